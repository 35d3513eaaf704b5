// jit.cpp -- NVRTC compilation and loading of generated kernels.
//
// libnvrtc is opened with dlopen so that the library itself loads (and its host-side API works)
// on machines without the CUDA toolkit; the cubin is loaded through the runtime's library API,
// so there is no link-time dependency on the driver either.
#include "jit.hpp"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "codegen.hpp"
#include "pdm.hpp"
#include "sysgen.hpp"

extern const char* const rb_embedded_names[];
extern const char* const rb_embedded_sources[];
extern const int rb_embedded_count;

namespace {

typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;

struct Nvrtc {
  void* handle = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  std::string error;
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* name : names) {
      n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (n.handle) break;
    }
    if (!n.handle) {
      n.error = "libnvrtc.so.12 not found";
      return;
    }
#define RB_SYM(field, sym)                                             \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, sym)); \
  if (!n.field) n.error = std::string("missing NVRTC symbol ") + sym;
    RB_SYM(CreateProgram, "nvrtcCreateProgram")
    RB_SYM(CompileProgram, "nvrtcCompileProgram")
    RB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    RB_SYM(GetProgramLog, "nvrtcGetProgramLog")
    RB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
    RB_SYM(GetCUBIN, "nvrtcGetCUBIN")
    RB_SYM(DestroyProgram, "nvrtcDestroyProgram")
    RB_SYM(GetErrorString, "nvrtcGetErrorString")
#undef RB_SYM
  });
  return n;
}

struct CacheEntry {
  cudaLibrary_t lib = nullptr;
  RbJitKernel k;
};
std::mutex g_mutex;
std::map<std::pair<int, std::string>, CacheEntry> g_cache;

}  // namespace

// ---- kernels compiled at build time (rebop_sysgen + nvcc) ----
static std::vector<RbPrebuilt>& prebuilt_registry() {
  static std::vector<RbPrebuilt> r;  // filled by static initialisers of the generated translation units
  return r;
}
void rb_register_prebuilt(const RbPrebuilt& entry) { prebuilt_registry().push_back(entry); }
const RbPrebuilt* rb_find_prebuilt(const std::string& key) {
  for (const RbPrebuilt& e : prebuilt_registry())
    if (key == e.key) return &e;
  return nullptr;
}
int rb_prebuilt_count() { return (int)prebuilt_registry().size(); }

extern "C" int rebop_b200_prebuilt_count(void) { return rb_prebuilt_count(); }
extern "C" int rebop_b200_prebuilt_name(int i, char* buf, size_t cap, size_t* needed) {
  if (i < 0 || i >= rb_prebuilt_count()) return rb_fail(REBOP_ERR_OUT_OF_RANGE, "prebuilt kernel index out of range");
  const std::string name = prebuilt_registry()[i].name;
  if (needed) *needed = name.size() + 1;
  if (buf && cap) {
    std::strncpy(buf, name.c_str(), cap);
    buf[cap - 1] = 0;
  }
  return REBOP_OK;
}
extern "C" int rebop_network_has_prebuilt(const rebop_network* net, int* yes) {
  if (!net || !yes) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::string why;
  *yes = rb_codegen_supported(*net, &why) && rb_find_prebuilt(rb_codegen_source(*net, "rb_ssa_jit", nullptr)) != nullptr;
  return REBOP_OK;
}

int rb_prebuilt_get(const rebop_network& net, RbJitKernel* out) {
  std::string why;
  if (!rb_codegen_supported(net, &why)) return rb_fail(REBOP_ERR_LIMIT, "network cannot be specialised: " + why);
  const RbPrebuilt* e = rb_find_prebuilt(rb_codegen_source(net, "rb_ssa_jit", nullptr));
  if (!e) return rb_fail(REBOP_ERR_INVALID, "no build-time kernel was generated for this network (see rebop_b200/systems)");
  for (int m = 0; m < 3; ++m) out->grid_kernel[m] = const_cast<void*>(e->grid_kernel[m]);
  out->kernel_evc = const_cast<void*>(e->kernel_evc);
  out->kernel_evw = const_cast<void*>(e->kernel_evw);
  out->block = e->block;
  out->net_words = e->net_words;
  out->static_smem = e->static_smem;
  out->large = rb_codegen_is_large(net);
  return REBOP_OK;
}

// NVRTC: source text -> sm_100a cubin.  Needs no GPU.
static int compile_to_cubin(const std::string& src, std::vector<char>* cubin) {
  Nvrtc& n = nvrtc();
  if (!n.error.empty()) return rb_fail(REBOP_ERR_NVRTC, "NVRTC unavailable: " + n.error);
  // REBOP_B200_JIT_DUMP=<dir>: keep the generated source (and the headers it includes) on disk and
  // compile it under that path, so that profilers (ncu --import-source) can show it.
  std::string prog_name = "rb_ssa_jit.cu";
  if (const char* dir = std::getenv("REBOP_B200_JIT_DUMP")) {
    char tag[32];
    std::snprintf(tag, sizeof tag, "%016zx", std::hash<std::string>{}(src));
    const std::string base = std::string(dir) + "/";
    auto dump = [](const std::string& path, const char* text) {
      if (FILE* fh = std::fopen(path.c_str(), "w")) {
        std::fputs(text, fh);
        std::fclose(fh);
      }
    };
    prog_name = base + "rb_ssa_jit_" + tag + ".cu";
    dump(prog_name, src.c_str());
    for (int i = 0; i < rb_embedded_count; ++i) dump(base + rb_embedded_names[i], rb_embedded_sources[i]);
  }
  nvrtcProgram prog = nullptr;
  nvrtcResult res = n.CreateProgram(&prog, src.c_str(), prog_name.c_str(), rb_embedded_count, rb_embedded_sources,
                                    rb_embedded_names);
  if (res != 0) return rb_fail(REBOP_ERR_NVRTC, std::string("nvrtcCreateProgram: ") + n.GetErrorString(res));
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-lineinfo", "--std=c++17"};
  res = n.CompileProgram(prog, 4, opts);
  if (res != 0) {
    size_t log_size = 0;
    n.GetProgramLogSize(prog, &log_size);
    std::string log(log_size, '\0');
    if (log_size) n.GetProgramLog(prog, &log[0]);
    n.DestroyProgram(&prog);
    return rb_fail(REBOP_ERR_NVRTC, std::string("nvrtcCompileProgram: ") + n.GetErrorString(res) + "\n" + log);
  }
  size_t cubin_size = 0;
  n.GetCUBINSize(prog, &cubin_size);
  cubin->resize(cubin_size);
  res = n.GetCUBIN(prog, cubin->data());
  n.DestroyProgram(&prog);
  if (res != 0 || cubin_size == 0) return rb_fail(REBOP_ERR_NVRTC, "nvrtcGetCUBIN failed");
  return REBOP_OK;
}

int rb_jit_compile(const rebop_network& net, std::string* source, std::vector<char>* cubin) {
  std::string why;
  if (!rb_codegen_supported(net, &why)) return rb_fail(REBOP_ERR_LIMIT, "network cannot be specialised: " + why);
  RbCodegenInfo info;
  const std::string src = rb_codegen_source(net, "rb_ssa_jit", &info);
  if (source) *source = src;
  if (cubin) return compile_to_cubin(src, cubin);
  return REBOP_OK;
}

static int jit_get_source(const std::string& src, const RbCodegenInfo& info, int device, RbJitKernel* out);

int rb_jit_get(const rebop_network& net, int device, RbJitKernel* out) {
  std::string why;
  if (!rb_codegen_supported(net, &why)) return rb_fail(REBOP_ERR_LIMIT, "network cannot be specialised: " + why);
  RbCodegenInfo info;
  const std::string src = rb_codegen_source(net, "rb_ssa_jit", &info);
  return jit_get_source(src, info, device, out);
}

int rb_jit_get_pdm(const rebop_network& net, const RbPdmLowered& low, int device, RbJitKernel* out) {
  RbCodegenInfo info;
  const std::string src = rb_codegen_pdm_source(net, low, "rb_ssa_jit", &info);
  return jit_get_source(src, info, device, out);
}

static int jit_get_source(const std::string& src, const RbCodegenInfo& info, int device, RbJitKernel* out) {
  std::lock_guard<std::mutex> lock(g_mutex);
  auto key = std::make_pair(device, src);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second.k;
    return REBOP_OK;
  }
  std::vector<char> cubin;
  int st = compile_to_cubin(src, &cubin);
  if (st) return st;

  CacheEntry e;
  cudaError_t err = cudaLibraryLoadData(&e.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(err));
  const char* const names[3] = {"rb_ssa_jit", "rb_ssa_jit_dyn", "rb_ssa_jit_dns"};
  for (int m = 0; m < 3 && err == cudaSuccess; ++m) {
    cudaKernel_t kernel = nullptr;
    err = cudaLibraryGetKernel(&kernel, e.lib, names[m]);
    e.k.grid_kernel[m] = kernel;
  }
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(err));
  e.k.block = info.block;
  e.k.net_words = info.net_words;
  e.k.static_smem = info.static_smem;
  e.k.large = info.large;
  g_cache[key] = e;
  *out = e.k;
  return REBOP_OK;
}

int rb_jit_get_events(const rebop_network& net, int device, RbJitKernel* out) {
  std::string why;
  if (!rb_codegen_supported(net, &why)) return rb_fail(REBOP_ERR_LIMIT, "network cannot be specialised: " + why);
  RbCodegenInfo info;
  const std::string src = rb_codegen_source(net, "rb_ssa_jit", &info, RB_VARIANT_EVENTS);
  std::lock_guard<std::mutex> lock(g_mutex);
  auto key = std::make_pair(device, src);
  auto it = g_cache.find(key);
  if (it == g_cache.end()) {
    std::vector<char> cubin;
    int st = compile_to_cubin(src, &cubin);
    if (st) return st;
    CacheEntry e;
    cudaError_t err = cudaLibraryLoadData(&e.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(err));
    cudaKernel_t evc = nullptr, evw = nullptr;
    err = cudaLibraryGetKernel(&evc, e.lib, "rb_ssa_jit_evc");
    if (err == cudaSuccess) err = cudaLibraryGetKernel(&evw, e.lib, "rb_ssa_jit_evw");
    if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(err));
    e.k.kernel_evc = evc;
    e.k.kernel_evw = evw;
    e.k.block = info.block;
    e.k.net_words = info.net_words;
    e.k.static_smem = info.static_smem;
    e.k.large = info.large;
    it = g_cache.emplace(key, e).first;
  }
  *out = it->second.k;
  return REBOP_OK;
}

cudaError_t rb_raise_smem_limit(const void* kernel, size_t smem_bytes) {
  static std::mutex mutex;
  static std::map<std::pair<int, const void*>, size_t> limit;
  int device = 0;
  cudaError_t err = cudaGetDevice(&device);
  if (err != cudaSuccess) return err;
  std::lock_guard<std::mutex> lock(mutex);
  size_t& cur = limit[std::make_pair(device, kernel)];
  if (smem_bytes <= cur) return cudaSuccess;
  err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (err == cudaSuccess) cur = smem_bytes;
  return err;
}

int rb_jit_launch_entry(void* kernel, unsigned block, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream) {
  cudaError_t err = rb_raise_smem_limit(kernel, smem_bytes);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
  void* args[] = {const_cast<SsaRunParams*>(&p)};
  err = cudaLaunchKernel(kernel, dim3(grid), dim3(block), args, smem_bytes, stream);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaLaunchKernel: ") + cudaGetErrorString(err));
  return REBOP_OK;
}

// ---- C ABI: inspection of the generated kernel (no GPU needed) ----
static int copy_out(const char* data, size_t size, char* buf, size_t cap, size_t* needed) {
  if (needed) *needed = size;
  if (buf && cap) std::memcpy(buf, data, size < cap ? size : cap);
  return REBOP_OK;
}

// Source / sm_100a cubin of the partial-propensity kernel (REBOP_KERNEL_PDM) of a mass-action network; needs no GPU.
static int pdm_compile(const rebop_network& net, std::string* source, std::vector<char>* cubin) {
  RbPdmLowered low;
  std::string why;
  int st = rb_pdm_lower(net, &low, &why);
  if (st) return rb_fail(st, why);
  const std::string src = rb_codegen_pdm_source(net, low, "rb_ssa_jit", nullptr);
  if (source) *source = src;
  if (cubin) return compile_to_cubin(src, cubin);
  return REBOP_OK;
}
extern "C" int rebop_network_codegen_pdm(const rebop_network* net, char* buf, size_t cap, size_t* needed) {
  if (!net) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::string src;
  int st = pdm_compile(*net, &src, nullptr);
  if (st) return st;
  return copy_out(src.c_str(), src.size() + 1, buf, cap, needed);
}
extern "C" int rebop_network_jit_cubin_pdm(const rebop_network* net, char* buf, size_t cap, size_t* needed) {
  if (!net) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::vector<char> cubin;
  int st = pdm_compile(*net, nullptr, &cubin);
  if (st) return st;
  return copy_out(cubin.data(), cubin.size(), buf, cap, needed);
}

extern "C" int rebop_network_codegen(const rebop_network* net, char* buf, size_t cap, size_t* needed) {
  if (!net) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::string src;
  int st = rb_jit_compile(*net, &src, nullptr);
  if (st) return st;
  return copy_out(src.c_str(), src.size() + 1, buf, cap, needed);
}

extern "C" int rebop_network_jit_cubin(const rebop_network* net, char* buf, size_t cap, size_t* needed) {
  if (!net) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::vector<char> cubin;
  int st = rb_jit_compile(*net, nullptr, &cubin);
  if (st) return st;
  return copy_out(cubin.data(), cubin.size(), buf, cap, needed);
}

int rb_jit_occupancy(const RbJitKernel& k, int mode, size_t smem_bytes, int* ctas_per_sm) {
  void* kernel = k.grid_kernel[mode];
  cudaError_t err = rb_raise_smem_limit(kernel, smem_bytes);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kernel, (int)k.block, smem_bytes);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaOccupancyMaxActiveBlocksPerMultiprocessor: ") + cudaGetErrorString(err));
  return REBOP_OK;
}

int rb_jit_launch(const RbJitKernel& k, int mode, const SsaRunParams& p, unsigned grid, size_t smem_bytes,
                  cudaStream_t stream) {
  void* kernel = k.grid_kernel[mode];
  cudaError_t err = rb_raise_smem_limit(kernel, smem_bytes);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
  void* args[] = {const_cast<SsaRunParams*>(&p)};
  err = cudaLaunchKernel(kernel, dim3(grid), dim3(k.block), args, smem_bytes, stream);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("cudaLaunchKernel: ") + cudaGetErrorString(err));
  return REBOP_OK;
}
