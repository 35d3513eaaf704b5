// jit.hpp -- run-time specialisation through NVRTC (K2 at run time).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <string>
#include <vector>

#include "network.hpp"
#include "ssa_params.h"

struct RbJitKernel {
  void* grid_kernel[3] = {nullptr, nullptr, nullptr};  // cudaKernel_t / __global__ function per RB_MODE_* schedule
  void* kernel_evc = nullptr;  // event-log mode: counting pass (filled by rb_jit_get_events / rb_prebuilt_get)
  void* kernel_evw = nullptr;  // event-log mode: writing pass
  unsigned block = 128;
  unsigned net_words = 0;
  unsigned static_smem = 0;  // bytes of static shared memory beyond the ensemble loop's own
  bool large = false;        // large form: the launch needs SsaRunParams::gtab
};

// Returns a compiled kernel for `net` on `device` (cached per process by source text).
// REBOP_ERR_LIMIT when the network is too large to specialise, REBOP_ERR_NVRTC when NVRTC is
// missing or the compilation fails (the message carries the log).
int rb_jit_get(const rebop_network& net, int device, RbJitKernel* out);
// The partial-propensity kernel of a mass-action network (REBOP_KERNEL_PDM; `low` from rb_pdm_lower).
struct RbPdmLowered;
int rb_jit_get_pdm(const rebop_network& net, const RbPdmLowered& low, int device, RbJitKernel* out);
// Same for the event-log kernels (a separate NVRTC program, compiled on first use).
int rb_jit_get_events(const rebop_network& net, int device, RbJitKernel* out);
// cudaFuncAttributeMaxDynamicSharedMemorySize of `kernel` on the current device, only ever raised (batches on
// several host threads share the kernels: lowering the limit under a concurrent launch would make it fail).
cudaError_t rb_raise_smem_limit(const void* kernel, size_t smem_bytes);
// Launch of an arbitrary entry point of a specialised kernel.
int rb_jit_launch_entry(void* kernel, unsigned block, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream);
// The kernel rebop_sysgen + nvcc compiled for this network at build time, if any (REBOP_ERR_INVALID if none).
int rb_prebuilt_get(const rebop_network& net, RbJitKernel* out);
// Source (and optionally the sm_100a cubin) of the specialised kernel; needs no GPU.
int rb_jit_compile(const rebop_network& net, std::string* source, std::vector<char>* cubin);
// Resident CTAs per SM of the kernel with this much dynamic shared memory.
int rb_jit_occupancy(const RbJitKernel& k, int mode, size_t smem_bytes, int* ctas_per_sm);
int rb_jit_launch(const RbJitKernel& k, int mode, const SsaRunParams& p, unsigned grid, size_t smem_bytes,
                  cudaStream_t stream);
