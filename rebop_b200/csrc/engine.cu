// engine.cu -- device-resident ensembles ("batches") and the launches that advance them.
//
// A batch is N instances of the reference's `Gillespie` struct (src/gillespie.rs:157-163)
// laid out for the GPU: species x[S][ldn] (int32, trajectory-contiguous), time t[ldn],
// xoshiro256++ state rng[4][ldn].  State persists between calls so that repeated
// advance_until / run_grid calls continue the same random streams, exactly like repeated
// calls on the CPU struct.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jit.hpp"
#include "network.hpp"
#include "samples.h"
#include "ssa_params.h"
#include "pdm.hpp"
#include "ssa_table.h"

#define RB_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t err__ = (call);                                                                \
    if (err__ != cudaSuccess)                                                                  \
      return rb_fail(REBOP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));   \
  } while (0)

static std::atomic<uint64_t> g_kernel_launches{0};
extern "C" uint64_t rebop_b200_kernel_launches(void) { return g_kernel_launches.load(); }

struct rebop_batch {
  rebop_network net;
  int device = 0;
  size_t n = 0, ldn = 0;
  int* d_x = nullptr;
  int* d_x0 = nullptr;  // [S] staging of a broadcast initial state
  double* d_t = nullptr;
  rb_u64* d_rng = nullptr;
  rb_u64* d_seeds = nullptr;
  rb_u64 seed_base = 0;
  unsigned seed_mode = 0;  // pending seeding for the next launch (0 = streams already live)
  void* d_out = nullptr;    // samples [rows][ldn] in the batch's sample type; event-log mode: int32 [n_save][total_rows]
  size_t out_capacity = 0;  // bytes
  int sample_bytes = 4;     // REBOP_SAMPLES_*: 2, 4 or 8
  int* d_raw = nullptr;     // lanes-claim-trajectories schedules: per-trajectory records [n][points][n_save] (see samples.cu)
  size_t raw_capacity = 0;  // int32 elements
  int* d_stat32 = nullptr;  // static schedule with a sample type other than int32: the kernel's int32 rows before conversion
  size_t stat32_capacity = 0;
  rb_u32* d_progress = nullptr;  // [ldn] RB_PROGRESS_*
  uint32_t out_rows = 0;    // (nb_steps+1) * n_save of the last run_grid
  uint32_t out_n_save = 0, out_nb_steps = 0;
  rb_u64* d_counters = nullptr;  // [0] events, [1] status (low 32 bits), [2] lane slots
  rb_u64* h_counters = nullptr;  // page-locked mirror (asynchronous launches report through it)
  rb_i64* d_sums = nullptr;
  size_t sums_capacity = 0;
  bool sums_ready = false;       // d_sums holds the row sums of the current samples (fused into the finishing kernel)
  bool async_pending = false;    // a launch on a caller-owned stream has not been accounted for yet
  RbTables* d_tables = nullptr;  // table-driven kernel: this batch's network image on the device
  bool tables_uploaded = false;
  rb_u64* d_pdm = nullptr;       // partial-propensity kernel: this batch's tables for the reaction choice on the device
  bool pdm_uploaded = false;
  RbPdmLowered pdm;              // host lowering (derived constants travel in the launch parameters)
  // a call cut short by the watchdog (REBOP_ERR_ITER_CAP): repeating it with the same arguments continues it
  struct Pending {
    bool active = false;
    int kind = 0;  // 0 advance_until, 1 run_grid
    double tmax = 0.0;
    uint32_t nb_steps = 0, n_save = 0;
    std::vector<uint32_t> save;
    unsigned segment = 0;  // run_grid: first segment that is not complete
    uint64_t events = 0, lane_slots = 0;
    float ms = 0.f;
  } pend;
  cudaStream_t stream = nullptr;      // the stream work is issued on
  cudaStream_t own_stream = nullptr;  // created with the batch; `stream` may point elsewhere
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;  // around the ensemble loop (ev0, ev1) and the sample finishing (ev1, ev2)
  uint64_t events_total = 0, events_last = 0, lane_slots_last = 0;
  int kernel_pref = REBOP_KERNEL_AUTO, kernel_used = REBOP_KERNEL_AUTO;
  uint32_t max_iters = 0;
  float last_ms = 0.f, last_finish_ms = 0.f;
  RbTables tables;
  bool tables_ok = false;
  std::string tables_error;
  int max_smem_optin = 0, sm_count = 0;
  int schedule = 0;               // 0 auto, 1 static (ring-staged coalesced samples), 2 lanes claim trajectories (sparse
                                  // samples), 3 lanes claim trajectories (dense samples)
  int mode_last = RB_MODE_STATIC;
  std::vector<int64_t> x_first;   // counts of the first trajectory as last uploaded (schedule heuristic)
  bool x_nonneg = true;           // every count uploaded so far was >= 0 (large specialised kernels need it)
  rb_u32* d_gtab = nullptr;       // large specialised kernels: reaction records + saved-species list
  size_t gtab_capacity = 0;       // words
  std::vector<rb_u32> h_gtab;
  cudaStream_t copy_stream = nullptr;  // run_grid(host_out): result rows go to the host while the next segment runs
  char* h_stage[2] = {nullptr, nullptr};  // page-locked staging for results that go to PAGEABLE host memory (see copy_rows_to_host)
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  double* d_grid_t = nullptr;      // grid times of the current run_grid launch
  size_t grid_t_capacity = 0;
  // event-log mode (nb_steps = 0)
  rb_u32* d_ev_counts = nullptr;
  rb_u64* d_ev_offsets = nullptr;
  double* d_ev_times = nullptr;
  size_t ev_times_capacity = 0;
  int* d_ev_out = nullptr;  // int32 [n_save][total_rows]
  size_t ev_out_capacity = 0;
  std::vector<uint64_t> ev_offsets;  // [n + 1] after run_events
  uint32_t ev_n_save = 0;
};

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------
__global__ void rb_fill_state_kernel(int* x, double* t, const int* x0, unsigned n_species,
                                     unsigned ldn, unsigned n, double t0, int fill_x, int fill_t) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ldn) return;
  if (fill_x)
    for (unsigned s = 0; s < n_species; ++s) x[(size_t)s * ldn + i] = i < n ? x0[s] : 0;
  if (fill_t) t[i] = t0;
}

// get_species: x[S][ldn] int32 -> rows [n][S] int64 (the layout the C ABI returns), transposed on the device.
__global__ void rb_species_rows_kernel(const int* __restrict__ x, long long* __restrict__ rows, unsigned n_species, unsigned ldn,
                                       unsigned n) {
  const size_t total = (size_t)n * n_species;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned traj = (unsigned)(i / n_species), sp = (unsigned)(i - (size_t)traj * n_species);
    rows[i] = x[(size_t)sp * ldn + traj];
  }
}

// K5: SmallRng::seed_from_u64 for every trajectory (src/gillespie.rs:184,190): SplitMix64 fills the xoshiro256++
// state.  A kernel of its own, so that every trajectory is seeded whatever the main launch gets to (a dynamic
// launch stopped by the watchdog leaves unclaimed trajectories untouched).
__global__ void rb_seed_kernel(rb_u64* rng, const rb_u64* seeds, rb_u64 seed_base, unsigned n, unsigned ldn) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rb_u64 st = seeds ? seeds[i] : seed_base + i;
  for (int w = 0; w < 4; ++w) {
    st += 0x9e3779b97f4a7c15ull;
    rb_u64 z = st;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    rng[(size_t)w * ldn + i] = z ^ (z >> 31);
  }
}

// K4: exact integer sum and sum of squares of every sample row over the trajectories, for samples of any of the three
// sample types (the claiming schedules get these sums from the finishing kernel; this one serves the static schedule).
// rows are 128-byte aligned (ldn % 32 == 0); one CTA reduces a segment of one row with 16-byte loads.
template <typename T>
__global__ void __launch_bounds__(256) rb_row_sums_kernel(const T* __restrict__ samples, unsigned n, unsigned ldn,
                                                          rb_i64* __restrict__ sums, unsigned n_rows) {
  constexpr unsigned per = 16u / (unsigned)sizeof(T);
  const unsigned row = blockIdx.x;  // rows on x: (nb_steps + 1) * n_save may exceed the 65535 limit of the other grid dimensions
  const T* base = samples + (size_t)row * ldn;
  const int4* src = reinterpret_cast<const int4*>(base);
  const unsigned nv = n / per;
  rb_i64 s1 = 0;
  rb_u64 s2 = 0;
  for (unsigned i = blockIdx.y * blockDim.x + threadIdx.x; i < nv; i += gridDim.y * blockDim.x) {
    const int4 v = __ldcs(src + i);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (unsigned j = 0; j < per; ++j) {
      const rb_i64 x = (rb_i64)e[j];
      s1 += x;
      s2 += (rb_u64)(x * x);
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < n - nv * per) {
    const rb_i64 x = (rb_i64)base[nv * per + threadIdx.x];
    s1 += x;
    s2 += (rb_u64)(x * x);
  }
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, off);
    s2 += __shfl_down_sync(0xffffffffu, s2, off);
  }
  __shared__ rb_i64 w1[8];
  __shared__ rb_u64 w2[8];
  if ((threadIdx.x & 31u) == 0) {
    w1[threadIdx.x >> 5] = s1;
    w2[threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      s1 += w1[w];
      s2 += w2[w];
    }
    atomicAdd(reinterpret_cast<rb_u64*>(sums) + row, (rb_u64)s1);
    atomicAdd(reinterpret_cast<rb_u64*>(sums) + n_rows + row, s2);
  }
}

// Sample rows [rows][ldn] of one sample type into dense rows [rows][n] of another (the widening / narrowing the
// host accessors offer, done on the device instead of in a host loop).
template <typename In, typename Out>
__global__ void __launch_bounds__(256) rb_rows_retype_kernel(const In* __restrict__ in, Out* __restrict__ out, unsigned n,
                                                             unsigned ldn, unsigned rows) {
  const size_t total = (size_t)rows * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / n;
    out[i] = (Out)in[r * ldn + (i - r * n)];
  }
}

// FP64 issue-rate probe: 8 independent chains per thread, alternating non-fused add / multiply.
__global__ void __launch_bounds__(256) rb_fp64_probe_kernel(double* sink, int iters, long long* clocks) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
  double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
  const double m = 1.0 - 1e-12, c = 1e-12;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a0 = __dmul_rn(a0, m); a1 = __dadd_rn(a1, c); a2 = __dmul_rn(a2, m); a3 = __dadd_rn(a3, c);
      a4 = __dmul_rn(a4, m); a5 = __dadd_rn(a5, c); a6 = __dmul_rn(a6, m); a7 = __dadd_rn(a7, c);
    }
  }
  const long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// SM clock probe: one warp per SM spins on a dependent integer chain; cycles (clock64) over nanoseconds
// (globaltimer).  Run right after the FP64 probe, while the clocks are up.
__global__ void rb_clock_probe_kernel(int iters, unsigned long long* out) {
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long c0 = clock64();
  unsigned v = threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) v = v * 1664525u + 1013904223u;
  const long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = (unsigned long long)(c1 - c0);
    out[2 * blockIdx.x + 1] = g1 - g0 + (v == 0xdeadbeefu);
  }
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
static int select_device(int device) {
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0)
    return rb_fail(REBOP_ERR_CUDA, std::string("no CUDA device available (this engine has no CPU fallback): ") +
                                       (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0"));
  if (device < 0 || device >= count) return rb_fail(REBOP_ERR_INVALID, "device index out of range");
  RB_CUDA(cudaSetDevice(device));
  return REBOP_OK;
}

extern "C" int rebop_b200_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

static int upload_x0(rebop_batch* b, const int64_t* x0, int per_traj) {
  const uint32_t S = b->net.n_species;
  if (S == 0) return REBOP_OK;
  if (!x0) return rb_fail(REBOP_ERR_INVALID, "x0 is NULL");
  const size_t count = per_traj ? b->n * S : S;
  bool nonneg = true;
  for (size_t i = 0; i < count; ++i) {
    if (x0[i] < INT32_MIN || x0[i] > INT32_MAX)
      return rb_fail(REBOP_ERR_LIMIT, "initial species count outside the int32 range carried on the device");
    nonneg = nonneg && x0[i] >= 0;
  }
  b->x_nonneg = nonneg;
  b->x_first.assign(x0, x0 + S);
  if (!per_traj) {
    std::vector<int> h(S);
    for (uint32_t s = 0; s < S; ++s) h[s] = (int)x0[s];
    // a copy from pageable memory returns once the source has been staged, so `h` may go out of scope
    RB_CUDA(cudaMemcpyAsync(b->d_x0, h.data(), S * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    rb_fill_state_kernel<<<(unsigned)((b->ldn + 255) / 256), 256, 0, b->stream>>>(
        b->d_x, b->d_t, b->d_x0, S, (unsigned)b->ldn, (unsigned)b->n, 0.0, 1, 0);
    RB_CUDA(cudaGetLastError());
    ++g_kernel_launches;
  } else {
    std::vector<int> h((size_t)S * b->ldn, 0);
    for (size_t n = 0; n < b->n; ++n)
      for (uint32_t s = 0; s < S; ++s) h[(size_t)s * b->ldn + n] = (int)x0[n * S + s];
    RB_CUDA(cudaMemcpyAsync(b->d_x, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  }
  b->pend.active = false;  // new state: a call cut short earlier cannot be continued
  return REBOP_OK;
}

static int upload_seeds(rebop_batch* b, const uint64_t* seeds, uint64_t seed_base) {
  if (seeds) {
    if (!b->d_seeds) RB_CUDA(cudaMalloc(&b->d_seeds, b->ldn * sizeof(rb_u64)));
    RB_CUDA(cudaMemcpyAsync(b->d_seeds, seeds, b->n * sizeof(rb_u64), cudaMemcpyHostToDevice, b->stream));
    RB_CUDA(cudaStreamSynchronize(b->stream));  // the caller's array may be page-locked: the copy must have read it before we return
    b->seed_mode = 1;
  } else {
    b->seed_base = seed_base;
    b->seed_mode = 2;
  }
  b->pend.active = false;
  return REBOP_OK;
}

// ---------------------------------------------------------------------------
// batch life cycle
// ---------------------------------------------------------------------------
extern "C" void rebop_batch_destroy(rebop_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  if (b->own_stream && b->own_stream != b->stream) cudaStreamSynchronize(b->own_stream);
  cudaFree(b->d_x); cudaFree(b->d_t); cudaFree(b->d_rng); cudaFree(b->d_seeds);
  cudaFree(b->d_out); cudaFree(b->d_counters); cudaFree(b->d_sums); cudaFree(b->d_gtab);
  cudaFree(b->d_raw); cudaFree(b->d_stat32); cudaFree(b->d_progress); cudaFree(b->d_tables); cudaFree(b->d_pdm); cudaFree(b->d_x0);
  if (b->h_counters) cudaFreeHost(b->h_counters);
  cudaFree(b->d_ev_counts); cudaFree(b->d_ev_offsets); cudaFree(b->d_ev_times); cudaFree(b->d_ev_out); cudaFree(b->d_grid_t);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev2) cudaEventDestroy(b->ev2);
  for (int i = 0; i < 2; ++i) {
    if (b->h_stage[i]) cudaFreeHost(b->h_stage[i]);
    if (b->ev_stage[i]) cudaEventDestroy(b->ev_stage[i]);
  }
  if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
  if (b->own_stream) cudaStreamDestroy(b->own_stream);
  delete b;
}

extern "C" int rebop_batch_create(const rebop_network* net, int device, size_t n_traj, const int64_t* x0,
                                  int x0_per_trajectory, const uint64_t* seeds, uint64_t seed_base,
                                  rebop_batch** out) {
  if (!net || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (n_traj == 0 || n_traj > 0x7fffffffu) return rb_fail(REBOP_ERR_INVALID, "n_traj must be in [1, 2^31)");
  int st = select_device(device);
  if (st) return st;
  rebop_batch* b = new rebop_batch();
  b->net = *net;
  b->device = device;
  b->n = n_traj;
  b->ldn = (n_traj + 31) / 32 * 32;
  st = rb_lower_tables(b->net, &b->tables);
  b->tables_ok = (st == REBOP_OK);
  if (!b->tables_ok) b->tables_error = rebop_b200_last_error();
#define RB_CREATE_CUDA(call)                                                                     \
  do {                                                                                           \
    cudaError_t err__ = (call);                                                                  \
    if (err__ != cudaSuccess) {                                                                  \
      rb_set_error(std::string(#call) + ": " + cudaGetErrorString(err__));                       \
      rebop_batch_destroy(b);                                                                    \
      return REBOP_ERR_CUDA;                                                                     \
    }                                                                                            \
  } while (0)
  const size_t S = net->n_species ? net->n_species : 1;
  RB_CREATE_CUDA(cudaDeviceGetAttribute(&b->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  RB_CREATE_CUDA(cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device));
  RB_CREATE_CUDA(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking));
  b->stream = b->own_stream;
  RB_CREATE_CUDA(cudaEventCreate(&b->ev0));
  RB_CREATE_CUDA(cudaEventCreate(&b->ev1));
  RB_CREATE_CUDA(cudaEventCreate(&b->ev2));
  RB_CREATE_CUDA(cudaMalloc(&b->d_x, S * b->ldn * sizeof(int)));
  RB_CREATE_CUDA(cudaMalloc(&b->d_t, b->ldn * sizeof(double)));
  RB_CREATE_CUDA(cudaMalloc(&b->d_rng, 4 * b->ldn * sizeof(rb_u64)));
  RB_CREATE_CUDA(cudaMalloc(&b->d_x0, S * sizeof(int)));
  RB_CREATE_CUDA(cudaMalloc(&b->d_progress, b->ldn * sizeof(rb_u32)));
  RB_CREATE_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b->h_counters), 4 * sizeof(rb_u64), cudaHostAllocDefault));
  RB_CREATE_CUDA(cudaMalloc(&b->d_counters, 4 * sizeof(rb_u64)));
  RB_CREATE_CUDA(cudaMemsetAsync(b->d_counters, 0, 4 * sizeof(rb_u64), b->stream));
  RB_CREATE_CUDA(cudaMemsetAsync(b->d_rng, 0, 4 * b->ldn * sizeof(rb_u64), b->stream));
  RB_CREATE_CUDA(cudaMemsetAsync(b->d_t, 0, b->ldn * sizeof(double), b->stream));
  RB_CREATE_CUDA(cudaMemsetAsync(b->d_x, 0, S * b->ldn * sizeof(int), b->stream));
#undef RB_CREATE_CUDA
  st = upload_x0(b, x0, x0_per_trajectory);
  if (!st) st = upload_seeds(b, seeds, seed_base);
  if (st) {
    rebop_batch_destroy(b);
    return st;
  }
  *out = b;
  return REBOP_OK;
}

extern "C" int rebop_batch_set_kernel(rebop_batch* b, int kind) {
  if (!b || kind < REBOP_KERNEL_AUTO || kind > REBOP_KERNEL_PDM) return rb_fail(REBOP_ERR_INVALID, "bad kernel kind");
  b->kernel_pref = kind;
  return REBOP_OK;
}
extern "C" int rebop_batch_get_kernel(const rebop_batch* b, int* kind) {
  if (!b || !kind) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *kind = b->kernel_used;
  return REBOP_OK;
}
extern "C" int rebop_batch_set_rates(rebop_batch* b, const double* k, size_t n_reactions) {
  if (!b || (!k && n_reactions)) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (n_reactions != b->net.rx.size()) return rb_fail(REBOP_ERR_INVALID, "n_reactions does not match the network");
  for (size_t r = 0; r < n_reactions; ++r)
    if (!b->net.rx[r].is_expr) b->net.rx[r].k = k[r];
  int st = rb_lower_tables(b->net, &b->tables);
  b->tables_ok = (st == REBOP_OK);
  b->tables_uploaded = false;
  b->pdm_uploaded = false;
  if (!b->tables_ok) b->tables_error = rebop_b200_last_error();
  return REBOP_OK;
}
extern "C" int rebop_batch_set_schedule(rebop_batch* b, int schedule) {
  if (!b || schedule < 0 || schedule > 3)
    return rb_fail(REBOP_ERR_INVALID, "schedule must be 0 (auto), 1 (static), 2 (dynamic, sparse samples) or 3 (dynamic, dense samples)");
  b->schedule = schedule;
  return REBOP_OK;
}
extern "C" int rebop_batch_get_schedule(const rebop_batch* b, int* schedule_used) {
  if (!b || !schedule_used) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *schedule_used = b->mode_last + 1;
  return REBOP_OK;
}
extern "C" int rebop_batch_set_max_iters(rebop_batch* b, uint32_t max_iters) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  b->max_iters = max_iters;
  return REBOP_OK;
}
extern "C" int rebop_batch_size(const rebop_batch* b, size_t* n_traj) {
  if (!b || !n_traj) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *n_traj = b->n;
  return REBOP_OK;
}
// Launches on a caller-owned stream (rebop_batch_set_stream) do not block the host: their counters and status arrive
// in the page-locked mirror and are accounted for here, after the stream has drained.
static int account_async(rebop_batch* b) {
  if (!b->async_pending) return REBOP_OK;
  b->async_pending = false;
  b->events_last = b->h_counters[0];
  b->lane_slots_last = b->h_counters[2];
  b->events_total += b->h_counters[0];
  const rb_u64 status = b->h_counters[1];
  if (status & RB_STATUS_NARROW)
    return rb_fail(REBOP_ERR_LIMIT, "a sample does not fit the batch's 16-bit sample type (rebop_batch_set_sample_dtype)");
  if (status & RB_STATUS_ITER_CAP)
    return rb_fail(REBOP_ERR_ITER_CAP, "a trajectory hit the per-launch iteration cap before reaching its target time");
  return REBOP_OK;
}

static int sync_stream(rebop_batch* b) {
  RB_CUDA(cudaStreamSynchronize(b->stream));
  return account_async(b);
}

extern "C" int rebop_batch_synchronize(rebop_batch* b) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  return sync_stream(b);
}

extern "C" int rebop_batch_get_stream(const rebop_batch* b, void** stream) {
  if (!b || !stream) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *stream = b->stream;
  return REBOP_OK;
}
extern "C" int rebop_batch_set_stream(rebop_batch* b, void* stream) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  int st = sync_stream(b);  // hand over with nothing in flight on the old stream
  b->stream = stream ? static_cast<cudaStream_t>(stream) : b->own_stream;
  return st;
}

extern "C" int rebop_b200_host_alloc(size_t bytes, void** out) {
  if (!out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (rebop_b200_device_count() == 0)
    return rb_fail(REBOP_ERR_CUDA, "no CUDA device available (page-locked memory needs a driver)");
  RB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return REBOP_OK;
}
extern "C" int rebop_b200_host_free(void* p) {
  if (p) RB_CUDA(cudaFreeHost(p));
  return REBOP_OK;
}

extern "C" int rebop_batch_seed(rebop_batch* b, const uint64_t* seeds, uint64_t seed_base) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  return upload_seeds(b, seeds, seed_base);
}

extern "C" int rebop_batch_get_time(rebop_batch* b, double* t) {
  if (!b || !t) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  RB_CUDA(cudaMemcpyAsync(t, b->d_t, b->n * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  return sync_stream(b);
}

extern "C" int rebop_batch_set_time(rebop_batch* b, double t) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  rb_fill_state_kernel<<<(unsigned)((b->ldn + 255) / 256), 256, 0, b->stream>>>(
      b->d_x, b->d_t, nullptr, 0, (unsigned)b->ldn, (unsigned)b->n, t, 0, 1);
  RB_CUDA(cudaGetLastError());
  ++g_kernel_launches;
  b->pend.active = false;
  return REBOP_OK;
}

extern "C" int rebop_batch_get_species(rebop_batch* b, int64_t* species) {
  if (!b || !species) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  const uint32_t S = b->net.n_species;
  if (S == 0) return REBOP_OK;
  long long* d_rows = nullptr;  // [n][S] int64, transposed and widened on the device
  const size_t total = b->n * S;
  RB_CUDA(cudaMallocAsync(&d_rows, total * sizeof(long long), b->stream));
  const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)b->sm_count * 16);
  rb_species_rows_kernel<<<blocks, 256, 0, b->stream>>>(b->d_x, d_rows, S, (unsigned)b->ldn, (unsigned)b->n);
  ++g_kernel_launches;
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemcpyAsync(species, d_rows, total * sizeof(long long), cudaMemcpyDeviceToHost, b->stream);
  cudaFreeAsync(d_rows, b->stream);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("rebop_batch_get_species: ") + cudaGetErrorString(err));
  return sync_stream(b);
}

extern "C" int rebop_batch_set_species(rebop_batch* b, const int64_t* species, int per_trajectory) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  return upload_x0(b, species, per_trajectory);
}

// ---------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------
static unsigned pow2_floor(unsigned v) {
  unsigned p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

// Ring depth: as many grid points as fit next to the kernel's other shared memory while leaving
// room for `ctas_per_sm` resident CTAs; a power of two in [1, 32], no deeper than the grid.
static unsigned choose_ring_depth(const rebop_batch* b, unsigned block, unsigned net_words, unsigned n_save,
                                  unsigned n_points, unsigned ctas_per_sm) {
  if (n_save == 0) return 1;
  const size_t fixed = RB_STATIC_SMEM_BYTES + 8u * block + 1024 + 4u * net_words;  // + the loop's grid-time slots + per-CTA reservation
  const size_t per_depth = (size_t)(block / 32u) * n_save * 32u * 4u;
  const size_t budget = (size_t)b->max_smem_optin / (ctas_per_sm ? ctas_per_sm : 1);
  if (budget < fixed + per_depth) return 0;  // no room for a ring: samples go straight to global memory
  unsigned depth = (unsigned)std::min<size_t>(32, (budget - fixed) / per_depth);
  depth = pow2_floor(depth ? depth : 1);
  unsigned cap = 1;
  while (cap < n_points && cap < 32) cap *= 2;
  return std::min(depth, cap);
}

// Reaction records (see ssa_params.h) + saved-species list, rebuilt and uploaded per launch (rate constants
// may have changed).
static int upload_gtab(rebop_batch* b, const uint32_t* save_idx, uint32_t n_save) {
  const uint32_t S = b->net.n_species;
  const size_t R = b->net.rx.size();
  b->h_gtab.assign(R * RB_GTAB_WORDS_PER_REACTION + std::max<uint32_t>(n_save, 1), 0u);
  for (size_t r = 0; r < R; ++r) {
    const RbReaction& rx = b->net.rx[r];
    rb_u32* w = b->h_gtab.data() + r * RB_GTAB_WORDS_PER_REACTION;
    std::memcpy(w, &rx.k, 8);
    const size_t nt = rx.term_idx.size();
    bool record = !rx.is_expr && nt <= 2;
    for (size_t j = 0; j < nt; ++j) record = record && rx.term_idx[j] <= 0xffffu && rx.term_exp[j] <= 0xffu;
    if (record) {
      w[2] = (nt > 0 ? rx.term_idx[0] : 0u) | ((nt > 1 ? rx.term_idx[1] : 0u) << 16);
      w[3] = (nt > 0 ? rx.term_exp[0] : 0u) | ((nt > 1 ? rx.term_exp[1] : 0u) << 8) | ((rb_u32)nt << 16);
    } else {
      w[3] = 0xffu << 16;
    }
    unsigned q = 0;
    for (uint32_t sp = 0; sp < S; ++sp) {
      if (rx.diff[sp] == 0) continue;
      if (q < 4) {
        w[4 + q / 2] |= sp << (16 * (q & 1));
        w[6 + q / 2] |= ((rb_u32)(uint16_t)(int16_t)rx.diff[sp]) << (16 * (q & 1));
      }
      ++q;
    }
    if (q > 4) w[3] |= 1u << 24;
  }
  for (uint32_t j = 0; j < n_save; ++j) b->h_gtab[R * RB_GTAB_WORDS_PER_REACTION + j] = save_idx[j];
  if (b->h_gtab.size() > b->gtab_capacity) {
    if (b->d_gtab) RB_CUDA(cudaFree(b->d_gtab));
    b->d_gtab = nullptr;
    b->gtab_capacity = 0;
    RB_CUDA(cudaMalloc(&b->d_gtab, b->h_gtab.size() * sizeof(rb_u32)));
    b->gtab_capacity = b->h_gtab.size();
  }
  RB_CUDA(cudaMemcpyAsync(b->d_gtab, b->h_gtab.data(), b->h_gtab.size() * sizeof(rb_u32), cudaMemcpyHostToDevice, b->stream));
  return REBOP_OK;
}

// Auto schedule.  Lanes claim trajectories in both automatic choices (no lane idles behind the slowest trajectory
// of its warp, absorbed trajectories stop at once); what differs is when the uniform is drawn.  SPARSE draws both
// random words ahead of the propensities and steps the stream back on a grid crossing: best when crossings are
// rare.  DENSE draws the uniform once the event is known to fire: best when most passes of a warp contain a
// crossing.  The number of events is not known in advance; the total propensity of the first trajectory's initial
// state times the horizon is a (low) estimate that orders the workloads correctly: SIR 0.04 events per sample,
// Dimers 3, Vilar 5, Michaelis-Menten 15.  The static schedule (ring-staged rows, thread n = trajectory n) is kept
// for callers that ask for it.
static int rb_auto_mode(const rebop_batch* b, double tmax, unsigned n_save, unsigned n_points) {
  if (n_save == 0) return RB_MODE_SPARSE;
  double a0 = 0.0;
  for (const RbReaction& rx : b->net.rx) {
    if (rx.is_expr) return RB_MODE_SPARSE;
    double a = rx.k;
    for (size_t j = 0; j < rx.term_idx.size(); ++j) {
      const double x = rx.term_idx[j] < b->x_first.size() ? (double)b->x_first[rx.term_idx[j]] : 0.0;
      for (uint32_t f = 0; f < std::max<uint32_t>(1u, rx.term_exp[j]); ++f) a *= x - f;
    }
    if (a > 0.0) a0 += a;
  }
  return a0 * tmax >= (double)n_save * n_points ? RB_MODE_SPARSE : RB_MODE_DENSE;
}

// Pending seeding (rebop_batch_create / rebop_batch_seed) is applied on the stream before the next launch.
static int apply_seeding(rebop_batch* b) {
  if (b->seed_mode == 0) return REBOP_OK;
  rb_seed_kernel<<<(unsigned)((b->n + 255) / 256), 256, 0, b->stream>>>(b->d_rng, b->seed_mode == 1 ? b->d_seeds : nullptr,
                                                                      b->seed_base, (unsigned)b->n, (unsigned)b->ldn);
  RB_CUDA(cudaGetLastError());
  ++g_kernel_launches;
  b->seed_mode = 0;  // streams are live on the device from now on
  return REBOP_OK;
}

static void fill_params(const rebop_batch* b, SsaRunParams* p) {
  std::memset(p, 0, sizeof *p);
  p->x = b->d_x;
  p->t = b->d_t;
  p->rng = b->d_rng;
  p->events = b->d_counters;
  p->status = reinterpret_cast<rb_u32*>(b->d_counters + 1);
  p->work_next = reinterpret_cast<rb_u32*>(b->d_counters + 3);
  p->progress = b->d_progress;
  p->n_traj = (rb_u32)b->n;
  p->ldn = (rb_u32)b->ldn;
  p->max_iters = b->max_iters;
  p->bias_hi = 0x43300000u;
  p->exp_one = 0x3ffu;
  p->bias = 0x1.0p52 + 0x1.0p31;
  p->one_m_eps = 1.0 - 0x1.0p-53;
  for (int l = 0; l < 4; ++l) p->byte_sel[l] = 1 << (8 * l);
  for (size_t r = 0; r < b->net.rx.size() && r < RB_MAX_K; ++r) p->k[r] = b->net.rx[r].k;
  p->n_species = (int)b->net.n_species;
  p->n_reactions = (int)b->net.rx.size();
  p->arith = b->net.arith;
}

// The table-driven kernel reads its network from the batch's own device image (nothing process-global).
static int upload_tables(rebop_batch* b) {
  if (b->tables_uploaded) return REBOP_OK;
  if (!b->d_tables) RB_CUDA(cudaMalloc(&b->d_tables, sizeof(RbTables)));
  RB_CUDA(cudaMemcpyAsync(b->d_tables, &b->tables, sizeof(RbTables), cudaMemcpyHostToDevice, b->stream));
  b->tables_uploaded = true;
  return REBOP_OK;
}

// The partial-propensity kernel: lowering (pdm.hpp) and the device image of its choice tables, once per batch and set of
// rate constants.
static int upload_pdm(rebop_batch* b) {
  if (b->pdm_uploaded) return REBOP_OK;
  std::string why;
  int st = rb_pdm_lower(b->net, &b->pdm, &why);
  if (st) return rb_fail(st, why);
  if (b->d_pdm) RB_CUDA(cudaFree(b->d_pdm));
  b->d_pdm = nullptr;
  RB_CUDA(cudaMalloc(&b->d_pdm, b->pdm.image.size() * sizeof(uint64_t)));
  RB_CUDA(cudaMemcpyAsync(b->d_pdm, b->pdm.image.data(), b->pdm.image.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, b->stream));
  b->pdm_uploaded = true;
  return REBOP_OK;
}

// Build-time specialised, else NVRTC-specialised, else table-driven (use_jit = false).  `events` asks
// for the event-log entry points instead of the time-grid ones.
static int pick_kernel(rebop_batch* b, bool events, RbJitKernel* jit, bool* use_jit, int* jit_kind) {
  *use_jit = false;
  *jit_kind = REBOP_KERNEL_NVRTC;
  if (b->kernel_pref == REBOP_KERNEL_AUTO || b->kernel_pref == REBOP_KERNEL_PREBUILT) {
    int st = rb_prebuilt_get(b->net, jit);
    if (st == REBOP_OK) {
      *use_jit = true;
      *jit_kind = REBOP_KERNEL_PREBUILT;
    } else if (b->kernel_pref == REBOP_KERNEL_PREBUILT) {
      return st;
    }
  }
  if (!*use_jit && (b->kernel_pref == REBOP_KERNEL_AUTO || b->kernel_pref == REBOP_KERNEL_NVRTC)) {
    int st = events ? rb_jit_get_events(b->net, b->device, jit) : rb_jit_get(b->net, b->device, jit);
    if (st == REBOP_OK) {
      *use_jit = true;
    } else if (b->kernel_pref == REBOP_KERNEL_NVRTC) {
      return st;
    }
  }
  // large specialised kernels assume non-decreasing cumulative rates: k >= 0 and counts >= 0
  if (*use_jit && jit->large) {
    bool ok = b->x_nonneg;
    for (const RbReaction& rx : b->net.rx) ok = ok && !(rx.k < 0.0);
    if (!ok) {
      if (b->kernel_pref != REBOP_KERNEL_AUTO)
        return rb_fail(REBOP_ERR_LIMIT, "the specialised kernel for large networks needs rate constants >= 0 and counts >= 0");
      *use_jit = false;
    }
  }
  if (!*use_jit && !b->tables_ok) return rb_fail(REBOP_ERR_LIMIT, b->tables_error);
  return REBOP_OK;
}

template <class T>
static int ensure_capacity(T** ptr, size_t* capacity, size_t need) {
  if (need <= *capacity) return REBOP_OK;
  if (*ptr) RB_CUDA(cudaFree(*ptr));
  *ptr = nullptr;
  *capacity = 0;
  RB_CUDA(cudaMalloc(reinterpret_cast<void**>(ptr), need * sizeof(T)));
  *capacity = need;
  return REBOP_OK;
}

static bool is_async(const rebop_batch* b) { return b->stream != b->own_stream; }

// Start of an API call that launches: per-call counters.  On a caller-owned stream the counters keep accumulating
// until the caller synchronises (rebop_batch_synchronize reports them).
static int begin_call(rebop_batch* b) {
  if (!b->async_pending) RB_CUDA(cudaMemsetAsync(b->d_counters, 0, 4 * sizeof(rb_u64), b->stream));
  return REBOP_OK;
}

// End of an API call that launched: counters and status back to the host.  Blocks unless the stream is the caller's.
static int end_call(rebop_batch* b, const char* what) {
  RB_CUDA(cudaMemcpyAsync(b->h_counters, b->d_counters, 4 * sizeof(rb_u64), cudaMemcpyDeviceToHost, b->stream));
  b->async_pending = true;
  if (is_async(b)) return REBOP_OK;
  RB_CUDA(cudaStreamSynchronize(b->stream));
  b->async_pending = false;
  b->events_last = b->h_counters[0];
  b->lane_slots_last = b->h_counters[2];
  b->events_total += b->h_counters[0];
  if (b->h_counters[1] & RB_STATUS_NARROW)
    return rb_fail(REBOP_ERR_LIMIT, "a sample does not fit the batch's 16-bit sample type (rebop_batch_set_sample_dtype)");
  if (b->h_counters[1] & RB_STATUS_ITER_CAP) return rb_fail(REBOP_ERR_ITER_CAP, what);
  return REBOP_OK;
}

// One launch of the ensemble loop over grid points step_first..step_last (inclusive) of a grid of nb_steps steps
// (nb_steps = 0: the single target tmax), followed -- when samples are taken -- by the kernel that brings them into
// the result layout at row `row_first` of d_out.  Everything is stream-ordered; nothing here blocks the host.
static int launch(rebop_batch* b, double tmax, uint32_t nb_steps, uint32_t step_first, uint32_t step_last, bool sample,
                  uint32_t n_save, const uint32_t* save_idx, bool resuming, unsigned grid_points_total = 0) {
  const uint32_t S = b->net.n_species;
  {
    int st = apply_seeding(b);
    if (st) return st;
  }
  SsaRunParams p;
  fill_params(b, &p);
  p.tmax = tmax;
  p.nb_steps = nb_steps;
  p.step_first = step_first;
  p.step_last = step_last;
  p.n_save = sample ? n_save : 0;
  p.resuming = resuming ? 1u : 0u;
  const unsigned n_points = step_last - step_first + 1;
  const size_t rows = (size_t)n_points * p.n_save;
  const size_t row_first = (size_t)step_first * p.n_save;

  if (!resuming) RB_CUDA(cudaMemsetAsync(b->d_progress, 0, b->ldn * sizeof(rb_u32), b->stream));
  RB_CUDA(cudaMemsetAsync(b->d_counters + 3, 0, sizeof(rb_u64), b->stream));  // work counter of this launch

  // grid times, exactly as the binding computes them: tmax * i as f64 / nb_steps as f64 (src/pyo3_gillespie.rs:201)
  // (nb_steps = 0, Gillespie::advance_until: the table holds the single target, so that the kernels read their grid
  // times one way only)
  std::vector<double> grid_t(n_points, tmax);
  if (nb_steps > 0) {
    for (unsigned i = 0; i < n_points; ++i) {
      volatile double prod = tmax * (double)(step_first + i);  // volatile: one rounding per operation, no contraction
      grid_t[i] = prod / (double)nb_steps;
    }
  }
  {
    int st = ensure_capacity(&b->d_grid_t, &b->grid_t_capacity, (size_t)n_points);
    if (st) return st;
    RB_CUDA(cudaMemcpyAsync(b->d_grid_t, grid_t.data(), n_points * sizeof(double), cudaMemcpyHostToDevice, b->stream));
    p.grid_t = b->d_grid_t;
  }

  // --- schedule
  int schedule = b->schedule;
  if (const char* env = std::getenv("REBOP_B200_SCHEDULE")) {
    if (!std::strcmp(env, "static")) schedule = 1;
    if (!std::strcmp(env, "dynamic") || !std::strcmp(env, "sparse")) schedule = 2;
    if (!std::strcmp(env, "dense")) schedule = 3;
  }
  const int mode = schedule ? schedule - 1 : rb_auto_mode(b, tmax, p.n_save, grid_points_total ? grid_points_total : n_points);
  const bool claim = mode != RB_MODE_STATIC;

  RbJitKernel jit;
  bool use_jit = false;
  int jit_kind = REBOP_KERNEL_NVRTC;
  const bool use_pdm = b->kernel_pref == REBOP_KERNEL_PDM;
  if (use_pdm) {
    if (!b->x_nonneg) return rb_fail(REBOP_ERR_LIMIT, "the partial-propensity kernel (REBOP_KERNEL_PDM) needs counts >= 0");
    int st = upload_pdm(b);
    if (!st) st = rb_jit_get_pdm(b->net, b->pdm, b->device, &jit);
    if (st) return st;
    use_jit = true;
    jit_kind = REBOP_KERNEL_PDM;
  } else {
    int st = pick_kernel(b, false, &jit, &use_jit, &jit_kind);
    if (st) return st;
  }

  // --- where the kernel puts its samples
  int* static_rows = nullptr;  // static schedule: int32 [rows][ldn]
  if (p.n_save) {
    if (claim) {
      int st = ensure_capacity(&b->d_raw, &b->raw_capacity, b->n * rows);
      if (st) return st;
      p.out = b->d_raw;
    } else if (b->sample_bytes == 4) {
      static_rows = static_cast<int*>(b->d_out) + row_first * b->ldn;
      p.out = static_rows;
    } else {
      int st = ensure_capacity(&b->d_stat32, &b->stat32_capacity, rows * b->ldn);
      if (st) return st;
      static_rows = b->d_stat32;
      p.out = static_rows;
    }
  }

  RB_CUDA(cudaEventRecord(b->ev0, b->stream));
  if (use_pdm) {
    for (size_t i = 0; i < b->pdm.consts.size(); ++i) p.k[i] = b->pdm.consts[i];  // c_i, K_ij in the order the kernel reads them
    p.pdm = b->d_pdm;
  }
  if (use_jit) {
    // saved species: the specialised kernels take a bit mask and emit rows in ascending species order
    for (uint32_t j = 0; j < p.n_save && save_idx[j] < 128; ++j) p.save_mask[save_idx[j] >> 6] |= 1ull << (save_idx[j] & 63u);
    const unsigned block = jit.block;
    unsigned ctas = std::max(1u, 2048u / block / 2u);
    if (jit.large) {
      ctas = 2;
      int st = upload_gtab(b, save_idx, p.n_save);
      if (st) return st;
      p.gtab = b->d_gtab;
    }
    p.ring_depth = claim ? 0u : choose_ring_depth(b, block, jit.net_words + jit.static_smem / 4u, p.n_save, n_points, ctas);
    const size_t smem = RB_SSA_SMEM_BYTES(jit.net_words, block, p.ring_depth, p.n_save);
    if (smem + RB_STATIC_SMEM_BYTES + 8u * block + jit.static_smem > (size_t)b->max_smem_optin)  // 8 * block: the loop's grid-time slots
      return rb_fail(REBOP_ERR_LIMIT, "specialised kernel: species state does not fit in shared memory");
    unsigned grid = (unsigned)((b->n + block - 1) / block);
    if (claim) {
      int resident = 0;
      int st = rb_jit_occupancy(jit, mode, smem, &resident);
      if (st) return st;
      grid = std::min(grid, (unsigned)std::max(1, resident) * (unsigned)b->sm_count);
      p.dynamic = 1;
      p.n_launched = grid * block;
    }
    int st = rb_jit_launch(jit, mode, p, grid, smem, b->stream);
    if (st) return st;
    b->kernel_used = jit_kind;
  } else {
    if (p.n_save > RB_TAB_MAX_SAVE) return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: more than 1024 saved species");
    {
      int st = upload_gtab(b, save_idx, p.n_save);
      if (!st) st = upload_tables(b);
      if (st) return st;
      p.gtab = b->d_gtab;
      p.tables = b->d_tables;
    }
    const unsigned block = RB_TABLE_BLOCK;
    const unsigned net_words = S * block;
    p.ring_depth = claim ? 0u : choose_ring_depth(b, block, net_words, p.n_save, n_points, 8);
    const size_t smem = RB_SSA_SMEM_BYTES(net_words, block, p.ring_depth, p.n_save);
    if (smem + RB_STATIC_SMEM_BYTES + 8u * block > (size_t)b->max_smem_optin)
      return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: species state and sample ring do not fit in shared memory");
    unsigned grid = (unsigned)((b->n + block - 1) / block);
    if (claim) {
      int resident = 0;
      RB_CUDA(rb_table_occupancy(mode, smem, &resident));
      grid = std::min(grid, (unsigned)std::max(1, resident) * (unsigned)b->sm_count);
      p.dynamic = 1;
      p.n_launched = grid * block;
    }
    RB_CUDA(rb_table_launch(mode, p, grid, smem, b->stream));
    b->kernel_used = REBOP_KERNEL_TABLE;
  }
  RB_CUDA(cudaEventRecord(b->ev1, b->stream));
  ++g_kernel_launches;
  b->mode_last = mode;

  // --- samples into the result layout and sample type
  if (p.n_save) {
    char* dst = static_cast<char*>(b->d_out) + row_first * b->ldn * (size_t)b->sample_bytes;
    if (claim) {
      RbFinishParams q;
      q.raw = b->d_raw;
      q.progress = b->d_progress;
      q.out = dst;
      q.sums = b->d_sums + row_first;
      q.sums_stride = b->out_rows;
      q.status = p.status;
      q.n = (rb_u32)b->n;
      q.ldn = (rb_u32)b->ldn;
      q.n_points = n_points;
      q.n_save = p.n_save;
      q.rows = (rb_u32)rows;
      RB_CUDA(cudaMemsetAsync(b->d_sums + row_first, 0, rows * sizeof(rb_i64), b->stream));
      RB_CUDA(cudaMemsetAsync(b->d_sums + b->out_rows + row_first, 0, rows * sizeof(rb_i64), b->stream));
      static const bool bulk = [] {
        const char* env = std::getenv("REBOP_B200_BULK_STORE");
        return !(env && env[0] == '0');
      }();
      RB_CUDA(rb_samples_finish(q, b->sample_bytes, bulk, (unsigned)b->sm_count, b->stream));
      ++g_kernel_launches;
      b->sums_ready = true;
    } else if (b->sample_bytes != 4) {
      RB_CUDA(rb_rows_convert(static_rows, dst, rows * b->ldn, b->sample_bytes, p.status, (unsigned)b->sm_count, b->stream));
      ++g_kernel_launches;
    }
    RB_CUDA(cudaEventRecord(b->ev2, b->stream));
  }
  return REBOP_OK;
}

// Times of the last launch: ensemble loop (ev0..ev1) and sample finishing (ev1..ev2).  Needs the stream drained.
static void read_times(rebop_batch* b, bool sampled, float* loop_ms, float* finish_ms) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) == cudaSuccess) *loop_ms += ms;
  if (sampled && cudaEventElapsedTime(&ms, b->ev1, b->ev2) == cudaSuccess) *finish_ms += ms;
}

extern "C" int rebop_batch_advance_until(rebop_batch* b, double tmax) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  // a repeat of the call the watchdog cut short continues it: finished trajectories are left alone
  const bool resuming = b->pend.active && b->pend.kind == 0 && b->pend.tmax == tmax;
  b->pend.active = false;
  int st = begin_call(b);
  if (!st) st = launch(b, tmax, 0, 0, 0, false, 0, nullptr, resuming);
  if (st) return st;
  st = end_call(b, "a trajectory hit the per-launch iteration cap before reaching its target time (call again to continue)");
  if (!is_async(b)) {
    b->last_ms = b->last_finish_ms = 0.f;
    read_times(b, false, &b->last_ms, &b->last_finish_ms);
  }
  if (st == REBOP_ERR_ITER_CAP) {
    b->pend.active = true;
    b->pend.kind = 0;
    b->pend.tmax = tmax;
  }
  return st;
}

// Result rows [row_first, row_first + rows) to the caller's buffer (row r at host + r * host_ld elements).
//
// Page-locked destination (rebop_b200_host_alloc, cudaHostRegister): one strided copy, asynchronous on `stream`.
// Pageable destination (a numpy array, a Vec): the driver would stage such a copy through a small buffer of its own,
// one row at a time and synchronously (measured: 4 GB in 0.9 s).  Instead the rows travel in chunks through two
// page-locked staging buffers of the batch, and while chunk c is in flight a few host threads move chunk c-1 into
// the caller's pages (first touch of a fresh allocation is the expensive part and parallelises).  Blocks the
// calling thread until the rows have arrived; the device keeps running whatever was launched before.
#define RB_STAGE_BYTES ((size_t)64 << 20)
static int copy_rows_to_host(rebop_batch* b, void* host, size_t host_ld, size_t row_first, size_t rows, cudaStream_t stream,
                             bool may_block) {
  const size_t sb = (size_t)b->sample_bytes;
  const size_t row_bytes = b->n * sb;
  cudaPointerAttributes attr;
  bool pinned = false;
  if (cudaPointerGetAttributes(&attr, host) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
  else cudaGetLastError();
  static const bool staging = [] {
    const char* env = std::getenv("REBOP_B200_STAGED_COPY");
    return !(env && env[0] == '0');
  }();
  if (pinned || !staging || !may_block || rows == 0 || row_bytes == 0 || row_bytes > RB_STAGE_BYTES) {
    RB_CUDA(cudaMemcpy2DAsync(static_cast<char*>(host) + row_first * host_ld * sb, host_ld * sb,
                              static_cast<const char*>(b->d_out) + row_first * b->ldn * sb, b->ldn * sb, row_bytes, rows,
                              cudaMemcpyDeviceToHost, stream));
    return REBOP_OK;
  }
  for (int i = 0; i < 2; ++i) {
    if (!b->h_stage[i]) RB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b->h_stage[i]), RB_STAGE_BYTES, cudaHostAllocDefault));
    if (!b->ev_stage[i]) RB_CUDA(cudaEventCreateWithFlags(&b->ev_stage[i], cudaEventDisableTiming));
  }
  {
    // A fresh allocation is faulted in page by page while it is being filled; with transparent huge pages that is
    // one fault per 2 MB instead of per 4 KB.  Advice only: ignored where the kernel does not offer it.
    const uintptr_t lo = (reinterpret_cast<uintptr_t>(host) + row_first * host_ld * sb + 4095u) & ~(uintptr_t)4095u;
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(host) + (row_first + rows - 1) * host_ld * sb + row_bytes) & ~(uintptr_t)4095u;
    if (hi > lo + ((uintptr_t)4 << 20)) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
  }
  const size_t chunk_rows = std::max<size_t>(1, RB_STAGE_BYTES / row_bytes);
  static const unsigned n_threads = [] {
    // (first touch of fresh pages is what the threads wait on: measured 1043 / 803 / 666 ms per 4 GB with 8 / 16 / 32
    // threads on 16 cores, profiles/r2n_probes.log)
    unsigned t = std::max(1u, std::min(32u, 2u * std::thread::hardware_concurrency()));
    if (const char* env = std::getenv("REBOP_B200_COPY_THREADS")) t = (unsigned)std::max(1, std::atoi(env));
    return t;
  }();
  auto drain = [&](int slot, size_t r0, size_t nr) -> int {  // staging slot -> caller's rows [r0, r0 + nr)
    RB_CUDA(cudaEventSynchronize(b->ev_stage[slot]));
    char* dst = static_cast<char*>(host) + r0 * host_ld * sb;
    const char* src = b->h_stage[slot];
    auto part = [&](unsigned k) {
      // every thread takes a contiguous byte range of every row, so that neighbouring pages go to the same thread
      const size_t lo = row_bytes * k / n_threads, hi = row_bytes * (k + 1) / n_threads;
      for (size_t r = 0; r < nr; ++r) std::memcpy(dst + r * host_ld * sb + lo, src + r * row_bytes + lo, hi - lo);
    };
    if (n_threads == 1 || nr * row_bytes < ((size_t)1 << 20)) {
      for (unsigned k = 0; k < n_threads; ++k) part(k);
    } else {
      std::vector<std::thread> pool;
      for (unsigned k = 1; k < n_threads; ++k) pool.emplace_back(part, k);
      part(0);
      for (std::thread& t : pool) t.join();
    }
    return REBOP_OK;
  };
  size_t prev_r0 = 0, prev_nr = 0;
  int slot = 0;
  for (size_t r0 = row_first; r0 < row_first + rows; r0 += chunk_rows, slot ^= 1) {
    const size_t nr = std::min(chunk_rows, row_first + rows - r0);
    RB_CUDA(cudaMemcpy2DAsync(b->h_stage[slot], row_bytes, static_cast<const char*>(b->d_out) + r0 * b->ldn * sb, b->ldn * sb, row_bytes, nr,
                              cudaMemcpyDeviceToHost, stream));
    RB_CUDA(cudaEventRecord(b->ev_stage[slot], stream));
    if (prev_nr) {
      int st = drain(slot ^ 1, prev_r0, prev_nr);
      if (st) return st;
    }
    prev_r0 = r0;
    prev_nr = nr;
  }
  return prev_nr ? drain(slot ^ 1, prev_r0, prev_nr) : REBOP_OK;
}

static int run_grid_impl(rebop_batch* b, double tmax, uint32_t nb_steps, const uint32_t* save_idx, uint32_t n_save, void* host_out,
                         size_t host_ld) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (nb_steps == 0) return rb_fail(REBOP_ERR_INVALID, "run_grid needs nb_steps >= 1 (nb_steps = 0 is the event-log mode)");
  if (host_out && host_ld < b->n) return rb_fail(REBOP_ERR_INVALID, "ld must be at least the number of trajectories of the batch");
  RB_CUDA(cudaSetDevice(b->device));
  const uint32_t S = b->net.n_species;
  std::vector<uint32_t> all;
  if (!save_idx) {
    all.resize(S);
    for (uint32_t s = 0; s < S; ++s) all[s] = s;
    save_idx = all.data();
    n_save = S;
  }
  for (uint32_t j = 0; j < n_save; ++j) {
    if (save_idx[j] >= S) return rb_fail(REBOP_ERR_OUT_OF_RANGE, "save_idx refers to a species index out of range");
    if (j > 0 && save_idx[j] <= save_idx[j - 1])
      return rb_fail(REBOP_ERR_INVALID, "save_idx must be strictly increasing");
  }
  const size_t rows = (size_t)(nb_steps + 1) * n_save;
  if (rows > 0x7fffffffu) return rb_fail(REBOP_ERR_LIMIT, "more than 2^31 sample rows");
  // a repeat of the call the watchdog cut short continues it from the segment and grid points reached
  const bool resuming = b->pend.active && b->pend.kind == 1 && b->pend.tmax == tmax && b->pend.nb_steps == nb_steps &&
                        b->pend.save == std::vector<uint32_t>(save_idx, save_idx + n_save);
  const unsigned seg_first = resuming ? b->pend.segment : 0u;
  b->pend.active = false;
  {
    int st = ensure_capacity(reinterpret_cast<char**>(&b->d_out), &b->out_capacity, std::max<size_t>(1, rows * b->ldn) * b->sample_bytes);
    if (!st) st = ensure_capacity(&b->d_sums, &b->sums_capacity, std::max<size_t>(2, 2 * rows));
    if (st) return st;
  }
  b->out_rows = (uint32_t)rows;
  b->out_n_save = n_save;
  b->out_nb_steps = nb_steps;
  if (!resuming) b->sums_ready = false;
  // The grid runs as a few segments of consecutive grid points when that pays: with a host buffer and a result
  // worth the trouble, the rows of a finished segment travel to the host while the next segment is being simulated;
  // and a very large result bounds the per-trajectory record buffer of the claiming schedules (a launch that ends
  // at grid point k leaves every trajectory exactly where a single launch would have it at that point, so the
  // result does not depend on the segmentation).
  const size_t bytes = rows * b->n * sizeof(int);
  unsigned segments = 1;
  if (host_out && n_save && bytes >= ((size_t)256 << 20) && nb_steps + 1 >= 8) segments = 4;
  while (n_save && bytes / segments > ((size_t)16 << 30) && segments * 2 <= nb_steps + 1) segments *= 2;
  const bool overlap = host_out && segments > 1 && !is_async(b);
  if (overlap && !b->copy_stream) RB_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
  uint64_t events = resuming ? b->pend.events : 0, lane_slots = resuming ? b->pend.lane_slots : 0;
  float ms = resuming ? b->pend.ms : 0.f, finish_ms = 0.f;
  int st = REBOP_OK;
  unsigned sgm = seg_first;
  // The rows of a finished segment go to the host AFTER the next segment has been launched, so that the copy --
  // asynchronous into page-locked memory, staged and blocking into pageable memory -- overlaps that segment's kernel.
  bool have_prev = false;
  uint32_t prev_first = 0, prev_last = 0;
  auto copy_prev = [&]() {
    have_prev = false;
    return copy_rows_to_host(b, host_out, host_ld, (size_t)prev_first * n_save, (size_t)(prev_last - prev_first + 1) * n_save,
                             overlap ? b->copy_stream : b->stream, true);
  };
  for (; sgm < segments && st == REBOP_OK; ++sgm) {
    const uint32_t first = (uint32_t)((uint64_t)(nb_steps + 1) * sgm / segments);
    const uint32_t last = (uint32_t)((uint64_t)(nb_steps + 1) * (sgm + 1) / segments) - 1;
    st = begin_call(b);
    if (!st) st = launch(b, tmax, nb_steps, first, last, n_save != 0, n_save, save_idx, resuming && sgm == seg_first, nb_steps + 1);
    if (st) break;
    if (host_out && n_save && is_async(b))
      st = copy_rows_to_host(b, host_out, host_ld, (size_t)first * n_save, (size_t)(last - first + 1) * n_save, b->stream, false);
    if (!st && have_prev) st = copy_prev();
    if (st) break;
    st = end_call(b, "a trajectory hit the per-launch iteration cap before reaching its last grid point (call again to continue)");
    if (!is_async(b)) {
      events += b->events_last;
      lane_slots += b->lane_slots_last;
      read_times(b, n_save != 0, &ms, &finish_ms);
    }
    if (st) break;  // (before ++sgm: a segment cut short by the watchdog is the one the repeated call continues)
    if (host_out && n_save && !is_async(b)) {
      have_prev = true;
      prev_first = first;
      prev_last = last;
    }
  }
  if (have_prev && st == REBOP_OK) st = copy_prev();
  if (!is_async(b)) {
    cudaError_t err = cudaStreamSynchronize(overlap ? b->copy_stream : b->stream);
    if (st == REBOP_OK && err != cudaSuccess) st = rb_fail(REBOP_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(err));
    b->events_last = events;
    b->lane_slots_last = lane_slots;
    b->last_ms = ms;
    b->last_finish_ms = finish_ms;
  }
  if (st == REBOP_ERR_ITER_CAP) {
    b->pend.active = true;
    b->pend.kind = 1;
    b->pend.tmax = tmax;
    b->pend.nb_steps = nb_steps;
    b->pend.n_save = n_save;
    b->pend.save.assign(save_idx, save_idx + n_save);
    b->pend.segment = sgm;
    b->pend.events = events;
    b->pend.lane_slots = lane_slots;
    b->pend.ms = ms;
  }
  return st;
}

extern "C" int rebop_batch_run_grid(rebop_batch* b, double tmax, uint32_t nb_steps, const uint32_t* save_idx,
                                    uint32_t n_save, int32_t* host_out) {
  if (b && host_out && b->sample_bytes != 4)
    return rb_fail(REBOP_ERR_INVALID, "rebop_batch_run_grid writes int32 samples: use rebop_batch_run_grid_typed with this batch's sample type");
  return run_grid_impl(b, tmax, nb_steps, save_idx, n_save, host_out, b ? b->n : 0);
}

extern "C" int rebop_batch_run_grid_typed(rebop_batch* b, double tmax, uint32_t nb_steps, const uint32_t* save_idx,
                                          uint32_t n_save, void* host_out) {
  return run_grid_impl(b, tmax, nb_steps, save_idx, n_save, host_out, b ? b->n : 0);
}

extern "C" int rebop_batch_run_grid_strided(rebop_batch* b, double tmax, uint32_t nb_steps, const uint32_t* save_idx,
                                            uint32_t n_save, void* host_out, size_t ld) {
  return run_grid_impl(b, tmax, nb_steps, save_idx, n_save, host_out, ld);
}

extern "C" int rebop_batch_set_sample_dtype(rebop_batch* b, int sample_bytes) {
  if (!b || (sample_bytes != 2 && sample_bytes != 4 && sample_bytes != 8))
    return rb_fail(REBOP_ERR_INVALID, "sample type must be REBOP_SAMPLES_I16, _I32 or _I64");
  b->sample_bytes = sample_bytes;
  b->out_rows = 0;  // samples of an earlier run are in the old type
  b->pend.active = false;
  return REBOP_OK;
}
extern "C" int rebop_batch_get_sample_dtype(const rebop_batch* b, int* sample_bytes) {
  if (!b || !sample_bytes) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *sample_bytes = b->sample_bytes;
  return REBOP_OK;
}

// Launch set-up shared by the event-log entry points (kernels <name>_evc / <name>_evw).
struct EventsLaunch {
  SsaRunParams p;
  RbJitKernel jit;
  bool use_jit = false;
  int jit_kind = REBOP_KERNEL_NVRTC;
  unsigned block = RB_TABLE_BLOCK, grid = 1;
  size_t smem = 0;
};

static int events_setup(rebop_batch* b, const uint32_t* save_idx, uint32_t n_save, EventsLaunch* e) {
  const uint32_t S = b->net.n_species;
  if (b->kernel_pref == REBOP_KERNEL_PDM)
    return rb_fail(REBOP_ERR_LIMIT, "the partial-propensity kernel (REBOP_KERNEL_PDM) has no event-log / single-step entry points: "
                                     "use one of the bit-exact kernels");
  int st = pick_kernel(b, true, &e->jit, &e->use_jit, &e->jit_kind);
  if (!st) st = apply_seeding(b);
  if (st) return st;
  SsaRunParams& p = e->p;
  fill_params(b, &p);
  p.n_save = n_save;
  if (e->use_jit) {
    for (uint32_t j = 0; j < n_save && save_idx[j] < 128; ++j) p.save_mask[save_idx[j] >> 6] |= 1ull << (save_idx[j] & 63u);
    e->block = e->jit.block;
    e->smem = RB_SSA_SMEM_BYTES(e->jit.net_words, e->block, 0, 0);
  } else {
    if (n_save > RB_TAB_MAX_SAVE) return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: more than 1024 saved species");
    e->smem = RB_SSA_SMEM_BYTES(S * e->block, e->block, 0, 0);
    if (e->smem + RB_STATIC_SMEM_BYTES > (size_t)b->max_smem_optin)
      return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: species state does not fit in shared memory");
    st = upload_tables(b);
    if (st) return st;
    p.tables = b->d_tables;
  }
  if (!e->use_jit || e->jit.large) {
    st = upload_gtab(b, save_idx, n_save);
    if (st) return st;
    p.gtab = b->d_gtab;
  }
  e->grid = (unsigned)((b->n + e->block - 1) / e->block);
  return REBOP_OK;
}

static int events_pass(rebop_batch* b, const EventsLaunch& e, bool write) {
  if (e.use_jit) {
    int rc = rb_jit_launch_entry(write ? e.jit.kernel_evw : e.jit.kernel_evc, e.block, e.p, e.grid, e.smem, b->stream);
    if (rc) return rc;
  } else {
    RB_CUDA(rb_table_launch_events(write, e.p, e.grid, e.smem, b->stream));
  }
  ++g_kernel_launches;
  return REBOP_OK;
}

// Gillespie::advance_one_reaction (src/gillespie.rs:270-297) on every trajectory: exactly one pass of the direct
// method whatever the time -- propensities; an absorbing state sets t = +inf; otherwise t += Exp1 / total, uniform,
// choice, update.  One launch of the event-log kernel limited to a single row, nothing logged.
extern "C" int rebop_batch_advance_one_reaction(rebop_batch* b) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  b->pend.active = false;
  EventsLaunch e;
  int st = events_setup(b, nullptr, 0, &e);
  if (st) return st;
  e.p.ev_single = 1;
  st = begin_call(b);
  if (st) return st;
  RB_CUDA(cudaEventRecord(b->ev0, b->stream));
  st = events_pass(b, e, true);
  if (st) return st;
  RB_CUDA(cudaEventRecord(b->ev1, b->stream));
  b->kernel_used = e.use_jit ? e.jit_kind : REBOP_KERNEL_TABLE;
  st = end_call(b, "iteration cap");
  if (!is_async(b)) {
    b->last_ms = b->last_finish_ms = 0.f;
    read_times(b, false, &b->last_ms, &b->last_finish_ms);
    b->lane_slots_last = 0;
  }
  return st;
}

// The nb_steps = 0 path of the binding (src/pyo3_gillespie.rs:209-223) for every trajectory: counting pass,
// host prefix sum, writing pass (see rb_ssa_events).
extern "C" int rebop_batch_run_events(rebop_batch* b, double tmax, const uint32_t* save_idx, uint32_t n_save) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  const uint32_t S = b->net.n_species;
  std::vector<uint32_t> all;
  if (!save_idx) {
    all.resize(S);
    for (uint32_t s = 0; s < S; ++s) all[s] = s;
    save_idx = all.data();
    n_save = S;
  }
  for (uint32_t j = 0; j < n_save; ++j) {
    if (save_idx[j] >= S) return rb_fail(REBOP_ERR_OUT_OF_RANGE, "save_idx refers to a species index out of range");
    if (j > 0 && save_idx[j] <= save_idx[j - 1]) return rb_fail(REBOP_ERR_INVALID, "save_idx must be strictly increasing");
  }
  b->pend.active = false;
  int st = sync_stream(b);  // this entry point blocks (the log's size is data): earlier asynchronous work is accounted first
  if (st) return st;
  EventsLaunch e;
  st = events_setup(b, save_idx, n_save, &e);
  if (st) return st;
  SsaRunParams& p = e.p;
  p.tmax = tmax;
  if (!b->d_ev_counts) RB_CUDA(cudaMalloc(&b->d_ev_counts, b->ldn * sizeof(rb_u32)));
  if (!b->d_ev_offsets) RB_CUDA(cudaMalloc(&b->d_ev_offsets, b->ldn * sizeof(rb_u64)));

  // pass 1: rows per trajectory
  p.ev_counts = b->d_ev_counts;
  RB_CUDA(cudaMemsetAsync(b->d_counters, 0, 4 * sizeof(rb_u64), b->stream));
  RB_CUDA(cudaEventRecord(b->ev0, b->stream));
  st = events_pass(b, e, false);
  if (st) return st;
  std::vector<rb_u32> counts(b->n);
  rb_u64 counters[4] = {0, 0, 0, 0};
  RB_CUDA(cudaMemcpyAsync(counts.data(), b->d_ev_counts, b->n * sizeof(rb_u32), cudaMemcpyDeviceToHost, b->stream));
  RB_CUDA(cudaMemcpyAsync(counters, b->d_counters, sizeof counters, cudaMemcpyDeviceToHost, b->stream));
  RB_CUDA(cudaStreamSynchronize(b->stream));
  if (counters[1] & RB_STATUS_ITER_CAP)
    return rb_fail(REBOP_ERR_ITER_CAP, "a trajectory hit the per-launch iteration cap before reaching tmax (nothing was changed)");
  b->ev_offsets.assign(b->n + 1, 0);
  for (size_t i = 0; i < b->n; ++i) b->ev_offsets[i + 1] = b->ev_offsets[i] + counts[i];
  const uint64_t total = b->ev_offsets[b->n];
  if (total >= 0xffffffffull) return rb_fail(REBOP_ERR_LIMIT, "event log of more than 2^32 rows: split the ensemble");
  size_t free_b = 0, total_b = 0;
  RB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t need_out = std::max<size_t>(1, (size_t)total * n_save);
  const size_t extra = (need_out > b->ev_out_capacity ? need_out * sizeof(int) : 0) +
                       (total > b->ev_times_capacity ? (size_t)total * sizeof(double) : 0);
  if (extra > free_b + b->ev_out_capacity * sizeof(int) + b->ev_times_capacity * sizeof(double))
    return rb_fail(REBOP_ERR_LIMIT, "event log does not fit in device memory: split the ensemble or save fewer species");
  st = ensure_capacity(&b->d_ev_out, &b->ev_out_capacity, need_out);
  if (!st) st = ensure_capacity(&b->d_ev_times, &b->ev_times_capacity, std::max<size_t>(1, total));
  if (st) return st;
  RB_CUDA(cudaMemcpyAsync(b->d_ev_offsets, b->ev_offsets.data(), b->n * sizeof(rb_u64), cudaMemcpyHostToDevice, b->stream));

  // pass 2: the same trajectories again, rows written at their offsets, final state written back
  p.ev_offsets = b->d_ev_offsets;
  p.ev_times = b->d_ev_times;
  p.ev_total = total;
  p.out = n_save ? b->d_ev_out : nullptr;
  RB_CUDA(cudaMemsetAsync(b->d_counters, 0, 4 * sizeof(rb_u64), b->stream));
  st = events_pass(b, e, true);
  if (st) return st;
  RB_CUDA(cudaEventRecord(b->ev1, b->stream));
  RB_CUDA(cudaMemcpyAsync(counters, b->d_counters, sizeof counters, cudaMemcpyDeviceToHost, b->stream));
  RB_CUDA(cudaStreamSynchronize(b->stream));
  RB_CUDA(cudaEventElapsedTime(&b->last_ms, b->ev0, b->ev1));
  b->last_finish_ms = 0.f;
  b->kernel_used = e.use_jit ? e.jit_kind : REBOP_KERNEL_TABLE;
  b->mode_last = RB_MODE_STATIC;
  b->events_last = counters[0];
  b->events_total += counters[0];
  b->lane_slots_last = 0;
  b->out_rows = 0;  // the grid accessors do not apply to an event log
  b->ev_n_save = n_save;
  return REBOP_OK;
}

extern "C" int rebop_batch_events_log_size(const rebop_batch* b, uint64_t* total_rows, uint32_t* n_save) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (b->ev_offsets.empty()) return rb_fail(REBOP_ERR_INVALID, "no event log: call rebop_batch_run_events first");
  if (total_rows) *total_rows = b->ev_offsets.back();
  if (n_save) *n_save = b->ev_n_save;
  return REBOP_OK;
}

extern "C" int rebop_batch_events_log_host(rebop_batch* b, uint64_t* offsets, double* times, int32_t* samples) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (b->ev_offsets.empty()) return rb_fail(REBOP_ERR_INVALID, "no event log: call rebop_batch_run_events first");
  RB_CUDA(cudaSetDevice(b->device));
  const uint64_t total = b->ev_offsets.back();
  if (offsets) std::memcpy(offsets, b->ev_offsets.data(), b->ev_offsets.size() * sizeof(uint64_t));
  if (times && total) RB_CUDA(cudaMemcpyAsync(times, b->d_ev_times, total * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (samples && total && b->ev_n_save)
    RB_CUDA(cudaMemcpyAsync(samples, b->d_ev_out, (size_t)total * b->ev_n_save * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  RB_CUDA(cudaStreamSynchronize(b->stream));
  return REBOP_OK;
}

extern "C" int rebop_batch_samples_device(const rebop_batch* b, const void** dev_ptr, size_t* ld, uint32_t* n_rows) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (dev_ptr) *dev_ptr = b->d_out;
  if (ld) *ld = b->ldn;
  if (n_rows) *n_rows = b->out_rows;
  return REBOP_OK;
}

// Samples in the batch's own sample type: row r goes to out[r * ld .. r * ld + n_traj).
static int samples_host_native(rebop_batch* b, void* out, size_t ld) {
  if (!b || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (ld < b->n) return rb_fail(REBOP_ERR_INVALID, "ld must be at least the number of trajectories of the batch");
  if (b->out_rows == 0) return REBOP_OK;
  RB_CUDA(cudaSetDevice(b->device));
  int st = copy_rows_to_host(b, out, ld, 0, b->out_rows, b->stream, true);
  if (st) return st;
  return sync_stream(b);
}

template <typename In, typename Out>
static int retype_chunks(rebop_batch* b, Out* out) {
  // a chunk of rows at a time through a device staging buffer of at most 256 MB
  const size_t rows_per_chunk = std::max<size_t>(1, ((size_t)256 << 20) / (b->n * sizeof(Out)));
  Out* d_stage = nullptr;
  const size_t chunk_rows = std::min<size_t>(rows_per_chunk, b->out_rows);
  RB_CUDA(cudaMalloc(&d_stage, chunk_rows * b->n * sizeof(Out)));
  cudaError_t err = cudaSuccess;
  for (size_t r0 = 0; r0 < b->out_rows && err == cudaSuccess; r0 += chunk_rows) {
    const unsigned rows = (unsigned)std::min<size_t>(chunk_rows, b->out_rows - r0);
    const unsigned blocks = (unsigned)std::min<size_t>(((size_t)rows * b->n + 255) / 256, (size_t)b->sm_count * 16);
    rb_rows_retype_kernel<In, Out><<<blocks, 256, 0, b->stream>>>(static_cast<const In*>(b->d_out) + r0 * b->ldn, d_stage, (unsigned)b->n,
                                                                 (unsigned)b->ldn, rows);
    ++g_kernel_launches;
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaMemcpyAsync(out + r0 * b->n, d_stage, (size_t)rows * b->n * sizeof(Out), cudaMemcpyDeviceToHost, b->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(b->stream);  // the staging buffer is reused by the next chunk
  }
  cudaFree(d_stage);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("sample conversion: ") + cudaGetErrorString(err));
  return account_async(b);
}

extern "C" int rebop_batch_samples_host(rebop_batch* b, void* out) { return samples_host_native(b, out, b ? b->n : 0); }

extern "C" int rebop_batch_samples_host_i32(rebop_batch* b, int32_t* out) {
  if (!b || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (b->out_rows == 0) return REBOP_OK;
  if (b->sample_bytes == 4) return samples_host_native(b, out, b->n);
  RB_CUDA(cudaSetDevice(b->device));
  if (b->sample_bytes == 2) return retype_chunks<short, int32_t>(b, out);
  return retype_chunks<long long, int32_t>(b, out);
}

extern "C" int rebop_batch_samples_host_i32_strided(rebop_batch* b, int32_t* out, size_t ld) {
  if (b && b->sample_bytes != 4) return rb_fail(REBOP_ERR_INVALID, "the batch's sample type is not int32: use rebop_batch_samples_host_strided");
  return samples_host_native(b, out, ld);
}

extern "C" int rebop_batch_samples_host_strided(rebop_batch* b, void* out, size_t ld) { return samples_host_native(b, out, ld); }

extern "C" int rebop_batch_samples_host_i64(rebop_batch* b, int64_t* out) {
  if (!b || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (b->out_rows == 0) return REBOP_OK;
  if (b->sample_bytes == 8) return samples_host_native(b, out, b->n);
  RB_CUDA(cudaSetDevice(b->device));
  if (b->sample_bytes == 2) return retype_chunks<short, long long>(b, reinterpret_cast<long long*>(out));
  return retype_chunks<int, long long>(b, reinterpret_cast<long long*>(out));
}

extern "C" int rebop_batch_sample_sums_device(rebop_batch* b, const int64_t** dev_ptr, uint32_t* n_rows) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  RB_CUDA(cudaSetDevice(b->device));
  const uint32_t rows = b->out_rows;
  if (rows == 0) return rb_fail(REBOP_ERR_INVALID, "no samples: call run_grid first");
  if (!b->sums_ready) {  // static schedule: the claiming schedules leave the sums behind with the samples
    RB_CUDA(cudaMemsetAsync(b->d_sums, 0, 2 * (size_t)rows * sizeof(rb_i64), b->stream));
    // enough CTAs per row to fill the machine, each thread moving 16 bytes per load
    const unsigned per_row = (unsigned)std::max<size_t>(1, std::min<size_t>((b->n / 4 + 255) / 256,
                                                          std::max<size_t>(1, (size_t)b->sm_count * 8 / rows + 1)));
    dim3 grid(rows, per_row);
    if (b->sample_bytes == 2)
      rb_row_sums_kernel<short><<<grid, 256, 0, b->stream>>>(static_cast<const short*>(b->d_out), (unsigned)b->n, (unsigned)b->ldn, b->d_sums, rows);
    else if (b->sample_bytes == 4)
      rb_row_sums_kernel<int><<<grid, 256, 0, b->stream>>>(static_cast<const int*>(b->d_out), (unsigned)b->n, (unsigned)b->ldn, b->d_sums, rows);
    else
      rb_row_sums_kernel<long long><<<grid, 256, 0, b->stream>>>(static_cast<const long long*>(b->d_out), (unsigned)b->n, (unsigned)b->ldn, b->d_sums, rows);
    RB_CUDA(cudaGetLastError());
    ++g_kernel_launches;
    b->sums_ready = true;
  }
  if (dev_ptr) *dev_ptr = reinterpret_cast<const int64_t*>(b->d_sums);
  if (n_rows) *n_rows = rows;
  return REBOP_OK;
}

extern "C" int rebop_batch_sample_sums(rebop_batch* b, int64_t* sum, uint64_t* sumsq) {
  if (!b || !sum || !sumsq) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  const int64_t* d = nullptr;
  uint32_t rows = 0;
  int st = rebop_batch_sample_sums_device(b, &d, &rows);
  if (st) return st;
  RB_CUDA(cudaMemcpyAsync(sum, d, rows * sizeof(int64_t), cudaMemcpyDeviceToHost, b->stream));
  RB_CUDA(cudaMemcpyAsync(sumsq, d + rows, rows * sizeof(int64_t), cudaMemcpyDeviceToHost, b->stream));
  return sync_stream(b);
}

extern "C" int rebop_batch_events(rebop_batch* b, uint64_t* total, uint64_t* last_launch) {
  if (!b) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (total) *total = b->events_total;
  if (last_launch) *last_launch = b->events_last;
  return REBOP_OK;
}

extern "C" int rebop_batch_lane_slots(rebop_batch* b, uint64_t* last_launch) {
  if (!b || !last_launch) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *last_launch = b->lane_slots_last;
  return REBOP_OK;
}

extern "C" int rebop_batch_last_kernel_ms(rebop_batch* b, float* ms) {
  if (!b || !ms) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *ms = b->last_ms;
  return REBOP_OK;
}

extern "C" int rebop_batch_last_finish_ms(rebop_batch* b, float* ms) {
  if (!b || !ms) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *ms = b->last_finish_ms;
  return REBOP_OK;
}

extern "C" int rebop_b200_measure_fp64_rate(int device, double* ops_per_second, double* sm_clock_mhz) {
  int st = select_device(device);
  if (st) return st;
  int sms = 0;
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* sink = nullptr;
  long long* clocks = nullptr;
  RB_CUDA(cudaMalloc(&sink, (size_t)blocks * threads * sizeof(double)));
  RB_CUDA(cudaMalloc(&clocks, sizeof(long long)));
  cudaEvent_t e0, e1;
  RB_CUDA(cudaEventCreate(&e0));
  RB_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  long long clk = 0;
  for (int rep = 0; rep < 5; ++rep) {  // first pass warms up
    RB_CUDA(cudaEventRecord(e0));
    rb_fp64_probe_kernel<<<blocks, threads>>>(sink, iters, clocks);
    ++g_kernel_launches;
    RB_CUDA(cudaEventRecord(e1));
    RB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) {
      best = ms;
      RB_CUDA(cudaMemcpy(&clk, clocks, sizeof clk, cudaMemcpyDeviceToHost));
    }
  }
  RB_CUDA(cudaGetLastError());
  const double ops = (double)blocks * threads * (double)iters * 32.0;
  if (ops_per_second) *ops_per_second = ops / (best * 1e-3);
  (void)clk;
  if (sm_clock_mhz) {
    unsigned long long* d_probe = nullptr;
    RB_CUDA(cudaMalloc(&d_probe, (size_t)sms * 2 * sizeof(unsigned long long)));
    rb_clock_probe_kernel<<<sms, 32>>>(1 << 20, d_probe);
    ++g_kernel_launches;
    std::vector<unsigned long long> h((size_t)sms * 2);
    RB_CUDA(cudaMemcpy(h.data(), d_probe, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(d_probe);
    std::vector<double> mhz;
    for (int i = 0; i < sms; ++i)
      if (h[2 * i + 1]) mhz.push_back((double)h[2 * i] / (double)h[2 * i + 1] * 1e3);
    std::sort(mhz.begin(), mhz.end());
    *sm_clock_mhz = mhz.empty() ? 0.0 : mhz[mhz.size() / 2];  // median over the SMs
  }
  cudaFree(sink);
  cudaFree(clocks);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return REBOP_OK;
}
