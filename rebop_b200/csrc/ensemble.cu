// ensemble.cu -- one ensemble over several GPUs of one node, inside the C ABI.
//
// Trajectories are independent units (the reference has no exchange step at all: its only "ensemble" is the host
// loop of examples/sir.rs:11-22 that averages repeated runs), so an ensemble of N trajectories shards into the
// contiguous ranges [g*N/G, (g+1)*N/G), one batch per device, every trajectory with its own seed: results do not
// depend on the number of devices.  One process, one host worker thread per device for the duration of a call.
// There is no data-path collective.  The only exchange is the optional ensemble statistics: every device reduces
// its shard's samples to exact int64 row sums and sums of squares (K4, fused into the sample-finishing kernel), the
// sums are all-reduced with ncclAllReduce(ncclInt64, ncclSum) over an ncclCommInitAll communicator, and a
// finalisation kernel turns them into mean and variance with 128-bit integer arithmetic (no cancellation however
// large the ensemble).  NCCL is opened with dlopen, like NVRTC: the library loads without it, and one device needs
// none.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "network.hpp"
#include "ssa_params.h"

#define RB_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t err__ = (call);                                                                \
    if (err__ != cudaSuccess)                                                                  \
      return rb_fail(REBOP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));   \
  } while (0)

namespace {

typedef struct ncclComm* ncclComm_t;
struct Nccl {
  void* handle = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};
void nccl_load(Nccl& n);
constexpr int kNcclInt64 = 4, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable across NCCL 2.x)

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] { nccl_load(n); });
  return n;
}

void nccl_load(Nccl& n) {
  std::vector<std::string> names;
  if (const char* env = std::getenv("REBOP_B200_NCCL_LIB")) names.push_back(env);
  names.insert(names.end(), {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"});
  for (const std::string& name : names) {
    n.handle = dlopen(name.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (n.handle) break;
  }
  if (!n.handle) {
    n.error = "libnccl.so.2 not found (set REBOP_B200_NCCL_LIB to its path)";
    return;
  }
#define RB_SYM(field, sym)                                             \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.handle, sym)); \
  if (!n.field) n.error = std::string("missing NCCL symbol ") + sym;
  RB_SYM(CommInitAll, "ncclCommInitAll")
  RB_SYM(CommDestroy, "ncclCommDestroy")
  RB_SYM(AllReduce, "ncclAllReduce")
  RB_SYM(GroupStart, "ncclGroupStart")
  RB_SYM(GroupEnd, "ncclGroupEnd")
  RB_SYM(GetErrorString, "ncclGetErrorString")
#undef RB_SYM
}

// mean = sum / n, unbiased variance = (n * sumsq - sum^2) / (n (n - 1)), the numerator in 128-bit integers.
__global__ void rb_stats_finalize_kernel(const rb_i64* __restrict__ sums, unsigned rows, rb_u64 n, double* __restrict__ mean,
                                         double* __restrict__ var) {
  const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const rb_i64 a = sums[r];
  const rb_u64 q = (rb_u64)sums[rows + r];
  mean[r] = (double)a / (double)n;
  if (n < 2) {
    var[r] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  const unsigned __int128 lhs = (unsigned __int128)n * q;
  const unsigned __int128 rhs = (unsigned __int128)((__int128)a * a);
  const unsigned __int128 num = lhs >= rhs ? lhs - rhs : 0;  // Cauchy-Schwarz: never negative
  const double numd = (double)(rb_u64)(num >> 64) * 0x1.0p64 + (double)(rb_u64)num;
  var[r] = numd / ((double)n * (double)(n - 1));
}

}  // namespace

struct rebop_ensemble {
  std::vector<int> devices;
  std::vector<rebop_batch*> shard;
  std::vector<size_t> first, count;
  size_t n_total = 0;
  std::vector<ncclComm_t> comms;  // created on the first statistics call over more than one device
  std::vector<int64_t*> d_reduced;  // per device: the ensemble's sums (the shards' own sums stay what they are, so that
  size_t reduced_capacity = 0;      // asking twice -- sums, then stats -- reduces the same inputs twice, not the result)
  uint32_t rows = 0;
  uint32_t n_species = 0;
  int sample_bytes = 4;
};

// Runs fn(g) for every shard on a worker thread of its own; the first failure (status and message) is reported.
template <class F>
static int for_each_shard(rebop_ensemble* e, F fn) {
  const size_t G = e->shard.size();
  std::vector<int> status(G, REBOP_OK);
  std::vector<std::string> message(G);
  auto work = [&](size_t g) {
    if (!e->shard[g]) return;
    status[g] = fn(g);
    if (status[g]) message[g] = rebop_b200_last_error();  // the message is thread-local: carry it over
  };
  if (G == 1) {
    work(0);
  } else {
    std::vector<std::thread> threads;
    for (size_t g = 0; g < G; ++g) threads.emplace_back(work, g);
    for (std::thread& t : threads) t.join();
  }
  for (size_t g = 0; g < G; ++g)
    if (status[g]) return rb_fail(status[g], "device " + std::to_string(e->devices[g]) + ": " + message[g]);
  return REBOP_OK;
}

extern "C" void rebop_ensemble_destroy(rebop_ensemble* e) {
  if (!e) return;
  for (ncclComm_t c : e->comms)
    if (c) nccl().CommDestroy(c);
  for (size_t g = 0; g < e->d_reduced.size(); ++g) {
    if (!e->d_reduced[g]) continue;
    cudaSetDevice(e->devices[g]);
    cudaFree(e->d_reduced[g]);
  }
  for (rebop_batch* b : e->shard) rebop_batch_destroy(b);
  delete e;
}

extern "C" int rebop_ensemble_create(const rebop_network* net, const int* devices, int n_devices, size_t n_traj, const int64_t* x0,
                                     int x0_per_trajectory, const uint64_t* seeds, uint64_t seed_base, rebop_ensemble** out) {
  if (!net || !out || !devices) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (n_devices < 1) return rb_fail(REBOP_ERR_INVALID, "an ensemble needs at least one device");
  if (n_traj == 0) return rb_fail(REBOP_ERR_INVALID, "n_traj must be at least 1");
  uint32_t S = 0;
  rebop_network_nb_species(net, &S);
  rebop_ensemble* e = new rebop_ensemble();
  e->devices.assign(devices, devices + n_devices);
  e->n_total = n_traj;
  e->n_species = S;
  e->shard.assign(n_devices, nullptr);
  e->first.resize(n_devices);
  e->count.resize(n_devices);
  for (int g = 0; g < n_devices; ++g) {
    e->first[g] = n_traj * (size_t)g / (size_t)n_devices;
    e->count[g] = n_traj * (size_t)(g + 1) / (size_t)n_devices - e->first[g];
  }
  // creation is serial (it is cheap and keeps the error path simple); a device with an empty range gets no batch
  for (int g = 0; g < n_devices; ++g) {
    if (e->count[g] == 0) continue;
    const int64_t* x0g = (x0 && x0_per_trajectory) ? x0 + e->first[g] * S : x0;
    int st = rebop_batch_create(net, devices[g], e->count[g], x0g, x0_per_trajectory, seeds ? seeds + e->first[g] : nullptr,
                                seed_base + e->first[g], &e->shard[g]);
    if (st) {
      const std::string msg = rebop_b200_last_error();
      rebop_ensemble_destroy(e);
      return rb_fail(st, msg);
    }
  }
  *out = e;
  return REBOP_OK;
}

extern "C" int rebop_ensemble_shards(const rebop_ensemble* e, int* n_shards) {
  if (!e || !n_shards) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *n_shards = (int)e->shard.size();
  return REBOP_OK;
}

extern "C" int rebop_ensemble_shard(const rebop_ensemble* e, int g, rebop_batch** batch, int* device, size_t* first, size_t* count) {
  if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  if (g < 0 || (size_t)g >= e->shard.size()) return rb_fail(REBOP_ERR_OUT_OF_RANGE, "shard index out of range");
  if (batch) *batch = e->shard[g];
  if (device) *device = e->devices[g];
  if (first) *first = e->first[g];
  if (count) *count = e->count[g];
  return REBOP_OK;
}

#define RB_ENSEMBLE_FORWARD(name, decl, call)                                            \
  extern "C" int rebop_ensemble_##name decl {                                             \
    if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");                           \
    return for_each_shard(e, [&](size_t g) { rebop_batch* b = e->shard[g]; return call; }); \
  }
RB_ENSEMBLE_FORWARD(set_kernel, (rebop_ensemble * e, int kind), rebop_batch_set_kernel(b, kind))
RB_ENSEMBLE_FORWARD(set_schedule, (rebop_ensemble * e, int schedule), rebop_batch_set_schedule(b, schedule))
RB_ENSEMBLE_FORWARD(set_max_iters, (rebop_ensemble * e, uint32_t max_iters), rebop_batch_set_max_iters(b, max_iters))
RB_ENSEMBLE_FORWARD(set_rates, (rebop_ensemble * e, const double* k, size_t n_reactions), rebop_batch_set_rates(b, k, n_reactions))
RB_ENSEMBLE_FORWARD(set_time, (rebop_ensemble * e, double t), rebop_batch_set_time(b, t))
RB_ENSEMBLE_FORWARD(advance_until, (rebop_ensemble * e, double tmax), rebop_batch_advance_until(b, tmax))
RB_ENSEMBLE_FORWARD(advance_one_reaction, (rebop_ensemble * e), rebop_batch_advance_one_reaction(b))
#undef RB_ENSEMBLE_FORWARD

extern "C" int rebop_ensemble_set_sample_dtype(rebop_ensemble* e, int dtype) {
  if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  int st = for_each_shard(e, [&](size_t g) { return rebop_batch_set_sample_dtype(e->shard[g], dtype); });
  if (!st) e->sample_bytes = dtype;
  return st;
}

extern "C" int rebop_ensemble_seed(rebop_ensemble* e, const uint64_t* seeds, uint64_t seed_base) {
  if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  return for_each_shard(e, [&](size_t g) {
    return rebop_batch_seed(e->shard[g], seeds ? seeds + e->first[g] : nullptr, seed_base + e->first[g]);
  });
}

extern "C" int rebop_ensemble_set_species(rebop_ensemble* e, const int64_t* species, int per_trajectory) {
  if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  return for_each_shard(e, [&](size_t g) {
    const int64_t* p = (species && per_trajectory) ? species + e->first[g] * e->n_species : species;
    return rebop_batch_set_species(e->shard[g], p, per_trajectory);
  });
}

extern "C" int rebop_ensemble_run_grid(rebop_ensemble* e, double tmax, uint32_t nb_steps, const uint32_t* save_idx, uint32_t n_save,
                                       void* host_out) {
  if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  int st = for_each_shard(e, [&](size_t g) {
    // every shard writes its own columns of the [step][save][all trajectories] host array, rows n_total apart
    void* dst = host_out ? static_cast<char*>(host_out) + e->first[g] * (size_t)e->sample_bytes : nullptr;
    return rebop_batch_run_grid_strided(e->shard[g], tmax, nb_steps, save_idx, n_save, dst, e->n_total);
  });
  if (st) return st;
  uint32_t rows = 0;
  for (rebop_batch* b : e->shard)
    if (b) {
      rebop_batch_samples_device(b, nullptr, nullptr, &rows);
      break;
    }
  e->rows = rows;
  return REBOP_OK;
}

extern "C" int rebop_ensemble_samples_host(rebop_ensemble* e, void* out) {
  if (!e || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  return for_each_shard(e, [&](size_t g) {
    return rebop_batch_samples_host_strided(e->shard[g], static_cast<char*>(out) + e->first[g] * (size_t)e->sample_bytes, e->n_total);
  });
}

static int first_live(const rebop_ensemble* e) {
  for (size_t g = 0; g < e->shard.size(); ++g)
    if (e->shard[g]) return (int)g;
  return -1;
}

// Row sums of every shard, all-reduced on the devices into buffers of the ensemble's own (every device ends up with
// the ensemble's sums; the shards' sums are left as they are).  *d_sums: where the ensemble's sums are, per device.
static int reduce_sums(rebop_ensemble* e, std::vector<const int64_t*>* d_sums) {
  const size_t G = e->shard.size();
  d_sums->assign(G, nullptr);
  if (e->rows == 0) return rb_fail(REBOP_ERR_INVALID, "no samples: call rebop_ensemble_run_grid first");
  int st = for_each_shard(e, [&](size_t g) {
    uint32_t rows = 0;
    return rebop_batch_sample_sums_device(e->shard[g], &(*d_sums)[g], &rows);
  });
  if (st) return st;
  size_t live = 0;
  bool distinct = true;
  for (size_t g = 0; g < G; ++g) {
    if (!e->shard[g]) continue;
    ++live;
    for (size_t h = 0; h < g; ++h) distinct = distinct && !(e->shard[h] && e->devices[h] == e->devices[g]);
  }
  if (live < 2) return REBOP_OK;
  // receive buffers of the reduction, one per device
  const size_t need = 2 * (size_t)e->rows;
  if (e->d_reduced.size() != G || e->reduced_capacity < need) {
    for (size_t g = 0; g < e->d_reduced.size(); ++g) {
      if (!e->d_reduced[g]) continue;
      cudaSetDevice(e->devices[g]);
      cudaFree(e->d_reduced[g]);
    }
    e->d_reduced.assign(G, nullptr);
    e->reduced_capacity = 0;
    for (size_t g = 0; g < G; ++g) {
      if (!e->shard[g]) continue;
      RB_CUDA(cudaSetDevice(e->devices[g]));
      RB_CUDA(cudaMalloc(&e->d_reduced[g], need * sizeof(int64_t)));
    }
    e->reduced_capacity = need;
  }
  if (!distinct) {
    // several shards on one device (a way to exercise the sharding on a single GPU): NCCL wants one rank per
    // device, so the few thousand integers are added on the host and handed back to the first shard's buffer
    const size_t count = 2 * (size_t)e->rows;
    std::vector<int64_t> total(count, 0), part(count);
    for (size_t g = 0; g < G; ++g) {
      if (!e->shard[g]) continue;
      RB_CUDA(cudaSetDevice(e->devices[g]));
      void* stream = nullptr;
      rebop_batch_get_stream(e->shard[g], &stream);
      RB_CUDA(cudaMemcpyAsync(part.data(), (*d_sums)[g], count * sizeof(int64_t), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
      RB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
      for (size_t i = 0; i < count; ++i) total[i] = (int64_t)((uint64_t)total[i] + (uint64_t)part[i]);
    }
    const int g0 = first_live(e);
    RB_CUDA(cudaSetDevice(e->devices[g0]));
    RB_CUDA(cudaMemcpy(e->d_reduced[g0], total.data(), count * sizeof(int64_t), cudaMemcpyHostToDevice));
    (*d_sums)[g0] = e->d_reduced[g0];
    return REBOP_OK;
  }
  Nccl& n = nccl();
  if (!n.error.empty()) return rb_fail(REBOP_ERR_NCCL, "NCCL unavailable: " + n.error);
  if (e->comms.empty()) {
    std::vector<int> devs;
    for (size_t g = 0; g < G; ++g)
      if (e->shard[g]) devs.push_back(e->devices[g]);
    std::vector<ncclComm_t> comms(devs.size(), nullptr);
    int rc = n.CommInitAll(comms.data(), (int)devs.size(), devs.data());
    if (rc != 0) return rb_fail(REBOP_ERR_NCCL, std::string("ncclCommInitAll: ") + n.GetErrorString(rc));
    e->comms.assign(G, nullptr);
    size_t j = 0;
    for (size_t g = 0; g < G; ++g)
      if (e->shard[g]) e->comms[g] = comms[j++];
  }
  int rc = n.GroupStart();
  for (size_t g = 0; g < G && rc == 0; ++g) {
    if (!e->shard[g]) continue;
    void* stream = nullptr;
    rebop_batch_get_stream(e->shard[g], &stream);
    rc = n.AllReduce((*d_sums)[g], e->d_reduced[g], 2 * (size_t)e->rows, kNcclInt64, kNcclSum, e->comms[g], static_cast<cudaStream_t>(stream));
  }
  const int rc_end = n.GroupEnd();
  if (rc == 0) rc = rc_end;
  if (rc != 0) return rb_fail(REBOP_ERR_NCCL, std::string("ncclAllReduce: ") + n.GetErrorString(rc));
  for (size_t g = 0; g < G; ++g)
    if (e->shard[g]) (*d_sums)[g] = e->d_reduced[g];
  return REBOP_OK;
}

extern "C" int rebop_ensemble_sums(rebop_ensemble* e, int64_t* sum, uint64_t* sumsq) {
  if (!e || !sum || !sumsq) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::vector<const int64_t*> d_sums;
  int st = reduce_sums(e, &d_sums);
  if (st) return st;
  const int g = first_live(e);
  void* stream = nullptr;
  rebop_batch_get_stream(e->shard[g], &stream);
  RB_CUDA(cudaSetDevice(e->devices[g]));
  RB_CUDA(cudaMemcpyAsync(sum, d_sums[g], e->rows * sizeof(int64_t), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  RB_CUDA(cudaMemcpyAsync(sumsq, d_sums[g] + e->rows, e->rows * sizeof(int64_t), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  // the other devices must have finished their part of the collective before their buffers may be reused
  return for_each_shard(e, [&](size_t h) { return rebop_batch_synchronize(e->shard[h]); });
}

extern "C" int rebop_ensemble_stats(rebop_ensemble* e, double* mean, double* var) {
  if (!e || !mean || !var) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  std::vector<const int64_t*> d_sums;
  int st = reduce_sums(e, &d_sums);
  if (st) return st;
  const int g = first_live(e);
  void* sp = nullptr;
  rebop_batch_get_stream(e->shard[g], &sp);
  cudaStream_t stream = static_cast<cudaStream_t>(sp);
  RB_CUDA(cudaSetDevice(e->devices[g]));
  double* d_stats = nullptr;
  RB_CUDA(cudaMallocAsync(&d_stats, 2 * (size_t)e->rows * sizeof(double), stream));
  rb_stats_finalize_kernel<<<(e->rows + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const rb_i64*>(d_sums[g]), e->rows, (rb_u64)e->n_total, d_stats, d_stats + e->rows);
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemcpyAsync(mean, d_stats, e->rows * sizeof(double), cudaMemcpyDeviceToHost, stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(var, d_stats + e->rows, e->rows * sizeof(double), cudaMemcpyDeviceToHost, stream);
  cudaFreeAsync(d_stats, stream);
  if (err != cudaSuccess) return rb_fail(REBOP_ERR_CUDA, std::string("rebop_ensemble_stats: ") + cudaGetErrorString(err));
  return for_each_shard(e, [&](size_t h) { return rebop_batch_synchronize(e->shard[h]); });
}

extern "C" int rebop_ensemble_events(rebop_ensemble* e, uint64_t* total, uint64_t* last_call) {
  if (!e) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  uint64_t t = 0, l = 0;
  for (rebop_batch* b : e->shard) {
    if (!b) continue;
    uint64_t bt = 0, bl = 0;
    rebop_batch_events(b, &bt, &bl);
    t += bt;
    l += bl;
  }
  if (total) *total = t;
  if (last_call) *last_call = l;
  return REBOP_OK;
}

extern "C" int rebop_ensemble_last_kernel_ms(rebop_ensemble* e, float* max_ms) {
  if (!e || !max_ms) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *max_ms = 0.f;
  for (rebop_batch* b : e->shard) {
    if (!b) continue;
    float ms = 0.f;
    rebop_batch_last_kernel_ms(b, &ms);
    if (ms > *max_ms) *max_ms = ms;
  }
  return REBOP_OK;
}

extern "C" int rebop_ensemble_size(const rebop_ensemble* e, size_t* n_traj) {
  if (!e || !n_traj) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *n_traj = e->n_total;
  return REBOP_OK;
}
