// samples.cu -- K6: turns the per-trajectory sample records of the lanes-claim-trajectories schedules into the
// result layout of the C ABI, [step][saved species][trajectory], in the sample type the caller asked for.
//
// The ensemble loop appends the rows of a trajectory to a record of its own, raw[traj][step][row] (int32):
// lanes of a warp run unrelated trajectories at unrelated grid points, so that is the only layout in which
// their stores fill whole sectors.  This kernel is the second half of the pyo3 grid loop's
// `push species[save_idx]` (src/pyo3_gillespie.rs:205-207): a tiled transpose that
//   * repeats the last row a trajectory wrote for every later grid point (the trajectory reached an absorbing
//     state and stopped there, see RB_PROGRESS_* in ssa_params.h),
//   * converts int32 counts to int16 / int32 / int64 (the reference returns isize; int16 halves HBM and PCIe
//     traffic for models whose counts are small; a count that does not fit raises RB_STATUS_NARROW),
//   * writes [row][trajectory] lines from a shared-memory tile with TMA bulk stores
//     (cp.async.bulk.global.shared::cta, one 512-byte..1-KB line per row of the tile), and
//   * accumulates the exact integer sum and sum of squares of every row (K4 fused in: the samples are not read
//     a second time for the ensemble statistics).
// It is HBM-bound: 4 B read + sizeof(sample) written per sample.
#include <cuda_runtime.h>

#include <algorithm>

#include "samples.h"

namespace {

constexpr int kRows = 32;     // rows of a tile
constexpr int kTraj = 128;    // trajectories of a tile
constexpr int kThreads = 256;

template <typename OutT>
struct Tile {
  static constexpr int kStride = kTraj + 16 / (int)sizeof(OutT);  // +16 bytes: rows stay 16-byte aligned, fewer bank conflicts
  OutT v[kRows][kStride];
};

__device__ __forceinline__ void bulk_store(void* dst, const void* src_shared, unsigned bytes) {
  const unsigned src = (unsigned)__cvta_generic_to_shared(src_shared);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <typename OutT, bool BULK>
__global__ void __launch_bounds__(kThreads, 2048 / kThreads) rb_samples_finish_kernel(const RbFinishParams q) {
  __shared__ __align__(128) Tile<OutT> tile;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const unsigned row0 = blockIdx.x * kRows;
  const unsigned my_row = row0 + lane;           // load phase: lane = row of the tile
  const bool row_ok = my_row < q.rows;
  const unsigned my_step = row_ok ? my_row / q.n_save : 0u;
  const unsigned my_col = row_ok ? my_row - my_step * q.n_save : 0u;
  const unsigned n_tiles = (q.ldn + kTraj - 1) / kTraj;
  OutT* const out = static_cast<OutT*>(q.out);
  bool narrow = false;
  long long s1[kRows / 8] = {0, 0, 0, 0};        // sum phase: warp w owns rows w, w + 8, w + 16, w + 24 of the tile
  unsigned long long s2[kRows / 8] = {0, 0, 0, 0};

  for (unsigned tt = blockIdx.y; tt < n_tiles; tt += gridDim.y) {
    const unsigned traj0 = tt * kTraj;
    const unsigned width = min((unsigned)kTraj, q.ldn - traj0);  // multiple of 32 (ldn is)
    // ---- load: 128-byte reads along a trajectory's record, transposed into the tile
    for (unsigned t = warp; t < width; t += kThreads / 32) {
      const unsigned traj = traj0 + t;
      int v = 0;
      if (row_ok && traj < q.n) {
        const unsigned written = __ldg(q.progress + traj) & RB_PROGRESS_ROWS;  // >= 1: grid points the trajectory wrote itself
        const unsigned step = min(my_step, written - 1u);
        v = __ldcs(q.raw + ((size_t)traj * q.n_points + step) * q.n_save + my_col);
      }
      const OutT o = (OutT)v;
      if (sizeof(OutT) < 4 && (int)o != v) narrow = true;
      tile.v[lane][t] = o;
    }
    if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- sums of the tile's rows (padding trajectories hold 0)
    if (q.sums) {
#pragma unroll
      for (int i = 0; i < kRows / 8; ++i) {
        const unsigned r = warp + 8u * i;
        for (unsigned t = lane; t < width; t += 32u) {
          const long long x = (long long)tile.v[r][t];
          s1[i] += x;
          s2[i] += (unsigned long long)(x * x);
        }
      }
    }
    // ---- store: one line of `width` samples per row of the tile
    if (BULK) {
      if (tid < kRows && row0 + tid < q.rows) {
        bulk_store(out + (size_t)(row0 + tid) * q.ldn + traj0, &tile.v[tid][0], width * (unsigned)sizeof(OutT));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile may be overwritten once it has been read
      }
    } else {
      constexpr unsigned per16 = 16u / (unsigned)sizeof(OutT);
      const unsigned vec_per_row = width / per16;
      for (unsigned i = tid; i < kRows * vec_per_row; i += kThreads) {
        const unsigned r = i / vec_per_row, c = (i - r * vec_per_row) * per16;
        if (row0 + r < q.rows)
          __stcs(reinterpret_cast<int4*>(out + (size_t)(row0 + r) * q.ldn + traj0 + c), *reinterpret_cast<const int4*>(&tile.v[r][c]));
      }
    }
    __syncthreads();
  }

  if (q.sums) {
#pragma unroll
    for (int i = 0; i < kRows / 8; ++i) {
      long long a = s1[i];
      unsigned long long b = s2[i];
      for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, off);
        b += __shfl_down_sync(0xffffffffu, b, off);
      }
      const unsigned row = row0 + warp + 8u * i;
      if (lane == 0 && row < q.rows) {
        atomicAdd(reinterpret_cast<unsigned long long*>(q.sums) + row, (unsigned long long)a);
        atomicAdd(reinterpret_cast<unsigned long long*>(q.sums) + q.sums_stride + row, b);
      }
    }
  }
  if (narrow) atomicOr(q.status, RB_STATUS_NARROW);
}

// Static schedule: the samples are [row][ldn] int32 already; only the sample type changes.
template <typename OutT>
__global__ void __launch_bounds__(256) rb_rows_convert_kernel(const int* __restrict__ in, OutT* __restrict__ out, size_t count,
                                                              rb_u32* status) {
  bool narrow = false;
  const size_t n4 = count / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int4 v = __ldcs(reinterpret_cast<const int4*>(in) + i);
    const OutT a = (OutT)v.x, b = (OutT)v.y, c = (OutT)v.z, d = (OutT)v.w;
    if (sizeof(OutT) < 4 && ((int)a != v.x || (int)b != v.y || (int)c != v.z || (int)d != v.w)) narrow = true;
    out[4 * i] = a;
    out[4 * i + 1] = b;
    out[4 * i + 2] = c;
    out[4 * i + 3] = d;
  }
  if (narrow) atomicOr(status, RB_STATUS_NARROW);
}

template <typename OutT>
cudaError_t finish_typed(const RbFinishParams& q, bool bulk, unsigned sm_count, cudaStream_t stream) {
  const unsigned row_blocks = (q.rows + kRows - 1) / kRows;
  const unsigned n_tiles = (q.ldn + kTraj - 1) / kTraj;
  // One wave of resident CTAs (the kernel waits on its reads: a partial second wave runs at a fraction of the
  // bandwidth -- ncu, profiles/r2r_k6_finish_int16_ncu_full.md: 1200 CTAs on 888 slots, 1.35 waves); every CTA walks a
  // strided share of the trajectory tiles, so that the row sums cost one atomic per (row, CTA column) instead of one
  // per tile
  int resident = 0;
  if (bulk) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, rb_samples_finish_kernel<OutT, true>, kThreads, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, rb_samples_finish_kernel<OutT, false>, kThreads, 0);
  if (resident < 1) resident = 1;
  unsigned cols = std::max(1u, sm_count * (unsigned)resident / row_blocks);
  if (cols > n_tiles) cols = n_tiles;
  if (cols > 65535u) cols = 65535u;
  if (cols == 0) cols = 1;
  dim3 grid(row_blocks, cols);
  if (bulk) rb_samples_finish_kernel<OutT, true><<<grid, kThreads, 0, stream>>>(q);
  else rb_samples_finish_kernel<OutT, false><<<grid, kThreads, 0, stream>>>(q);
  return cudaGetLastError();
}

}  // namespace

cudaError_t rb_samples_finish(const RbFinishParams& q, int sample_bytes, bool bulk, unsigned sm_count, cudaStream_t stream) {
  if (q.rows == 0 || q.ldn == 0) return cudaSuccess;
  switch (sample_bytes) {
    case 2: return finish_typed<short>(q, bulk, sm_count, stream);
    case 4: return finish_typed<int>(q, bulk, sm_count, stream);
    case 8: return finish_typed<long long>(q, bulk, sm_count, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t rb_rows_convert(const int* in, void* out, size_t count, int sample_bytes, rb_u32* status, unsigned sm_count,
                            cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)std::min<size_t>((count / 4 + 255) / 256, (size_t)sm_count * 16);
  switch (sample_bytes) {
    case 2: rb_rows_convert_kernel<short><<<blocks ? blocks : 1, 256, 0, stream>>>(in, static_cast<short*>(out), count, status); break;
    case 8: rb_rows_convert_kernel<long long><<<blocks ? blocks : 1, 256, 0, stream>>>(in, static_cast<long long*>(out), count, status); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
