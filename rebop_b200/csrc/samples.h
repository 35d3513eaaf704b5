// samples.h -- host entry of the sample-finishing kernels (samples.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "ssa_params.h"

struct RbFinishParams {
  const int* raw;            // [n][n_points][n_save] rows as the trajectories wrote them
  const rb_u32* progress;    // [ldn] RB_PROGRESS_*: rows each trajectory wrote itself
  void* out;                 // [rows][ldn] samples of `sample_bytes` each (first row of this launch's grid points)
  rb_i64* sums;              // [rows] sums, then at sums + sums_stride the sums of squares; accumulated into; or null
  rb_u32* status;            // |= RB_STATUS_NARROW
  rb_u32 n, ldn, n_points, n_save, rows;  // rows = n_points * n_save
  rb_u32 sums_stride;
};

// raw records -> [row][trajectory] samples (int16 / int32 / int64 by sample_bytes), row sums fused in.
// bulk: write the lines with TMA bulk stores (cp.async.bulk) instead of 16-byte vector stores.
cudaError_t rb_samples_finish(const RbFinishParams& q, int sample_bytes, bool bulk, unsigned sm_count, cudaStream_t stream);
// [count] int32 -> int16 / int64 (count a multiple of 4).
cudaError_t rb_rows_convert(const int* in, void* out, size_t count, int sample_bytes, rb_u32* status, unsigned sm_count,
                            cudaStream_t stream);
