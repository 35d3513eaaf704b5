// ssa_table.h -- host entry of the table-driven kernel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#define RB_TABLE_BLOCK 128

struct SsaRunParams;

// Launches the table-driven kernel in schedule `mode` (RB_MODE_*).  The network travels in the launch
// parameters (SsaRunParams::gtab, ::tables: device images owned by the batch).
cudaError_t rb_table_launch(int mode, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream);
// Resident CTAs per SM of the table-driven kernel with this much dynamic shared memory.
cudaError_t rb_table_occupancy(int mode, size_t smem_bytes, int* ctas_per_sm);
// Event-log mode (nb_steps = 0): counting pass (write = false) or writing pass.
cudaError_t rb_table_launch_events(bool write, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream);
