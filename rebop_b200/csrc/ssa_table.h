// ssa_table.h -- host entry of the table-driven kernel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#define RB_TABLE_BLOCK 128

struct RbTables;
struct SsaRunParams;

// Uploads the tables to __constant__ memory on `stream`, then launches.  The host image must
// stay valid until the copy has been issued (pageable memory: the call returns after staging).
cudaError_t rb_table_launch(const RbTables* host_tables, bool dynamic, const SsaRunParams& p, unsigned grid,
                            size_t smem_bytes, cudaStream_t stream);
// Resident CTAs per SM of the table-driven kernel with this much dynamic shared memory.
cudaError_t rb_table_occupancy(bool dynamic, size_t smem_bytes, int* ctas_per_sm);
// Event-log mode (nb_steps = 0): counting pass (write = false) or writing pass.
cudaError_t rb_table_launch_events(const RbTables* host_tables, bool write, const SsaRunParams& p, unsigned grid,
                                   size_t smem_bytes, cudaStream_t stream);
