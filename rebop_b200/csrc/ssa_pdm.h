// ssa_pdm.h -- K6: the dependency-driven kernel for large mass-action networks (tier-2 parity: statistically
// exact, not stream-exact).  Host side: lowering of a network into the partial-propensity tables, launches.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#define RB_PDM_BLOCK 64
#define RB_PDM_SUPER 8          // groups per super-group (two-level sum tree)
#define RB_PDM_REFRESH 4096u    // passes between full rebuilds of the partial propensities and sums

struct SsaRunParams;
struct rebop_network;

// Device image of the tables (one flat buffer of 8-byte words; offsets in words from the start).
//   c[n_groups]                      pi_i at x = 0 (sum of first-order constants, minus the 2A constants)
//   col_ptr[n_groups + 1] (u32)      CSC over the CHANGED species s: entries (K, i) with pi_i += K * dx_s
//   col[nnz]      {double K; u32 i; u32 pad}
//   own_ptr[n_groups + 1] (u32)      reactions owned by group i, for the choice inside a group
//   own[n_reactions] {double k; u32 reaction; u32 partner}   partner: RB_PDM_NONE (a = x_i k), i (a = x_i k (x_i - 1)),
//                                                            or j (a = x_i k x_j)
struct RbPdmHeader {
  uint32_t n_groups;   // species + 1: the last group owns the zeroth-order reactions (its "count" is the constant 1)
  uint32_t n_super;    // ceil(n_groups / RB_PDM_SUPER)
  uint32_t off_c, off_col_ptr, off_col, off_own_ptr, off_own;
  uint32_t n_reactions;
};
#define RB_PDM_NONE 0xffffffffu

// Shared memory words (32-bit) per CTA of the kernel for a network with `n_species` species.
static inline unsigned rb_pdm_net_words(unsigned n_species) {
  const unsigned ng = n_species + 1, nsg = (ng + RB_PDM_SUPER - 1) / RB_PDM_SUPER;
  return RB_PDM_BLOCK * (2u * ng + 2u * nsg + ng);  // pi (f64), super sums (f64), counts (int32)
}

// Lowers `net` into the image above.  REBOP_OK, or REBOP_ERR_LIMIT with the reason in *why: the partial-propensity
// form needs elementary mass action (total reactant order <= 2, no expression rates), rate constants >= 0 and at
// most four species changed per reaction.
int rb_pdm_build(const rebop_network& net, std::vector<uint64_t>* image, std::string* why);

cudaError_t rb_pdm_launch(int mode, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t rb_pdm_occupancy(int mode, size_t smem_bytes, int* ctas_per_sm);
