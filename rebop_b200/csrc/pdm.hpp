// pdm.hpp -- partial-propensity form of a mass-action network (host lowering shared by the code generator and the
// engine).  See codegen.cpp (rb_codegen_pdm_source) for what the kernel does with it.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "network.hpp"

#define RB_PDM_NONE 0xffffffffu
#define RB_PDM_MAX_CHECKPOINTS 50  // 100 owner groups in blocks of two: 1.45e10 vs 1.40e10 events/s with 32 (profiles/r2af_pdm.log)

// Every reaction is owned by its first reactant i (zeroth-order reactions by a pseudo-species whose count is the
// constant 1, stored as species index n_species):
//      a_r = x_i * k_r              (A -> ..)
//      a_r = x_i * k_r * x_j        (A + B -> ..)
//      a_r = x_i * k_r * (x_i - 1)  (2A -> ..; the reference's falling factorial, src/gillespie.rs:73-87)
// so the propensities owned by i sum to x_i * pi_i with pi_i = c_i + sum_j K_ij x_j.
struct RbPdmGroup {
  uint32_t species;             // owner i
  uint32_t const_first;         // consts[const_first] = c_i, followed by one K_ij per entry of `partners`
  std::vector<uint32_t> partners;  // j of every K_ij, in evaluation order
  struct Own { double k; uint32_t reaction, partner; };  // partner: RB_PDM_NONE, i itself (2A) or j
  std::vector<Own> own;         // the reactions of this group, for the choice inside the group
};

struct RbPdmLowered {
  std::vector<RbPdmGroup> groups;  // owners that have at least one reaction, ascending species index
  std::vector<double> consts;      // c_i and K_ij of every group in evaluation order (travel in SsaRunParams::k)
  unsigned group_size = 1;         // groups per checkpoint of the running sum
  unsigned n_checkpoints = 1;
  // device image for the choice (one flat buffer of 8-byte words):
  //   header {u32 n_blocks, off_block_ptr, off_entries, n_entries}
  //   block_ptr[n_blocks + 1]  u32: the reactions owned by the groups of checkpoint block b are entries [ptr[b], ptr[b+1])
  //   entries[]                {f64 k; u32 i | j << 16; u32 reaction | kind << 30}  (16 bytes), kind 0: a = k x_i,
  //                            1: a = k x_i x_j, 2: a = k x_i (x_i - 1); blocks in the order of the unrolled pass, inside a block heaviest first
  std::vector<uint64_t> image;
};

// REBOP_OK, or REBOP_ERR_LIMIT with the reason in *why: the form needs elementary mass action (total reactant order
// <= 2, no expression rates), rate constants >= 0, at most four species changed per reaction, and reactions that
// cannot drive a count negative.
int rb_pdm_lower(const rebop_network& net, RbPdmLowered* out, std::string* why);
