// codegen.hpp -- network -> CUDA source of a network-specialised ensemble kernel.
#pragma once
#include <string>

#include "network.hpp"

// Register-resident state and cumulative sums bound the networks that can be specialised;
// larger ones use the table-driven kernel (shared-memory state).
#define RB_GEN_MAX_SPECIES 64
#define RB_GEN_MAX_REACTIONS 96
// ... and together they must fit a thread's 255 registers: 2 per species + 2 per reaction + loop state
#define RB_GEN_MAX_REGISTERS 255
#define RB_GEN_LOOP_REGISTERS 86
// what the compiler needs next to state and cumulative rates when the launch bound asks it to be tight: 40
// registers of loop state plus temporaries that grow with the width of the stoichiometry rows
// (Vilar: 18 + 32 + 40 + 6 = 96 registers at 5 resident CTAs); used to choose the launch bound
#define RB_GEN_LOOP_REGISTERS_TIGHT(S, R) (40u + ((S) + (R)) / 4u)
// Large form (f64 state columns in shared memory): 32-thread CTAs hold up to 100 KB / (32 * 8 B) species.
#define RB_GEN_LARGE_MAX_SPECIES 400

struct RbCodegenInfo {
  unsigned block = 128;      // threads per CTA the kernel was generated for
  unsigned net_words = 0;    // 32-bit words of dynamic shared memory the network needs (none)
  unsigned static_smem = 0;  // bytes of static shared memory for the packed stoichiometry table
  bool uses_param_k = true;  // LMA rate constants are read from SsaRunParams::k
  bool large = false;        // large form: needs SsaRunParams::gtab (reaction records + saved-species list)
};

bool rb_codegen_supported(const rebop_network& net, std::string* why);
// True when the network gets the large form (state in shared memory) instead of register-resident state.
bool rb_codegen_is_large(const rebop_network& net);
// Source text of `extern "C" __global__ void <kernel_name>(SsaRunParams)`; includes "ssa_kernel.cuh".
// `variant` selects the entry points: the time-grid kernels <name>, <name>_dyn and/or the event-log kernels
// <name>_evc, <name>_evw.  The grid variant of "rb_ssa_jit" is the text build-time kernels are registered under.
#define RB_VARIANT_GRID 1
#define RB_VARIANT_EVENTS 2
std::string rb_codegen_source(const rebop_network& net, const std::string& kernel_name, RbCodegenInfo* info,
                              int variant = RB_VARIANT_GRID);

// Partial-propensity form of a mass-action network (opt-in kernel REBOP_KERNEL_PDM, tier-2 parity); `low` comes from
// rb_pdm_lower (pdm.hpp).  Entry points <name>, <name>_dyn, <name>_dns.
struct RbPdmLowered;
std::string rb_codegen_pdm_source(const rebop_network& net, const RbPdmLowered& low, const std::string& kernel_name,
                                  RbCodegenInfo* info);
