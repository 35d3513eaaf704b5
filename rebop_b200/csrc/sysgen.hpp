// sysgen.hpp -- parsed `define_system!` text (src/gillespie_macro.rs:49-61).
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "network.hpp"

struct RbRateExpr {
  enum { CONST, PARAM, NEG, ADD, SUB, MUL, DIV };
  int kind = CONST;
  double value = 0.0;
  int index = 0;
  std::unique_ptr<RbRateExpr> a, b;
};

struct RbSystemReaction {
  std::string name;
  std::vector<std::pair<uint32_t, uint32_t>> lhs, rhs;  // (species index, coefficient) in the order written
  std::unique_ptr<RbRateExpr> rate;
};

struct rebop_system {
  std::string name;
  std::vector<std::string> params, species;
  std::vector<RbSystemReaction> reactions;
};

int rb_system_parse(const char* text, rebop_system* sys);
// Name::with_parameters + lowering in define_system! arithmetic.
int rb_system_network(const rebop_system& sys, const double* params, size_t n_params, rebop_network* net);

// Kernels compiled at build time (rebop_sysgen + nvcc) register themselves here, keyed by the
// source text the run-time generator would produce for the same network.
struct RbPrebuilt {
  const char* key;      // rb_codegen_source(net, "rb_ssa_jit") of the network it was generated from
  const void* grid_kernel[3];  // __global__ functions of the time-grid loop, one per RB_MODE_* schedule
  const void* kernel_evc;  // event-log mode: counting pass
  const void* kernel_evw;  // event-log mode: writing pass
  unsigned block, static_smem, net_words;
  const char* name;     // system name
};
void rb_register_prebuilt(const RbPrebuilt& entry);
const RbPrebuilt* rb_find_prebuilt(const std::string& key);
int rb_prebuilt_count();
