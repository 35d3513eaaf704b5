// network.hpp -- host-side description of a reaction network and its lowering.
//
// Mirrors the reference's `Vec<(Rate, Jump)>` (src/gillespie.rs:157-163) in the sparse forms
// (Rate::LMASparse / Jump::Sparse), which the reference pins as bit-identical to the dense ones
// (tests/test_rebop.py:55-65, src/gillespie.rs:448-472).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rebop_b200.h"
#include "rb_tables.h"

struct RbReaction {
  bool is_expr = false;
  double k = 0.0;
  std::vector<uint32_t> term_idx;  // reactant terms in evaluation order
  std::vector<uint32_t> term_exp;
  std::vector<rebop_expr_op> prog;  // post-order expression program (is_expr)
  std::vector<int64_t> diff;        // dense stoichiometry, length n_species
};

struct rebop_network {
  uint32_t n_species = 0;
  int arith = REBOP_ARITH_API;
  std::vector<RbReaction> rx;
};

// Thread-local error message shared by the whole library.
void rb_set_error(const std::string& msg);
int rb_fail(int status, const std::string& msg);

// Checks an expression program (well-formed stack discipline, indices in range, depth).
int rb_check_program(const rebop_expr_op* prog, size_t n_ops, uint32_t n_species);

// Fills the constant-memory image used by the table-driven kernel.  save_idx is left untouched.
int rb_lower_tables(const rebop_network& net, RbTables* out);
