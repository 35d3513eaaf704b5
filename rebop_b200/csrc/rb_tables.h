// rb_tables.h -- constant-memory image of a lowered network for the table-driven kernel.
// Shared between the host lowering (network.cpp) and the device code (ssa_table.cu).
#pragma once

#define RB_TAB_MAX_REACTIONS 1024
#define RB_TAB_MAX_TERMS 3072
#define RB_TAB_MAX_JUMPS 4096
#define RB_TAB_MAX_OPS 512
#define RB_TAB_MAX_SAVE 1024
#define RB_TAB_MAX_SPECIES 65535
#define RB_EXPR_STACK 16

// Expression byte-code: post-order walk of the reference's `Expr` tree (src/expr.rs:9-21).
enum RbOp {
  RB_OP_CONST = 0, RB_OP_SPECIES = 1, RB_OP_NEG = 2, RB_OP_ADD = 3, RB_OP_SUB = 4, RB_OP_MUL = 5,
  RB_OP_DIV = 6, RB_OP_POW = 7, RB_OP_MAX = 8, RB_OP_MIN = 9, RB_OP_EXP = 10
};

struct RbTables {
  double k[RB_TAB_MAX_REACTIONS];        // rate constant of LMA reactions
  double op_val[RB_TAB_MAX_OPS];         // constants of expression programs
  unsigned short term_ptr[RB_TAB_MAX_REACTIONS + 1];  // CSR: reactant terms of reaction r
  unsigned short jump_ptr[RB_TAB_MAX_REACTIONS + 1];  // CSR: stoichiometry of reaction r
  unsigned short expr_ptr[RB_TAB_MAX_REACTIONS + 1];  // CSR: expression program (empty => LMA)
  unsigned short term_idx[RB_TAB_MAX_TERMS];
  unsigned short jump_idx[RB_TAB_MAX_JUMPS];
  short jump_diff[RB_TAB_MAX_JUMPS];
  unsigned short op_idx[RB_TAB_MAX_OPS];
  unsigned short save_idx[RB_TAB_MAX_SAVE];
  unsigned char term_exp[RB_TAB_MAX_TERMS];
  unsigned char op_code[RB_TAB_MAX_OPS];
  int n_species, n_reactions;
  int arith;  // 0: function API arithmetic + count select; 1: define_system! arithmetic + first match
};
