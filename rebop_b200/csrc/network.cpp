// network.cpp -- network construction (add_reaction validation) and table lowering.
#include "network.hpp"

#include <cmath>
#include <cstring>
#include <limits>

static thread_local std::string g_last_error;

void rb_set_error(const std::string& msg) { g_last_error = msg; }
int rb_fail(int status, const std::string& msg) {
  g_last_error = msg;
  return status;
}

extern "C" const char* rebop_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* rebop_b200_version(void) { return "0.1.0 (rebop 0.9.7 semantics)"; }

int rb_check_program(const rebop_expr_op* prog, size_t n_ops, uint32_t n_species) {
  if (n_ops == 0 || prog == nullptr) return rb_fail(REBOP_ERR_INVALID, "empty expression program");
  int sp = 0, max_sp = 0;
  for (size_t i = 0; i < n_ops; ++i) {
    switch (prog[i].op) {
      case REBOP_OP_CONST: ++sp; break;
      case REBOP_OP_SPECIES:
        if (prog[i].index < 0 || (uint32_t)prog[i].index >= n_species)
          return rb_fail(REBOP_ERR_OUT_OF_RANGE, "expression refers to a species index out of range");
        ++sp;
        break;
      case REBOP_OP_NEG: case REBOP_OP_EXP:
        if (sp < 1) return rb_fail(REBOP_ERR_INVALID, "malformed expression program");
        break;
      case REBOP_OP_ADD: case REBOP_OP_SUB: case REBOP_OP_MUL: case REBOP_OP_DIV:
      case REBOP_OP_POW: case REBOP_OP_MAX: case REBOP_OP_MIN:
        if (sp < 2) return rb_fail(REBOP_ERR_INVALID, "malformed expression program");
        --sp;
        break;
      default: return rb_fail(REBOP_ERR_INVALID, "unknown expression opcode");
    }
    if (sp > max_sp) max_sp = sp;
  }
  if (sp != 1) return rb_fail(REBOP_ERR_INVALID, "malformed expression program");
  if (max_sp > RB_EXPR_STACK)
    return rb_fail(REBOP_ERR_LIMIT, "expression needs a deeper evaluation stack than the kernels provide");
  return REBOP_OK;
}

static int check_diff(const rebop_network* net, const int64_t* differences, RbReaction* rx) {
  if (!differences && net->n_species) return rb_fail(REBOP_ERR_INVALID, "differences is NULL");
  rx->diff.assign(differences, differences + net->n_species);
  for (int64_t d : rx->diff)
    if (d < -32767 || d > 32767)
      return rb_fail(REBOP_ERR_LIMIT, "stoichiometric difference outside [-32767, 32767]");
  return REBOP_OK;
}

extern "C" int rebop_network_create(uint32_t n_species, int arith, rebop_network** out) {
  if (!out) return rb_fail(REBOP_ERR_INVALID, "out is NULL");
  if (arith != REBOP_ARITH_API && arith != REBOP_ARITH_MACRO)
    return rb_fail(REBOP_ERR_INVALID, "unknown arithmetic mode");
  if (n_species > RB_TAB_MAX_SPECIES) return rb_fail(REBOP_ERR_LIMIT, "too many species");
  rebop_network* net = new rebop_network();
  net->n_species = n_species;
  net->arith = arith;
  *out = net;
  return REBOP_OK;
}

extern "C" void rebop_network_destroy(rebop_network* net) { delete net; }

extern "C" int rebop_network_add_reaction_lma_sparse(rebop_network* net, double k, const uint32_t* index,
                                                     const uint32_t* exponent, size_t n_terms,
                                                     const int64_t* differences) {
  if (!net) return rb_fail(REBOP_ERR_INVALID, "network is NULL");
  RbReaction rx;
  rx.k = k;
  for (size_t j = 0; j < n_terms; ++j) {
    // the reference asserts every sparse index is a known species (src/gillespie.rs:229-233)
    if (index[j] >= net->n_species)
      return rb_fail(REBOP_ERR_OUT_OF_RANGE, "assertion failed: reactant species index out of range");
    if (exponent[j] > 255) return rb_fail(REBOP_ERR_LIMIT, "reactant exponent above 255");
    if (net->arith == REBOP_ARITH_API && exponent[j] == 0) continue;  // empty factor range
    rx.term_idx.push_back(index[j]);
    rx.term_exp.push_back(exponent[j]);
  }
  int st = check_diff(net, differences, &rx);
  if (st) return st;
  net->rx.push_back(std::move(rx));
  return REBOP_OK;
}

extern "C" int rebop_network_add_reaction_lma(rebop_network* net, double k, const uint32_t* exponents,
                                              const int64_t* differences) {
  if (!net) return rb_fail(REBOP_ERR_INVALID, "network is NULL");
  if (!exponents && net->n_species) return rb_fail(REBOP_ERR_INVALID, "exponents is NULL");
  // Rate::sparse (src/gillespie.rs:55-69): ascending species index, zero exponents dropped.
  std::vector<uint32_t> idx, ex;
  for (uint32_t s = 0; s < net->n_species; ++s)
    if (exponents[s] > 0) {
      idx.push_back(s);
      ex.push_back(exponents[s]);
    }
  return rebop_network_add_reaction_lma_sparse(net, k, idx.data(), ex.data(), idx.size(), differences);
}

extern "C" int rebop_network_add_reaction_expr(rebop_network* net, const rebop_expr_op* program,
                                               size_t n_ops, const int64_t* differences) {
  if (!net) return rb_fail(REBOP_ERR_INVALID, "network is NULL");
  int st = rb_check_program(program, n_ops, net->n_species);
  if (st) return st;
  RbReaction rx;
  rx.is_expr = true;
  rx.k = std::numeric_limits<double>::quiet_NaN();
  rx.prog.assign(program, program + n_ops);
  st = check_diff(net, differences, &rx);
  if (st) return st;
  net->rx.push_back(std::move(rx));
  return REBOP_OK;
}

extern "C" int rebop_network_nb_species(const rebop_network* net, uint32_t* out) {
  if (!net || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *out = net->n_species;
  return REBOP_OK;
}

extern "C" int rebop_network_nb_reactions(const rebop_network* net, uint32_t* out) {
  if (!net || !out) return rb_fail(REBOP_ERR_INVALID, "NULL argument");
  *out = (uint32_t)net->rx.size();
  return REBOP_OK;
}

int rb_lower_tables(const rebop_network& net, RbTables* t) {
  std::memset(t, 0, sizeof *t);
  if (net.rx.size() > RB_TAB_MAX_REACTIONS)
    return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: more than 1024 reactions");
  t->n_species = (int)net.n_species;
  t->n_reactions = (int)net.rx.size();
  t->arith = net.arith;
  size_t nt = 0, nj = 0, no = 0;
  for (size_t r = 0; r < net.rx.size(); ++r) {
    const RbReaction& rx = net.rx[r];
    t->k[r] = rx.k;
    t->term_ptr[r] = (unsigned short)nt;
    t->jump_ptr[r] = (unsigned short)nj;
    t->expr_ptr[r] = (unsigned short)no;
    if (rx.is_expr) {
      if (no + rx.prog.size() > RB_TAB_MAX_OPS)
        return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: expression programs exceed 512 operations");
      for (const rebop_expr_op& op : rx.prog) {
        t->op_code[no] = (unsigned char)op.op;
        t->op_idx[no] = (unsigned short)(op.op == REBOP_OP_SPECIES ? op.index : 0);
        t->op_val[no] = op.value;
        ++no;
      }
    } else {
      if (nt + rx.term_idx.size() > RB_TAB_MAX_TERMS)
        return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: more than 3072 reactant terms");
      for (size_t j = 0; j < rx.term_idx.size(); ++j) {
        t->term_idx[nt] = (unsigned short)rx.term_idx[j];
        t->term_exp[nt] = (unsigned char)rx.term_exp[j];
        ++nt;
      }
    }
    for (uint32_t s = 0; s < net.n_species; ++s) {
      if (rx.diff[s] == 0) continue;  // Jump::sparse (src/gillespie.rs:125-140)
      if (nj >= RB_TAB_MAX_JUMPS)
        return rb_fail(REBOP_ERR_LIMIT, "table-driven kernel: more than 4096 stoichiometry entries");
      t->jump_idx[nj] = (unsigned short)s;
      t->jump_diff[nj] = (short)rx.diff[s];
      ++nj;
    }
  }
  const size_t R = net.rx.size();
  t->term_ptr[R] = (unsigned short)nt;
  t->jump_ptr[R] = (unsigned short)nj;
  t->expr_ptr[R] = (unsigned short)no;
  return REBOP_OK;
}
