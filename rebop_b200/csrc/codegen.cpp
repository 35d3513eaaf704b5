// codegen.cpp -- network -> CUDA source of a network-specialised ensemble kernel (K2).
//
// This is the GPU analogue of what `define_system!` does at the user's compile time
// (src/gillespie_macro.rs:49-129): species become scalars (registers), every propensity is one
// straight-line expression, the cumulative sum and the reaction choice are fully unrolled.
// The same generator serves the run-time path (NVRTC, jit.cpp) and the build-time path
// (rebop_sysgen, sysgen_main.cpp -> nvcc).  It has no CUDA dependency.
#include "codegen.hpp"
#include "pdm.hpp"

#include "ssa_params.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <sstream>

namespace {

std::string f64_bits(double v) {
  unsigned long long bits;
  std::memcpy(&bits, &v, sizeof bits);
  char buf[64];
  std::snprintf(buf, sizeof buf, "__longlong_as_double(0x%016llxll)", bits);
  return buf;
}

// Rate::Expr -> nested intrinsic calls; same operations, same order as Expr::eval (src/expr.rs:24-38).
std::string emit_expr(const std::vector<rebop_expr_op>& prog) {
  std::vector<std::string> st;
  for (const rebop_expr_op& op : prog) {
    if (op.op == REBOP_OP_CONST) {
      st.push_back(f64_bits(op.value));
    } else if (op.op == REBOP_OP_SPECIES) {
      st.push_back("d" + std::to_string(op.index));
    } else if (op.op == REBOP_OP_NEG) {
      st.back() = "(-(" + st.back() + "))";
    } else if (op.op == REBOP_OP_EXP) {
      st.back() = "exp(" + st.back() + ")";
    } else {
      std::string b = st.back();
      st.pop_back();
      std::string a = st.back();
      const char* fn = "";
      switch (op.op) {
        case REBOP_OP_ADD: fn = "__dadd_rn"; break;
        case REBOP_OP_SUB: fn = "__dsub_rn"; break;
        case REBOP_OP_MUL: fn = "__dmul_rn"; break;
        case REBOP_OP_DIV: fn = "__ddiv_rn"; break;
        case REBOP_OP_POW: fn = "pow"; break;
        case REBOP_OP_MAX: fn = "fmax"; break;
        case REBOP_OP_MIN: fn = "fmin"; break;
      }
      st.back() = std::string(fn) + "(" + a + ", " + b + ")";
    }
  }
  return st.empty() ? "0.0" : st.back();
}

}  // namespace

// Entry points of a generated translation unit.  RB_VARIANT_GRID: static and dynamic schedule of the
// time-grid loop; RB_VARIANT_EVENTS: counting and writing pass of the event-log mode.
static void emit_kernels(std::ostringstream& o, const std::string& kernel_name, unsigned block, unsigned minctas, int variant) {
  struct Entry { const char* suffix; const char* body; int in; };
  const Entry entries[] = {
      {"", "rb_ssa_loop<RbGenNet, RB_MODE_STATIC>(net, p, rb_smem);", RB_VARIANT_GRID},
      {"_dyn", "rb_ssa_loop<RbGenNet, RB_MODE_SPARSE>(net, p, rb_smem);", RB_VARIANT_GRID},
      {"_dns", "rb_ssa_loop<RbGenNet, RB_MODE_DENSE>(net, p, rb_smem);", RB_VARIANT_GRID},
      {"_evc", "rb_ssa_events<RbGenNet, false>(net, p, rb_smem);", RB_VARIANT_EVENTS},
      {"_evw", "rb_ssa_events<RbGenNet, true>(net, p, rb_smem);", RB_VARIANT_EVENTS},
  };
  for (const Entry& e : entries) {
    if (!(variant & e.in)) continue;
    o << "extern \"C\" __global__ void __launch_bounds__(" << block;
    if (minctas) o << ", " << minctas;
    o << ") " << kernel_name << e.suffix << "(const __grid_constant__ SsaRunParams p) {\n";
    o << "  extern __shared__ __align__(16) int rb_smem[];\n";
    o << "  RbGenNet net;\n";
    o << "  " << e.body << "\n";
    o << "}\n";
  }
}

static bool small_form_ok(const rebop_network& net) {
  unsigned max_s = RB_GEN_MAX_SPECIES, max_r = RB_GEN_MAX_REACTIONS;
  if (const char* env = std::getenv("REBOP_B200_CODEGEN")) {  // development knobs: maxs=, maxr=
    const std::string e(env);
    size_t pos = e.find("maxs=");
    if (pos != std::string::npos) max_s = (unsigned)std::atoi(e.c_str() + pos + 5);
    pos = e.find("maxr=");
    if (pos != std::string::npos) max_r = (unsigned)std::atoi(e.c_str() + pos + 5);
  }
  return net.n_species <= max_s && net.rx.size() <= max_r &&
         2 * net.n_species + 2 * net.rx.size() + RB_GEN_LOOP_REGISTERS <= RB_GEN_MAX_REGISTERS;
}

// Large form: state as f64 columns in shared memory, unrolled propensity pass with checkpoints,
// block re-walk from per-reaction records (ssa_kernel.cuh: rb_large_fire).
static bool large_form_ok(const rebop_network& net, std::string* why) {
  auto no = [&](const std::string& msg) {
    if (why) *why = msg;
    return false;
  };
  if (net.n_species > RB_GEN_LARGE_MAX_SPECIES) return no("more than " + std::to_string(RB_GEN_LARGE_MAX_SPECIES) + " species (shared-memory state)");
  if (net.rx.size() > RB_MAX_K) return no("more than " + std::to_string(RB_MAX_K) + " reactions (rate constants in the launch parameters)");
  for (const RbReaction& rx : net.rx) {
    if (rx.is_expr) return no("expression rates in a network too large for register-resident state");
    if (rx.term_idx.size() > 2) return no("a reaction with more than two reactant terms in a large network");
    for (uint32_t e : rx.term_exp)
      if (e > 2) return no("a reactant exponent above 2 in a large network");
    int nz = 0;
    for (int64_t d : rx.diff) nz += d != 0;
    if (nz > 4) return no("a reaction changing more than four species in a large network");
    // non-decreasing cumulative rates need counts that cannot go negative: every consumed species
    // must be a reactant of at least that order (the propensity is 0 before the count would cross 0)
    for (size_t sp = 0; sp < rx.diff.size(); ++sp) {
      if (rx.diff[sp] >= 0) continue;
      int64_t order = 0;
      for (size_t j = 0; j < rx.term_idx.size(); ++j)
        if (rx.term_idx[j] == sp) order += rx.term_exp[j] ? rx.term_exp[j] : 1;
      if (order < -rx.diff[sp]) return no("a reaction consumes more of a species than its reactant order (counts could go negative)");
    }
  }
  return true;
}

bool rb_codegen_supported(const rebop_network& net, std::string* why) {
  return small_form_ok(net) || large_form_ok(net, why);
}

bool rb_codegen_is_large(const rebop_network& net) { return !small_form_ok(net); }

static std::string large_source(const rebop_network& net, const std::string& kernel_name, RbCodegenInfo* info, int variant) {
  const int S = (int)net.n_species;
  const int R = (int)net.rx.size();
  const bool macro = net.arith == REBOP_ARITH_MACRO;
  // CTA size: two CTAs of f64 state columns must fit in an SM's shared memory
  unsigned block = 128;
  while (block > 32 && (size_t)S * block * 8 > 100 * 1024) block /= 2;
  int ckn = 8;
  while ((R + ckn - 1) / ckn > 32) ckn *= 2;  // at most 32 checkpoints (64 registers)
  const int nck = (R + ckn - 1) / ckn;
  unsigned tick = 16;
  if (const char* env = std::getenv("REBOP_B200_CODEGEN")) {
    const std::string e(env);
    const size_t pos = e.find("lblock=");
    if (pos != std::string::npos) block = (unsigned)std::atoi(e.c_str() + pos + 7);
  }
  if (info) {
    info->block = block;
    info->net_words = (unsigned)(S * block * 2);
    info->static_smem = 16;
    info->large = true;
  }
  std::ostringstream o;
  o << "// generated by rebop_b200 codegen (large form): " << S << " species, " << R << " reactions, "
    << (macro ? "define_system! arithmetic" : "function-API arithmetic") << ", " << nck << " checkpoints of " << ckn
    << " reactions\n";
  o << "#define RB_TICK " << tick << "u\n";
  o << "#include \"ssa_kernel.cuh\"\n\n";
  o << "struct RbGenNet {\n";
  o << "  static constexpr int BLOCK = " << block << ";\n";
  o << "  static constexpr bool NAN_PICKS_NONE = false;\n";
  o << "  double* xs;  // this thread's column of the f64 state: species s at xs[s * BLOCK]\n";
  o << "  double ck[" << nck << "];  // cumulative rate after every " << ckn << " reactions\n";
  o << "  static __device__ __forceinline__ int smem_words(const SsaRunParams&) { return " << S * (int)block * 2 << "; }\n";
  o << "  __device__ __forceinline__ void init(const SsaRunParams&, int* smem, rb_u32 tid, rb_u32) {\n";
  o << "    xs = reinterpret_cast<double*>(smem) + tid;\n  }\n";
  o << "  __device__ __forceinline__ void load(const SsaRunParams& p, rb_u32 traj, bool valid) {\n";
  o << "    for (int s = 0; s < " << S << "; ++s) xs[s * BLOCK] = valid ? (double)p.x[(size_t)s * p.ldn + traj] : 0.0;\n  }\n";
  o << "  __device__ __forceinline__ void store(const SsaRunParams& p, rb_u32 traj) {\n";
  o << "    for (int s = 0; s < " << S << "; ++s) p.x[(size_t)s * p.ldn + traj] = __double2int_rn(xs[s * BLOCK]);\n  }\n";
  // unrolled propensities; the arithmetic of every term is the one rb_large_term re-walks
  o << "  double tot;\n  __device__ __forceinline__ double& total_ref(const SsaRunParams& p) {\n    tot = propensities(p);\n    return tot;\n  }\n";
  o << "  __device__ __forceinline__ double propensities(const SsaRunParams& p) {\n";
  o << "    double c = 0.0, a;\n";
  if (R == 0) o << "    (void)a;\n";
  for (int r = 0; r < R; ++r) {
    const RbReaction& rx = net.rx[r];
    std::string a = "p.k[" + std::to_string(r) + "]";
    for (size_t j = 0; j < rx.term_idx.size(); ++j) {
      const std::string x = "xs[" + std::to_string(rx.term_idx[j]) + " * BLOCK]";
      const unsigned e = rx.term_exp[j];
      if (e <= 1) a = "__dmul_rn(" + a + ", " + x + ")";
      else if (macro) a = "__dmul_rn(" + a + ", __dmul_rn(" + x + ", __dsub_rn(" + x + ", 1.0)))";
      else a = "__dmul_rn(__dmul_rn(" + a + ", __dsub_rn(" + x + ", 1.0)), " + x + ")";
    }
    o << "    a = " << a << ";\n";
    o << "    c = __dadd_rn(c, a);\n";
    if (r % ckn == ckn - 1 || r == R - 1) o << "    ck[" << r / ckn << "] = c;\n";
  }
  o << "    return c;\n  }\n";
  o << "  __device__ __forceinline__ int select(const SsaRunParams& p, double chosen) const {\n";
  if (R == 0) o << "    return 0;\n";
  else o << "    return rb_large_select<" << nck << ", " << ckn << ", " << R << ", " << (macro ? "true" : "false")
         << ", BLOCK>(ck, chosen, xs, p.gtab);\n";
  o << "  }\n";
  o << "  __device__ __forceinline__ void apply(const SsaRunParams& p, int pick, rb_u32& nev) {\n";
  if (R == 0) o << "    (void)pick; (void)nev;\n";
  else o << "    if (rb_large_apply<" << R << ", BLOCK>(pick, xs, p.gtab)) ++nev;\n";
  o << "  }\n";
  o << "  static __device__ __forceinline__ int none() { return " << R << "; }\n";
  o << "  __device__ __forceinline__ void record(const SsaRunParams& p, int* dst, rb_u32 stride) const {\n";
  o << "    const rb_u32* save = p.gtab + " << R * 8 << ";\n";
  o << "    for (rb_u32 j = 0; j < p.n_save; ++j) dst[(size_t)j * stride] = __double2int_rn(xs[__ldg(save + j) * BLOCK]);\n";
  o << "  }\n};\n\n";
  emit_kernels(o, kernel_name, block, 2, variant);
  return o.str();
}

// Partial-propensity form (REBOP_KERNEL_PDM; tier-2 parity -- statistically exact, not stream-exact).
//
// The reference recomputes every propensity at every event (src/gillespie.rs:357-364); its Python docstring promises
// "heuristics" for large systems (python/rebop/gillespie.py:123-126) that the crate does not have
// (src/pyo3_gillespie.rs:161).  For elementary mass action the sum of the propensities factors by owner species:
// total = sum_i x_i * pi_i with pi_i = c_i + sum_j K_ij x_j (pdm.hpp), i.e. one fused multiply-add per owner species
// and per distinct reactant pair instead of two to three multiplies and an add per reaction -- 252 instead of 1661
// FP64 operations per event for the synthetic 100 x 500 network -- and the pass stays the same straight-line code for
// every lane.  It is still the direct method: the same two random numbers per event in the reference's order, the
// same distributions of waiting time and reaction choice; only the order (and fusing) of the floating-point sums
// differs from the reference's running sum, so a trajectory eventually takes a different reaction than the reference
// would from the same random word.  Opt-in, validated against the oracle by ensemble statistics.
std::string rb_codegen_pdm_source(const rebop_network& net, const RbPdmLowered& low, const std::string& kernel_name,
                                  RbCodegenInfo* info) {
  const int S = (int)net.n_species;
  const int R = (int)net.rx.size();
  unsigned block = 128;
  while (block > 32 && (size_t)(S + 1) * block * 8 > 104960) block /= 2;  // two CTAs of f64 state columns per SM
  if (info) {
    info->block = block;
    info->net_words = (unsigned)((S + 1) * block * 2);
    info->static_smem = 16;
    info->large = true;
  }
  const unsigned nck = low.n_checkpoints, gs = low.group_size;
  std::ostringstream o;
  o << "// generated by rebop_b200 codegen (partial-propensity form): " << S << " species, " << R << " reactions, "
    << low.groups.size() << " owner groups, " << low.consts.size() << " constants, " << nck << " checkpoints of " << gs << " groups\n";
  o << "#define RB_TICK 16u\n";
  o << "#include \"ssa_kernel.cuh\"\n\n";
  o << "struct RbGenNet {\n";
  o << "  static constexpr int BLOCK = " << block << ";\n";
  o << "  static constexpr bool NAN_PICKS_NONE = false;\n";
  o << "  double* xs;  // this thread's column of the f64 state: species s at xs[s * BLOCK]; xs[" << S << " * BLOCK] is the constant 1\n";
  o << "  double ck[" << nck << "];  // running sum of x_i * pi_i after every " << gs << " owner groups\n";
  o << "  static __device__ __forceinline__ int smem_words(const SsaRunParams&) { return " << (S + 1) * (int)block * 2 << "; }\n";
  o << "  __device__ __forceinline__ void init(const SsaRunParams&, int* smem, rb_u32 tid, rb_u32) {\n";
  o << "    xs = reinterpret_cast<double*>(smem) + tid;\n  }\n";
  o << "  __device__ __forceinline__ void load(const SsaRunParams& p, rb_u32 traj, bool valid) {\n";
  o << "    for (int s = 0; s < " << S << "; ++s) xs[s * BLOCK] = valid ? (double)p.x[(size_t)s * p.ldn + traj] : 0.0;\n";
  o << "    xs[" << S << " * BLOCK] = valid ? 1.0 : 0.0;\n  }\n";
  o << "  __device__ __forceinline__ void store(const SsaRunParams& p, rb_u32 traj) {\n";
  o << "    for (int s = 0; s < " << S << "; ++s) p.x[(size_t)s * p.ldn + traj] = __double2int_rn(xs[s * BLOCK]);\n  }\n";
  o << "  double tot;\n  __device__ __forceinline__ double& total_ref(const SsaRunParams& p) {\n    tot = propensities(p);\n    return tot;\n  }\n";
  o << "  __device__ __forceinline__ double propensities(const SsaRunParams& p) {\n";
  o << "    double c = 0.0, pi;\n";
  if (low.groups.empty()) o << "    (void)pi; (void)p;\n    ck[0] = c;\n";
  for (size_t g = 0; g < low.groups.size(); ++g) {
    const RbPdmGroup& grp = low.groups[g];
    o << "    pi = p.k[" << grp.const_first << "];\n";
    for (size_t e = 0; e < grp.partners.size(); ++e)
      o << "    pi = fma(p.k[" << grp.const_first + 1 + e << "], xs[" << grp.partners[e] << " * BLOCK], pi);\n";
    o << "    c = fma(xs[" << grp.species << " * BLOCK], pi, c);\n";
    if (g % gs == gs - 1 || g + 1 == low.groups.size()) o << "    ck[" << g / gs << "] = c;\n";
  }
  o << "    return c;\n  }\n";
  o << "  __device__ __forceinline__ int select(const SsaRunParams& p, double chosen) const {\n";
  o << "    return rb_pp_select<" << nck << ", BLOCK>(ck, chosen, xs, p.pdm, " << R << ");\n  }\n";
  o << "  __device__ __forceinline__ void apply(const SsaRunParams& p, int pick, rb_u32& nev) {\n";
  if (R == 0) o << "    (void)p; (void)pick; (void)nev;\n";
  else o << "    if (rb_large_apply<" << R << ", BLOCK>(pick, xs, p.gtab)) ++nev;\n";
  o << "  }\n";
  o << "  static __device__ __forceinline__ int none() { return " << R << "; }\n";
  o << "  __device__ __forceinline__ void record(const SsaRunParams& p, int* dst, rb_u32 stride) const {\n";
  o << "    const rb_u32* save = p.gtab + " << R * 8 << ";\n";
  o << "    for (rb_u32 j = 0; j < p.n_save; ++j) dst[(size_t)j * stride] = __double2int_rn(xs[__ldg(save + j) * BLOCK]);\n";
  o << "  }\n};\n\n";
  emit_kernels(o, kernel_name, block, 2, RB_VARIANT_GRID);
  return o.str();
}

std::string rb_codegen_source(const rebop_network& net, const std::string& kernel_name, RbCodegenInfo* info, int variant) {
  if (!small_form_ok(net)) return large_source(net, kernel_name, info, variant);
  const int S = (int)net.n_species;
  const int R = (int)net.rx.size();
  const bool macro = net.arith == REBOP_ARITH_MACRO;

  // stoichiometry packing: int8 lanes (dp4a) unless some |difference| needs int16 (dp2a)
  bool wide = false;
  for (const RbReaction& rx : net.rx)
    for (int64_t d : rx.diff)
      if (d < -127 || d > 127) wide = true;
  // lane S of every row counts the event: 1 for a reaction, 0 for the all-zero "no reaction" row, so that the
  // event counter is updated by the same instruction kind as the species (no test of the pick in the loop)
  const int per_word = wide ? 2 : 4;
  int dw = (S + 1 + per_word - 1) / per_word;
  int dwp = dw <= 2 ? dw : (dw + 3) / 4 * 4;  // 1, 2 or a multiple of 4 words per reaction
  std::vector<bool> touched(S, false);
  for (const RbReaction& rx : net.rx)
    for (int s = 0; s < S; ++s)
      if (rx.diff[s] != 0) touched[s] = true;

  // development knobs (REBOP_B200_CODEGEN="block=128,minctas=0,tick=32"); the defaults are the tuned values
  // (tick 32: a lane whose trajectory ends leaves the pass loop until the next tick, so a longer tick costs idle lane
  // slots and saves tick code; +1 % on Vilar, Dimers and Michaelis-Menten, profiles/r2v_sweep.log)
  // Resident CTAs asked of the compiler: as many (up to 5 = 96 registers per thread, the Vilar sweet spot)
  // as leave room for the register-resident state (2 per species), cumulative rates (2 per reaction) and
  // ~40 registers of loop state; larger networks get fewer CTAs instead of spills.
  unsigned block = 128, minctas = 5, tick = 32, unroll = 1, conv = 0, prescale = 1;
  {
    const unsigned need = 2u * (unsigned)S + 2u * (unsigned)R + RB_GEN_LOOP_REGISTERS_TIGHT((unsigned)S, (unsigned)R);
    while (minctas > 1 && std::min(255u, 65536u / (128u * minctas) / 8u * 8u) < need) --minctas;
  }
  if (const char* env = std::getenv("REBOP_B200_CODEGEN")) {
    const std::string e(env);
    auto get = [&](const char* key, unsigned def) {
      const size_t pos = e.find(std::string(key) + "=");
      return pos == std::string::npos ? def : (unsigned)std::atoi(e.c_str() + pos + std::strlen(key) + 1);
    };
    block = get("block", block);
    minctas = get("minctas", minctas);
    tick = get("tick", tick);
    unroll = get("unroll", unroll);
    conv = get("conv", conv);
    prescale = get("prescale", prescale);
  }
  if (info) {
    info->block = block;
    info->net_words = 0;  // the stoichiometry table is static shared memory of the generated kernel
    // + the second half of the wide ziggurat tables (RB_ZIG_WIDE, ssa_kernel.cuh) beyond RB_STATIC_SMEM_BYTES
    info->static_smem = (unsigned)(4 * (R + 1) * dwp) + RB_ZIG_WIDE_EXTRA_BYTES;
    info->uses_param_k = true;
  }

  std::ostringstream o;
  o << "// generated by rebop_b200 codegen: " << S << " species, " << R << " reactions, "
    << (macro ? "define_system! arithmetic" : "function-API arithmetic") << "\n";
  o << "#define RB_NET_STATIC_WORDS " << ((R + 1) * dwp) << "\n";  // + one all-zero row: \"no reaction\"
  o << "#define RB_TICK " << tick << "u\n";
  o << "#define RB_ZIG_WIDE\n";
  if (unroll != 1) o << "#define RB_INNER_UNROLL " << unroll << "\n";
  if (conv == 1) o << "#define RB_STATE_INT\n";
  if (const char* env = std::getenv("REBOP_B200_CODEGEN")) {  // defs=RB_A+RB_B: experimental switches of ssa_kernel.cuh
    const std::string e(env);
    const size_t pos = e.find("defs=");
    if (pos != std::string::npos) {
      std::string list = e.substr(pos + 5, e.find(',', pos) == std::string::npos ? std::string::npos : e.find(',', pos) - pos - 5);
      size_t a = 0;
      while (a < list.size()) {
        size_t b = list.find('+', a);
        if (b == std::string::npos) b = list.size();
        std::string d = list.substr(a, b - a);
        const size_t eq = d.find(':');
        if (eq != std::string::npos) d[eq] = ' ';
        if (!d.empty()) o << "#define " << d << "\n";
        a = b + 1;
      }
    }
  }
  o << "#include \"ssa_kernel.cuh\"\n\n";

  // packed stoichiometry table.  Function-API arithmetic: rows 0..R-1, then the all-zero "no reaction" row.
  // define_system! arithmetic: the all-zero row FIRST, then the reactions in descending order (row R - r): the
  // first-match chains of select() then start from the zero register and join by a max, one instruction less per pass.
  auto emit_row = [&](int r, bool first) {
    for (int w = 0; w < dwp; ++w) {
      unsigned word = 0;
      for (int l = 0; l < per_word && r >= 0; ++l) {
        const int s = w * per_word + l;
        if (s > S) continue;
        const long long d = s == S ? 1 : net.rx[r].diff[s];
        if (wide) word |= ((unsigned)(d & 0xffff)) << (16 * l);
        else word |= ((unsigned)(d & 0xff)) << (8 * l);
      }
      char buf[32];
      std::snprintf(buf, sizeof buf, "%s0x%08x", (first && w == 0) ? "" : ", ", word);
      o << buf;
    }
  };
  o << "__constant__ int rb_delta_c[" << (R + 1) * dwp << "] = {";
  if (macro) {
    emit_row(-1, true);
    for (int r = R - 1; r >= 0; --r) emit_row(r, false);
  } else {
    for (int r = 0; r < R; ++r) emit_row(r, r == 0);
    emit_row(-1, R == 0);
  }
  o << "};\n\n";

  o << "struct RbGenNet {\n";
  o << "  static constexpr int BLOCK = " << block << ";\n";
  // define_system! arithmetic: a NaN `chosen` matches no reaction (first-match chains), see rb_ssa_loop
  // (A/B against the NaN-waiting-time form: profiles/r2ad_sweep.log)
  o << "  static constexpr bool NAN_PICKS_NONE = " << (macro ? "true" : "false") << ";\n";
  o << "  rb_state x[" << (S ? S : 1) << "];  // biased-double form, see ssa_kernel.cuh\n";
  o << "  double c[" << (R ? R : 1) << "];\n";
  o << "  static __device__ __forceinline__ int smem_words(const SsaRunParams&) { return 0; }\n";
  o << "  rb_u32 tab;  // shared-window address of the packed stoichiometry rows\n";
  o << "  __device__ __forceinline__ void init(const SsaRunParams&, int*, rb_u32 tid, rb_u32 sbase) {\n";
  o << "    for (int i = (int)tid; i < " << (R + 1) * dwp << "; i += BLOCK) rb_zig.net[i] = rb_delta_c[i];\n";
  o << "    tab = sbase + RB_SMEM_OFF_NET;\n";
  o << "  }\n";
  o << "  __device__ __forceinline__ void load(const SsaRunParams& p, rb_u32 traj, bool valid) {\n";
  for (int s = 0; s < S; ++s)
    o << "    x[" << s << "] = rb_bias_pack(valid ? p.x[(size_t)" << s << " * p.ldn + traj] : 0, p.bias_hi);\n";
  o << "  }\n";
  o << "  __device__ __forceinline__ void store(const SsaRunParams& p, rb_u32 traj) {\n";
  for (int s = 0; s < S; ++s) o << "    p.x[(size_t)" << s << " * p.ldn + traj] = rb_bias_int(x[" << s << "]);\n";
  o << "  }\n";

  // propensities + cumulative sum (make_cumrates, src/gillespie.rs:357-364;
  // macro: src/gillespie_macro.rs:106-107)
  o << "  __device__ __forceinline__ double propensities(const SsaRunParams& p) {\n";
  std::vector<bool> need_d(S, false);
  for (const RbReaction& rx : net.rx) {
    if (rx.is_expr) {
      for (const rebop_expr_op& op : rx.prog)
        if (op.op == REBOP_OP_SPECIES) need_d[op.index] = true;
    } else {
      for (size_t j = 0; j < rx.term_idx.size(); ++j) need_d[rx.term_idx[j]] = true;
    }
  }
  for (int s = 0; s < S; ++s)
    if (need_d[s]) o << "    const double d" << s << " = rb_bias_f64(x[" << s << "], p);\n";
  if (R == 0) o << "    c[0] = 0.0;\n    return 0.0;\n";
  for (int r = 0; r < R; ++r) {
    const RbReaction& rx = net.rx[r];
    std::string a;
    if (rx.is_expr) {
      a = emit_expr(rx.prog);
    } else {
      a = "p.k[" + std::to_string(r) + "]";
      for (size_t j = 0; j < rx.term_idx.size(); ++j) {
        const int s = (int)rx.term_idx[j];
        const int e = (int)rx.term_exp[j];
        if (e == 1 || (macro && e == 0)) {
          a = "__dmul_rn(" + a + ", d" + std::to_string(s) + ")";
        } else if (!macro) {
          // factors (n+1-e)..=n ascending, one f64 multiply each (src/gillespie.rs:73-87);
          // n - f is exact in f64
          for (int f = e - 1; f >= 0; --f) {
            if (f == 0) a = "__dmul_rn(" + a + ", d" + std::to_string(s) + ")";
            else a = "__dmul_rn(" + a + ", __dsub_rn(d" + std::to_string(s) + ", " + std::to_string(f) + ".0))";
          }
        } else if (e == 2) {
          // n * (n - 1) as isize, then `as f64` (src/gillespie_macro.rs:133-146): one rounding of
          // the exact product, which is what the f64 multiply of the two exact factors gives too
          a = "__dmul_rn(" + a + ", __dmul_rn(d" + std::to_string(s) + ", __dsub_rn(d" + std::to_string(s) + ", 1.0)))";
        } else {
          // wrapping integer falling factorial, converted once (src/gillespie_macro.rs:133-146)
          const std::string n = "(rb_u64)(rb_i64)rb_bias_int(x[" + std::to_string(s) + "])";
          std::string prod = n;
          for (int i = 1; i < e; ++i) prod += " * (" + n + " - " + std::to_string(i) + "ull)";
          a = "__dmul_rn(" + a + ", __ll2double_rn((rb_i64)(" + prod + ")))";
        }
      }
    }
    // `0.0 + r_0` only turns -0.0 into +0.0, which no comparison downstream can see
    if (r == 0) o << "    c[0] = " << a << ";\n";
    else o << "    c[" << r << "] = __dadd_rn(c[" << r - 1 << "], " << a << ");\n";
  }
  // pin the cumulative rates: under register pressure the compiler otherwise re-computes some of them
  // (FP64 pipe work) next to the comparisons of select()
  for (int r = 0; r < R; ++r) o << "    asm volatile(\"\" : \"+d\"(c[" << r << "]));\n";
  if (R > 0) o << "    return c[" << R - 1 << "];\n";
  o << "  }\n";

  // the total as the last cumulative rate itself (not a copy: the sparse pass overwrites it on a cold path)
  o << "  __device__ __forceinline__ double& total_ref(const SsaRunParams& p) {\n    propensities(p);\n    return c["
    << (R ? R - 1 : 0) << "];\n  }\n";

  // select + update
  o << "  __device__ __forceinline__ int select(const SsaRunParams&, double chosen) const {\n";
  if (R == 0) {
    o << "    return 0;\n  }\n";
    o << "  __device__ __forceinline__ void apply(const SsaRunParams&, int, rb_u32&) {\n";
  } else {
    // The value handed from select()/none() to apply() is the byte offset of the reaction's stoichiometry row, not
    // its index: that saves the shift in front of the row's load (+0.4 %, profiles/r2t_sweep.log; knob prescale=0)
    const int rs = prescale ? 4 * dwp : 1;
    if (!macro) {
      // choose_cumrate_sum (src/gillespie.rs:402-407): index = number of cum < chosen (two
      // interleaved counters halve the dependency chain)
      o << "    int i = 0, i2 = 0;\n";
      for (int r = 0; r < R; ++r)
        o << "    rb_count_lt_by<" << rs << ">(" << ((r & 1) ? "i2" : "i") << ", c[" << r << "], chosen);\n";
      o << "    i += i2;\n";
      o << "    i = i < " << (R - 1) * rs << " ? i : " << (R - 1) * rs << ";\n";
    } else {
      // _choice! (src/gillespie_macro.rs:150-171): first r with chosen < c[r]; none => nothing happens.
      // Two half-range chains walked from the last reaction down, joined by a max (rows are stored in descending
      // order behind the all-zero row, see above), halve the dependency chain.
      const int H = R / 2;
      o << "    int i = 0, i2 = 0;\n";
      for (int r = R - 1; r >= H; --r) {
        o << "    rb_first_lt<" << (R - r) * rs << ">(i2, chosen, c[" << r << "]);\n";
        if (r - H >= 0 && r - H < H) o << "    rb_first_lt<" << (R - (r - H)) * rs << ">(i, chosen, c[" << r - H << "]);\n";
      }
      o << "    i = max(i, i2);\n";
    }
    o << "    return i;\n  }\n";
    // row R of the table is all zeros: applying it is the branch-free \"no reaction\" (macro arithmetic: nothing
    // matched, src/gillespie_macro.rs:150-171; any arithmetic: the ensemble loop's lanes without an event)
    o << "  __device__ __forceinline__ void apply(const SsaRunParams& p, int i, rb_u32& nev) {\n";
    // fetch the packed stoichiometry row of reaction i from shared memory
    const std::string row = prescale ? "(rb_u32)i" : std::to_string(4 * dwp) + "u * i";
    if (dwp == 1) {
      o << "    const int w0 = rb_lds_i32(tab + " << row << ");\n";
    } else if (dwp == 2) {
      o << "    const int2 v0 = rb_lds_i32x2(tab + " << row << ");\n";
      o << "    const int w0 = v0.x, w1 = v0.y;\n";
    } else {
      for (int q = 0; q < dwp / 4; ++q) {
        o << "    const int4 v" << q << " = rb_lds_i32x4(tab + " << row << " + " << 16 * q << "u);\n";
        o << "    const int w" << 4 * q << " = v" << q << ".x, w" << 4 * q + 1 << " = v" << q << ".y, w" << 4 * q + 2
          << " = v" << q << ".z, w" << 4 * q + 3 << " = v" << q << ".w;\n";
      }
    }
    for (int s = 0; s < S; ++s) {
      if (!touched[s]) continue;
      const int w = s / per_word, l = s % per_word;
      if (wide) o << "    rb_bias_dp2a(x[" << s << "], w" << w << ", " << l << ");\n";
      else {
        o << "    rb_bias_dp4a(x[" << s << "], w" << w << ", p.byte_sel[" << l << "]);\n";
      }
    }
    {
      const int w = S / per_word, l = S % per_word;
      if (wide) o << "    nev = (rb_u32)__dp2a_lo(w" << w << ", " << (l ? "0x100" : "0x1") << ", (int)nev);\n";
      else o << "    nev = (rb_u32)__dp4a(w" << w << ", p.byte_sel[" << l << "], (int)nev);\n";
    }
    for (int w = 0; w < dwp; ++w) o << "    (void)w" << w << ";\n";
  }
  o << "  }\n";

  o << "  static __device__ __forceinline__ int none() { return " << (macro ? 0 : R * (prescale ? 4 * dwp : 1)) << "; }\n";
  // samples: saved species in ascending index order, selected by a launch-time bit mask
  o << "  __device__ __forceinline__ void record(const SsaRunParams& p, int* dst, rb_u32 stride) const {\n";
  // every species saved (the usual case): rows at compile-time offsets, no tests of the mask
  o << "    if (p.n_save == " << S << "u) {\n";
  for (int s = 0; s < S; ++s) o << "      dst[(size_t)" << s << " * stride] = rb_bias_int(x[" << s << "]);\n";
  o << "      return;\n    }\n";
  o << "    rb_u32 row = 0;\n";
  for (int s = 0; s < S; ++s) {
    o << "    if (p.save_mask[" << s / 64 << "] & " << (1ull << (s % 64)) << "ull) { dst[(size_t)row * stride] = rb_bias_int(x[" << s
      << "]); ++row; }\n";
  }
  o << "    (void)row;\n  }\n";
  o << "};\n\n";

  emit_kernels(o, kernel_name, block, minctas, variant);
  return o.str();
}
