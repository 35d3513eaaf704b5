// codegen.cpp -- network -> CUDA source of a network-specialised ensemble kernel (K2).
//
// This is the GPU analogue of what `define_system!` does at the user's compile time
// (src/gillespie_macro.rs:49-129): species become scalars (registers), every propensity is one
// straight-line expression, the cumulative sum and the reaction choice are fully unrolled.
// The same generator serves the run-time path (NVRTC, jit.cpp) and the build-time path
// (tools/rebop_sysgen -> nvcc).  It has no CUDA dependency.
#include "codegen.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
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
      st.push_back("rb_i2d(x[" + std::to_string(op.index) + "])");
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

// Cumulative rates are non-decreasing for every reachable state when all reactions are mass-action
// with a non-negative constant and no reaction can drive a count negative (every species a reaction
// consumes is a reactant of at least that order, so the propensity is 0 before the count would go
// below 0).  The initial state must be non-negative too, which the engine checks per batch.
bool network_is_monotone(const rebop_network& net) {
  for (const RbReaction& rx : net.rx) {
    if (rx.is_expr) return false;
    if (!(rx.k >= 0.0)) return false;  // also rejects NaN
    for (size_t s = 0; s < rx.diff.size(); ++s) {
      if (rx.diff[s] >= 0) continue;
      int64_t order = 0;
      for (size_t j = 0; j < rx.term_idx.size(); ++j)
        if (rx.term_idx[j] == s) order += rx.term_exp[j];
      if (order < -rx.diff[s]) return false;
    }
  }
  return true;
}

// Branch-free binary search over the register-resident cumulative rates c[0..R-2] (c[R-1] and the
// padding up to a power of two count as +inf, which is what clamping the index to R-1 means).
// `probe(j)` is the source text of the predicate "move right past c[j]".
struct BinarySearchEmitter {
  std::ostringstream& o;
  int R;
  std::string (*probe)(const std::string& value);
  // emits code selecting c[lo_index(bits)] for the current level; returns the expression
  std::string select(int level, int prefix, int depth, int half) const {
    // candidates at this level: prefix (already decided high bits) -> index prefix + half - 1
    if (depth == level) {
      const int j = prefix + half - 1;
      return j <= R - 2 ? "c[" + std::to_string(j) + "]" : std::string();
    }
    const int bit = half << (level - depth);  // weight of decision `depth`
    const std::string hi = select(level, prefix + bit, depth + 1, half);
    const std::string lo = select(level, prefix, depth + 1, half);
    if (hi.empty()) return lo.empty() ? std::string() : "(q" + std::to_string(depth) + " ? RB_INF : " + lo + ")";
    return "(q" + std::to_string(depth) + " ? " + hi + " : " + lo + ")";
  }
};

}  // namespace

bool rb_codegen_monotone(const rebop_network& net) { return network_is_monotone(net); }

RbCodegenOptions rb_codegen_default_options() {
  RbCodegenOptions opt;
  if (const char* env = std::getenv("REBOP_B200_CODEGEN")) {
    const std::string e(env);
    auto get = [&](const char* key, int def) {
      const size_t pos = e.find(std::string(key) + "=");
      return pos == std::string::npos ? def : std::atoi(e.c_str() + pos + std::strlen(key) + 1);
    };
    opt.conv = get("conv", opt.conv);
    opt.select = get("select", opt.select);
    opt.bake = get("bake", opt.bake);
    opt.block = get("block", opt.block);
    opt.min_ctas = get("minctas", opt.min_ctas);
  }
  return opt;
}

bool rb_codegen_supported(const rebop_network& net, std::string* why) {
  if (net.n_species > RB_GEN_MAX_SPECIES) {
    if (why) *why = "more than " + std::to_string(RB_GEN_MAX_SPECIES) + " species (register-resident state)";
    return false;
  }
  if (net.rx.size() > RB_GEN_MAX_REACTIONS) {
    if (why) *why = "more than " + std::to_string(RB_GEN_MAX_REACTIONS) + " reactions (register-resident cumulative sums)";
    return false;
  }
  return true;
}

std::string rb_codegen_source(const rebop_network& net, const std::string& kernel_name, RbCodegenInfo* info) {
  const int S = (int)net.n_species;
  const int R = (int)net.rx.size();
  const bool macro = net.arith == REBOP_ARITH_MACRO;

  // stoichiometry packing: int8 lanes (dp4a) unless some |difference| needs int16 (dp2a)
  bool wide = false;
  for (const RbReaction& rx : net.rx)
    for (int64_t d : rx.diff)
      if (d < -127 || d > 127) wide = true;
  const int per_word = wide ? 2 : 4;
  int dw = (S + per_word - 1) / per_word;
  if (dw == 0) dw = 1;
  int dwp = dw <= 2 ? dw : (dw + 3) / 4 * 4;  // 1, 2 or a multiple of 4 words per reaction
  std::vector<bool> touched(S, false);
  for (const RbReaction& rx : net.rx)
    for (int s = 0; s < S; ++s)
      if (rx.diff[s] != 0) touched[s] = true;

  const unsigned block = 128;
  if (info) {
    info->block = block;
    info->net_words = (unsigned)(R * dwp);
    info->uses_param_k = true;
  }

  std::ostringstream o;
  o << "// generated by rebop_b200 codegen: " << S << " species, " << R << " reactions, "
    << (macro ? "define_system! arithmetic" : "function-API arithmetic") << "\n";
  o << "#include \"ssa_kernel.cuh\"\n\n";

  // packed stoichiometry table
  o << "__constant__ int rb_delta_c[" << (R * dwp > 0 ? R * dwp : 1) << "] = {";
  for (int r = 0; r < R; ++r) {
    for (int w = 0; w < dwp; ++w) {
      unsigned word = 0;
      for (int l = 0; l < per_word; ++l) {
        const int s = w * per_word + l;
        if (s >= S) continue;
        const long long d = net.rx[r].diff[s];
        if (wide) word |= ((unsigned)(d & 0xffff)) << (16 * l);
        else word |= ((unsigned)(d & 0xff)) << (8 * l);
      }
      char buf[32];
      std::snprintf(buf, sizeof buf, "%s0x%08x", (r || w) ? ", " : "", word);
      o << buf;
    }
  }
  if (R * dwp == 0) o << "0";
  o << "};\n\n";

  o << "struct RbGenNet {\n";
  o << "  static constexpr int BLOCK = " << block << ";\n";
  o << "  int x[" << (S ? S : 1) << "];\n";
  o << "  double c[" << (R ? R : 1) << "];\n";
  o << "  const int* tab;\n";
  o << "  static __device__ __forceinline__ int smem_words(const SsaRunParams&) { return " << R * dwp << "; }\n";
  o << "  __device__ __forceinline__ void init(const SsaRunParams&, int* smem, rb_u32 tid) {\n";
  o << "    for (int i = (int)tid; i < " << R * dwp << "; i += BLOCK) smem[i] = rb_delta_c[i];\n";
  o << "    tab = smem;\n  }\n";
  o << "  __device__ __forceinline__ void load(const SsaRunParams& p, rb_u32 traj, bool valid) {\n";
  for (int s = 0; s < S; ++s)
    o << "    x[" << s << "] = valid ? p.x[(size_t)" << s << " * p.ldn + traj] : 0;\n";
  o << "  }\n";
  o << "  __device__ __forceinline__ void store(const SsaRunParams& p, rb_u32 traj) {\n";
  for (int s = 0; s < S; ++s) o << "    p.x[(size_t)" << s << " * p.ldn + traj] = x[" << s << "];\n";
  o << "  }\n";

  // propensities + cumulative sum (make_cumrates, src/gillespie.rs:357-364;
  // macro: src/gillespie_macro.rs:106-107)
  o << "  __device__ __forceinline__ double propensities(const SsaRunParams& p) {\n";
  std::vector<bool> need_d(S, false);
  for (const RbReaction& rx : net.rx)
    if (!rx.is_expr)
      for (size_t j = 0; j < rx.term_idx.size(); ++j)
        if (rx.term_exp[j] == 1 || (macro && rx.term_exp[j] == 0)) need_d[rx.term_idx[j]] = true;
  for (int s = 0; s < S; ++s)
    if (need_d[s]) o << "    const double d" << s << " = rb_i2d(x[" << s << "]);\n";
  if (R == 0) o << "    return 0.0;\n";
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
          // factors (n+1-e)..=n ascending, one f64 multiply each (src/gillespie.rs:73-87)
          for (int f = e - 1; f >= 0; --f)
            a = "__dmul_rn(" + a + ", rb_i2d(x[" + std::to_string(s) + "] - " + std::to_string(f) + "))";
        } else {
          // wrapping integer falling factorial, converted once (src/gillespie_macro.rs:133-146)
          std::string prod = "(rb_u64)(rb_i64)x[" + std::to_string(s) + "]";
          for (int i = 1; i < e; ++i)
            prod += " * (rb_u64)(rb_i64)(x[" + std::to_string(s) + "] - " + std::to_string(i) + ")";
          a = "__dmul_rn(" + a + ", __ll2double_rn((rb_i64)(" + prod + ")))";
        }
      }
    }
    // `0.0 + r_0` only turns -0.0 into +0.0, which no comparison downstream can see
    if (r == 0) o << "    c[0] = " << a << ";\n";
    else o << "    c[" << r << "] = __dadd_rn(c[" << r - 1 << "], " << a << ");\n";
  }
  if (R > 0) o << "    return c[" << R - 1 << "];\n";
  o << "  }\n";

  // select + update
  o << "  __device__ __forceinline__ bool fire(const SsaRunParams&, double chosen) {\n";
  if (R == 0) {
    o << "    return false;\n";
  } else {
    if (!macro) {
      // choose_cumrate_sum (src/gillespie.rs:402-407): index = number of cum < chosen
      o << "    int i = 0;\n";
      for (int r = 0; r < R; ++r) o << "    i += (c[" << r << "] < chosen) ? 1 : 0;\n";
      o << "    i = i < " << R - 1 << " ? i : " << R - 1 << ";\n";
    } else {
      // _choice! (src/gillespie_macro.rs:150-171): first r with chosen < c[r]; none => nothing happens
      o << "    int i = " << R << ";\n";
      for (int r = R - 1; r >= 0; --r) o << "    i = (chosen < c[" << r << "]) ? " << r << " : i;\n";
      o << "    if (i == " << R << ") return false;\n";
    }
    // fetch the packed stoichiometry row of reaction i from shared memory
    if (dwp == 1) {
      o << "    const int w0 = tab[i];\n";
    } else if (dwp == 2) {
      o << "    const int2 v0 = *reinterpret_cast<const int2*>(tab + 2 * i);\n";
      o << "    const int w0 = v0.x, w1 = v0.y;\n";
    } else {
      for (int q = 0; q < dwp / 4; ++q) {
        o << "    const int4 v" << q << " = *reinterpret_cast<const int4*>(tab + " << dwp << " * i + " << 4 * q << ");\n";
        o << "    const int w" << 4 * q << " = v" << q << ".x, w" << 4 * q + 1 << " = v" << q << ".y, w" << 4 * q + 2
          << " = v" << q << ".z, w" << 4 * q + 3 << " = v" << q << ".w;\n";
      }
    }
    for (int s = 0; s < S; ++s) {
      if (!touched[s]) continue;
      const int w = s / per_word, l = s % per_word;
      if (wide) o << "    x[" << s << "] = __dp2a_lo(w" << w << ", " << (l ? "0x100" : "0x1") << ", x[" << s << "]);\n";
      else {
        char sel[16];
        std::snprintf(sel, sizeof sel, "0x%x", 1u << (8 * l));
        o << "    x[" << s << "] = __dp4a(w" << w << ", " << sel << ", x[" << s << "]);\n";
      }
    }
    o << "    return true;\n";
  }
  o << "  }\n";

  // samples: saved species in ascending index order, selected by a launch-time bit mask
  o << "  __device__ __forceinline__ void record(const SsaRunParams& p, int* dst, rb_u32 stride) const {\n";
  o << "    rb_u32 row = 0;\n";
  for (int s = 0; s < S; ++s) {
    o << "    if (p.save_mask[" << s / 64 << "] & " << (1ull << (s % 64)) << "ull) { dst[(size_t)row * stride] = x[" << s
      << "]; ++row; }\n";
  }
  o << "    (void)row;\n  }\n";
  o << "};\n\n";

  o << "extern \"C\" __global__ void __launch_bounds__(" << block << ") " << kernel_name
    << "(const __grid_constant__ SsaRunParams p) {\n";
  o << "  extern __shared__ __align__(16) int rb_smem[];\n";
  o << "  RbGenNet net;\n";
  o << "  rb_ssa_loop(net, p, rb_smem);\n";
  o << "}\n";
  return o.str();
}
