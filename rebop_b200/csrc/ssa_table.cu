// ssa_table.cu -- K1: table-driven direct-method kernel (any network, no JIT).
//
// Species counts live in shared memory (one column per thread, conflict-free).  The network is
// data: one 16-byte record per reaction for the common mass-action case (SsaRunParams::gtab), and
// the batch's own RbTables image in global memory (SsaRunParams::tables: CSR reactant terms,
// stoichiometry, expression byte-code) for everything else.  Both are read at warp-uniform
// addresses (every lane evaluates reaction r at the same time), so each fetch is a broadcast out
// of L1.  Nothing is process-global: batches with different networks may run side by side on one
// device from different host threads and streams.
//
// Replaces: Gillespie::advance_until + the pyo3 grid loop
// (src/gillespie.rs:315-344, src/pyo3_gillespie.rs:197-208) with Rate::rate
// (src/gillespie.rs:71-90), Expr::eval (src/expr.rs:24-38) and Jump::affect
// (src/gillespie.rs:142-152) for LMASparse/Sparse forms, which the reference
// pins as bit-identical to the dense forms (tests/test_rebop.py:55-65).
#include <cuda_runtime.h>

#include "rb_tables.h"
#include "ssa_kernel.cuh"
#include "ssa_table.h"
#include "jit.hpp"

struct RbTableNet {
  static constexpr int BLOCK = RB_TABLE_BLOCK;
  static constexpr bool NAN_PICKS_NONE = false;
  int* xs;  // this thread's column: species s at xs[s * BLOCK]
  const RbTables* __restrict__ tab;
  int n_reactions, arith;

  static __device__ __forceinline__ int smem_words(const SsaRunParams& p) { return p.n_species * BLOCK; }
  __device__ __forceinline__ void init(const SsaRunParams& p, int* smem, rb_u32 tid, rb_u32) {
    xs = smem + tid;
    tab = static_cast<const RbTables*>(p.tables);
    n_reactions = p.n_reactions;
    arith = p.arith;
  }
  __device__ __forceinline__ void load(const SsaRunParams& p, rb_u32 traj, bool valid) {
    const int S = p.n_species;
    for (int s = 0; s < S; ++s) xs[s * BLOCK] = valid ? p.x[(size_t)s * p.ldn + traj] : 0;
  }
  __device__ __forceinline__ void store(const SsaRunParams& p, rb_u32 traj) {
    const int S = p.n_species;
    for (int s = 0; s < S; ++s) p.x[(size_t)s * p.ldn + traj] = xs[s * BLOCK];
  }

  // Expr::eval on the post-order program of reaction r.
  __device__ __noinline__ double eval_expr(int lo, int hi) const {
    double stack[RB_EXPR_STACK];
    int sp = 0;
    for (int i = lo; i < hi; ++i) {
      const int op = __ldg(&tab->op_code[i]);
      if (op == RB_OP_CONST) {
        stack[sp++] = __ldg(&tab->op_val[i]);
      } else if (op == RB_OP_SPECIES) {
        stack[sp++] = rb_i2d(xs[__ldg(&tab->op_idx[i]) * BLOCK]);
      } else if (op == RB_OP_NEG) {
        stack[sp - 1] = -stack[sp - 1];
      } else if (op == RB_OP_EXP) {
        stack[sp - 1] = exp(stack[sp - 1]);
      } else {
        const double b = stack[--sp], a = stack[sp - 1];
        double v;
        switch (op) {
          case RB_OP_ADD: v = __dadd_rn(a, b); break;
          case RB_OP_SUB: v = __dsub_rn(a, b); break;
          case RB_OP_MUL: v = __dmul_rn(a, b); break;
          case RB_OP_DIV: v = __ddiv_rn(a, b); break;
          case RB_OP_POW: v = pow(a, b); break;
          case RB_OP_MAX: v = fmax(a, b); break;
          case RB_OP_MIN: v = fmin(a, b); break;
          default: v = __longlong_as_double(0x7ff8000000000000ll);
        }
        stack[sp - 1] = v;
      }
    }
    return stack[0];
  }

  // One reactant term: acc * n(n-1)...(n-e+1) in the arithmetic of the selected engine.
  static __device__ __forceinline__ double term(double acc, int n, int e, int arith) {
    if (e == 1) return __dmul_rn(acc, rb_i2d(n));
    if (arith == 0) {
      // src/gillespie.rs:73-87: factors (n+1-e)..=n ascending, one f64 multiply each
      for (int f = n + 1 - e; f <= n; ++f) acc = __dmul_rn(acc, rb_i2d(f));
      return acc;
    }
    // src/gillespie_macro.rs:133-146: wrapping integer falling factorial, one conversion
    rb_u64 prod = (rb_u64)(rb_i64)n;
    for (int i = 1; i < e; ++i) prod *= (rb_u64)(rb_i64)(n - i);
    return __dmul_rn(acc, __ll2double_rn((rb_i64)prod));
  }

  // Rate::rate for reaction r through the CSR tables in constant memory (any number of terms, expressions).
  __device__ __noinline__ double rate_general(int r) const {
    const int e0 = __ldg(&tab->expr_ptr[r]), e1 = __ldg(&tab->expr_ptr[r + 1]);
    if (e1 > e0) return eval_expr(e0, e1);
    double acc = __ldg(&tab->k[r]);
    const int j1 = __ldg(&tab->term_ptr[r + 1]);
    for (int j = __ldg(&tab->term_ptr[r]); j < j1; ++j)
      acc = term(acc, xs[__ldg(&tab->term_idx[j]) * BLOCK], __ldg(&tab->term_exp[j]), arith);
    return acc;
  }

  // Rate::rate for reaction r.  The common case -- mass action with at most two reactant terms -- is
  // served by ONE 16-byte record (SsaRunParams::gtab, see ssa_params.h) read at a warp-uniform address,
  // so consecutive reactions do not wait on chains of dependent table loads.
  __device__ __forceinline__ double rate(const uint4* __restrict__ rec, int r) const {
    const uint4 w = __ldg(rec + 2 * r);
    const rb_u32 n = (w.w >> 16) & 0xffu;
    if (n == 0xffu) return rate_general(r);
    double acc = __hiloint2double((int)w.y, (int)w.x);
    if (n >= 1u) acc = term(acc, xs[(w.z & 0xffffu) * BLOCK], (int)(w.w & 0xffu), arith);
    if (n >= 2u) acc = term(acc, xs[(w.z >> 16) * BLOCK], (int)((w.w >> 8) & 0xffu), arith);
    return acc;
  }

  // make_cumrates (src/gillespie.rs:357-364); only the total is kept, select() re-walks the sum.
  double tot;
  __device__ __forceinline__ double& total_ref(const SsaRunParams& p) {
    tot = propensities(p);
    return tot;
  }
  __device__ __forceinline__ double propensities(const SsaRunParams& p) const {
    const int R = n_reactions;
    const uint4* rec = reinterpret_cast<const uint4*>(p.gtab);
    double total = 0.0;
#pragma unroll 4
    for (int r = 0; r < R; ++r) total = __dadd_rn(total, rate(rec, r));
    return total;
  }

  __device__ __forceinline__ int select(const SsaRunParams& p, double chosen) const {
    const int R = n_reactions;
    const uint4* rec = reinterpret_cast<const uint4*>(p.gtab);
    double cum = 0.0;
    int i;
    if (arith == 0) {
      // choose_cumrate_sum (src/gillespie.rs:402-407): the index is a count
      i = 0;
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        cum = __dadd_rn(cum, rate(rec, r));
        i += (cum < chosen) ? 1 : 0;
      }
      if (i >= R) i = R - 1;  // unreachable for finite totals (src/gillespie.rs:339)
      if (R == 0) i = 0;
    } else {
      // _choice! (src/gillespie_macro.rs:150-171): first r with chosen < carry + r_r
      i = R;
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        cum = __dadd_rn(cum, rate(rec, r));
        if (i == R && chosen < cum) i = r;
      }
    }
    return i;  // R: nothing matches (macro arithmetic), or no reactions at all
  }

  __device__ __forceinline__ int none() const { return n_reactions; }

  __device__ __forceinline__ void apply(const SsaRunParams& p, int i, rb_u32& nev) {
    if (i >= n_reactions) return;
    ++nev;
    const uint4* rec = reinterpret_cast<const uint4*>(p.gtab);
    const uint4 j = __ldg(rec + 2 * i + 1);
    if ((__ldg(rec + 2 * i).w >> 24) == 0u) {
      // at most four species change: (species, difference) pairs from the record
      const rb_u32 idx[4] = {j.x & 0xffffu, j.x >> 16, j.y & 0xffffu, j.y >> 16};
      const int diff[4] = {(int)(short)(j.z & 0xffffu), (int)(short)(j.z >> 16), (int)(short)(j.w & 0xffffu),
                           (int)(short)(j.w >> 16)};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (diff[q] != 0) xs[idx[q] * BLOCK] += diff[q];
    } else {
      const int j1 = __ldg(&tab->jump_ptr[i + 1]);
      for (int q = __ldg(&tab->jump_ptr[i]); q < j1; ++q) xs[__ldg(&tab->jump_idx[q]) * BLOCK] += __ldg(&tab->jump_diff[q]);
    }
  }

  __device__ __forceinline__ void record(const SsaRunParams& p, int* dst, rb_u32 stride) const {
    const rb_u32* save = p.gtab + n_reactions * RB_GTAB_WORDS_PER_REACTION;  // saved-species list behind the records
    for (rb_u32 j = 0; j < p.n_save; ++j) dst[(size_t)j * stride] = xs[__ldg(save + j) * BLOCK];
  }
};

__global__ void __launch_bounds__(RB_TABLE_BLOCK) rb_ssa_table_kernel(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbTableNet net;
  rb_ssa_loop<RbTableNet, RB_MODE_STATIC>(net, p, rb_smem);
}
__global__ void __launch_bounds__(RB_TABLE_BLOCK) rb_ssa_table_kernel_dyn(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbTableNet net;
  rb_ssa_loop<RbTableNet, RB_MODE_SPARSE>(net, p, rb_smem);
}
__global__ void __launch_bounds__(RB_TABLE_BLOCK) rb_ssa_table_kernel_dns(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbTableNet net;
  rb_ssa_loop<RbTableNet, RB_MODE_DENSE>(net, p, rb_smem);
}

__global__ void __launch_bounds__(RB_TABLE_BLOCK) rb_ssa_table_kernel_evc(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbTableNet net;
  rb_ssa_events<RbTableNet, false>(net, p, rb_smem);
}
__global__ void __launch_bounds__(RB_TABLE_BLOCK) rb_ssa_table_kernel_evw(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbTableNet net;
  rb_ssa_events<RbTableNet, true>(net, p, rb_smem);
}

typedef void (*RbTableKernel)(const SsaRunParams);
static RbTableKernel grid_kernel(int mode) {
  return mode == RB_MODE_STATIC ? rb_ssa_table_kernel : mode == RB_MODE_SPARSE ? rb_ssa_table_kernel_dyn : rb_ssa_table_kernel_dns;
}

// Event-log mode: counting (write = false) or writing pass.
cudaError_t rb_table_launch_events(bool write, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream) {
  auto kernel = write ? rb_ssa_table_kernel_evw : rb_ssa_table_kernel_evc;
  cudaError_t err = rb_raise_smem_limit(reinterpret_cast<const void*>(kernel), smem_bytes);
  if (err != cudaSuccess) return err;
  kernel<<<grid, RB_TABLE_BLOCK, smem_bytes, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t rb_table_occupancy(int mode, size_t smem_bytes, int* ctas_per_sm) {
  auto kernel = grid_kernel(mode);
  cudaError_t err = rb_raise_smem_limit(reinterpret_cast<const void*>(kernel), smem_bytes);
  if (err != cudaSuccess) return err;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kernel, RB_TABLE_BLOCK, smem_bytes);
}

cudaError_t rb_table_launch(int mode, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream) {
  auto kernel = grid_kernel(mode);
  cudaError_t err = rb_raise_smem_limit(reinterpret_cast<const void*>(kernel), smem_bytes);
  if (err != cudaSuccess) return err;
  kernel<<<grid, RB_TABLE_BLOCK, smem_bytes, stream>>>(p);
  return cudaGetLastError();
}
