// ssa_pdm.cu -- K6: dependency-driven direct method for large mass-action networks.
//
// The reference recomputes every propensity at every event (src/gillespie.rs:357-364), which for a network of
// hundreds of reactions is the whole cost of a step; its Python docstring promises "heuristics" for large
// systems (python/rebop/gillespie.py:123-126) that the crate does not have (src/pyo3_gillespie.rs:161).  This kernel
// is that missing piece for the GPU: still the direct method -- the same two random numbers per event, drawn in
// the reference's order; the same distribution of waiting times and reaction choices -- but with the total
// propensity maintained incrementally through PARTIAL PROPENSITIES, so that an event costs work proportional to
// the species it changes instead of to the size of the network.
//
// It is an OPT-IN mode (REBOP_KERNEL_PDM), never chosen automatically: sums are updated in a different order than
// the reference's running sum, so the floating-point rounding differs and a trajectory eventually picks a
// different reaction than the reference would from the same random word.  Parity for this kernel is therefore
// tier 2 of the north star: ensemble distributions against the oracle (tests/test_ensemble_stats.py), not bits.
//
// Formulation (elementary mass action: orders 0, 1 and 2).  Every reaction is owned by its first reactant i:
//      a_r = x_i * k_r              (A -> ..)
//      a_r = x_i * k_r * x_j        (A + B -> ..)
//      a_r = x_i * k_r * (x_i - 1)  (2A -> ..; the reference's falling factorial, src/gillespie.rs:73-87)
// so the propensities owned by i sum to x_i * pi_i with the partial propensity pi = c + K x (K sparse: one entry
// per second-order reaction).  Zeroth-order reactions are owned by a pseudo-species whose count is the constant 1.
// Per trajectory the kernel keeps x (int32), pi (f64) and the sums of x_i * pi_i over groups of RB_PDM_SUPER species
// (f64) as columns of shared memory, and the total in a register.  An event that changes species s by d costs
//      pi_i += K_is * d, sums += x_i * K_is * d     for the few i that have s as a partner,
//      one re-evaluation of x_s * pi_s,
// the choice walks the super-group sums, the groups of one super-group and the reactions of one group, and every
// RB_PDM_REFRESH passes everything is recomputed from the counts, which bounds the rounding drift of the
// incremental sums (relative 1e-16 per update) far below anything an ensemble statistic can resolve.
#include <cuda_runtime.h>

#include <cstring>

#include "network.hpp"
#include "ssa_kernel.cuh"
#include "ssa_pdm.h"
#include "jit.hpp"

// ---------------------------------------------------------------------------
// host: lowering
// ---------------------------------------------------------------------------
int rb_pdm_build(const rebop_network& net, std::vector<uint64_t>* image, std::string* why) {
  auto no = [&](const std::string& msg) {
    if (why) *why = "the dependency-driven kernel (REBOP_KERNEL_PDM) needs " + msg;
    return REBOP_ERR_LIMIT;
  };
  const uint32_t S = net.n_species, R = (uint32_t)net.rx.size();
  const uint32_t NG = S + 1, NSG = (NG + RB_PDM_SUPER - 1) / RB_PDM_SUPER;
  if (S > 0xfffeu) return no("at most 65534 species");
  struct Col { double K; uint32_t i; };
  struct Own { double k; uint32_t r, partner; };
  std::vector<double> c(NG, 0.0);
  std::vector<std::vector<Col>> col(NG);
  std::vector<std::vector<Own>> own(NG);
  for (uint32_t r = 0; r < R; ++r) {
    const RbReaction& rx = net.rx[r];
    if (rx.is_expr) return no("mass-action rates (reaction " + std::to_string(r) + " has an expression rate)");
    if (!(rx.k >= 0.0)) return no("rate constants >= 0");
    uint32_t order = 0;
    for (uint32_t e : rx.term_exp) order += e;
    if (order > 2) return no("reactions of total order <= 2 (reaction " + std::to_string(r) + " has order " + std::to_string(order) + ")");
    int changed = 0;
    for (int64_t d : rx.diff) changed += d != 0;
    if (changed > 4) return no("at most four species changed per reaction");
    // counts must not go negative (a negative count would make partial propensities negative): every consumed species
    // has to be a reactant of at least that order, so that the propensity is 0 before the count would cross 0
    for (size_t sp = 0; sp < rx.diff.size(); ++sp) {
      if (rx.diff[sp] >= 0) continue;
      int64_t ord = 0;
      for (size_t j = 0; j < rx.term_idx.size(); ++j)
        if (rx.term_idx[j] == sp) ord += rx.term_exp[j];
      if (ord < -rx.diff[sp]) return no("reactions that consume no more of a species than their reactant order (counts could go negative)");
      if (rx.diff[sp] < -32767) return no("stoichiometric differences within int16");
    }
    // drop exponent-0 terms (they multiply by 1)
    std::vector<uint32_t> sp;
    for (size_t j = 0; j < rx.term_idx.size(); ++j)
      for (uint32_t e = 0; e < rx.term_exp[j]; ++e) sp.push_back(rx.term_idx[j]);
    if (sp.empty()) {
      c[S] += rx.k;
      own[S].push_back({rx.k, r, RB_PDM_NONE});
    } else if (sp.size() == 1) {
      c[sp[0]] += rx.k;
      own[sp[0]].push_back({rx.k, r, RB_PDM_NONE});
    } else if (sp[0] == sp[1]) {
      c[sp[0]] -= rx.k;  // k * (x - 1) = -k + k * x
      col[sp[0]].push_back({rx.k, sp[0]});
      own[sp[0]].push_back({rx.k, r, sp[0]});
    } else {
      col[sp[1]].push_back({rx.k, sp[0]});
      own[sp[0]].push_back({rx.k, r, sp[1]});
    }
  }
  // merge duplicate (s, i) entries of K (several reactions with the same reactant pair)
  for (auto& list : col) {
    std::vector<Col> merged;
    for (const Col& e : list) {
      bool found = false;
      for (Col& m : merged)
        if (m.i == e.i) { m.K += e.K; found = true; }
      if (!found) merged.push_back(e);
    }
    list.swap(merged);
  }
  size_t nnz = 0;
  for (const auto& list : col) nnz += list.size();
  RbPdmHeader h;
  std::memset(&h, 0, sizeof h);
  h.n_groups = NG;
  h.n_super = NSG;
  h.n_reactions = R;
  size_t off = (sizeof(RbPdmHeader) + 7) / 8;
  h.off_c = (uint32_t)off; off += NG;
  h.off_col_ptr = (uint32_t)off; off += (NG + 1 + 1) / 2;
  off += off & 1;  // 16-byte alignment of the entry arrays
  h.off_col = (uint32_t)off; off += 2 * nnz;
  h.off_own_ptr = (uint32_t)off; off += (NG + 1 + 1) / 2;
  off += off & 1;
  h.off_own = (uint32_t)off; off += 2 * (size_t)R;
  image->assign(off + 2, 0);
  std::memcpy(image->data(), &h, sizeof h);
  std::memcpy(image->data() + h.off_c, c.data(), NG * sizeof(double));
  uint32_t* col_ptr = reinterpret_cast<uint32_t*>(image->data() + h.off_col_ptr);
  uint32_t* own_ptr = reinterpret_cast<uint32_t*>(image->data() + h.off_own_ptr);
  uint64_t* colw = image->data() + h.off_col;
  uint64_t* ownw = image->data() + h.off_own;
  uint32_t nc = 0, no_ = 0;
  for (uint32_t g = 0; g < NG; ++g) {
    col_ptr[g] = nc;
    for (const Col& e : col[g]) {
      std::memcpy(colw + 2 * nc, &e.K, 8);
      colw[2 * nc + 1] = e.i;
      ++nc;
    }
    own_ptr[g] = no_;
    for (const Own& e : own[g]) {
      std::memcpy(ownw + 2 * no_, &e.k, 8);
      ownw[2 * no_ + 1] = (uint64_t)e.r | ((uint64_t)e.partner << 32);
      ++no_;
    }
  }
  col_ptr[NG] = nc;
  own_ptr[NG] = no_;
  return REBOP_OK;
}

// ---------------------------------------------------------------------------
// device
// ---------------------------------------------------------------------------
struct RbPdmNet {
  static constexpr int BLOCK = RB_PDM_BLOCK;
  double* pi;   // this thread's columns: pi[i * BLOCK]
  double* ss;   // super-group sums: ss[g * BLOCK]
  int* xs;      // counts: xs[i * BLOCK]; xs[S * BLOCK] is the constant 1 of the zeroth-order group
  const double* __restrict__ c;
  const rb_u32* __restrict__ col_ptr;
  const double2* __restrict__ col;
  const rb_u32* __restrict__ own_ptr;
  const double2* __restrict__ own;
  int ng, nsg, n_reactions;
  double total, scale;
  rb_u32 age;  // passes since the sums were rebuilt from the counts

  static __device__ __forceinline__ int smem_words(const SsaRunParams& p) {
    const int ng = p.n_species + 1, nsg = (ng + RB_PDM_SUPER - 1) / RB_PDM_SUPER;
    return BLOCK * (2 * ng + 2 * nsg + ng);
  }
  __device__ __forceinline__ void init(const SsaRunParams& p, int* smem, rb_u32 tid, rb_u32) {
    const rb_u64* img = static_cast<const rb_u64*>(p.pdm);
    const RbPdmHeader* h = reinterpret_cast<const RbPdmHeader*>(img);
    ng = (int)h->n_groups;
    nsg = (int)h->n_super;
    n_reactions = (int)h->n_reactions;
    c = reinterpret_cast<const double*>(img + h->off_c);
    col_ptr = reinterpret_cast<const rb_u32*>(img + h->off_col_ptr);
    col = reinterpret_cast<const double2*>(img + h->off_col);
    own_ptr = reinterpret_cast<const rb_u32*>(img + h->off_own_ptr);
    own = reinterpret_cast<const double2*>(img + h->off_own);
    double* d = reinterpret_cast<double*>(smem);
    pi = d + tid;
    ss = d + (size_t)ng * BLOCK + tid;
    xs = smem + 2 * (ng + nsg) * BLOCK + tid;
    total = scale = 0.0;
    age = 0;
  }

  // Everything from the counts: pi = c + K x, the group sums and the total.  Loop bounds and table addresses are
  // the same for every lane (they depend on the network only).
  __device__ __noinline__ void rebuild() {
    for (int i = 0; i < ng; ++i) pi[i * BLOCK] = __ldg(c + i);
    for (int s = 0; s < ng; ++s) {
      const rb_u32 e1 = __ldg(col_ptr + s + 1);
      const double x = (double)xs[s * BLOCK];
      for (rb_u32 e = __ldg(col_ptr + s); e < e1; ++e) {
        const double2 w = __ldg(col + e);
        const int i = (int)(rb_u32)__double_as_longlong(w.y);
        pi[i * BLOCK] += w.x * x;
      }
    }
    total = 0.0;
    for (int g = 0; g < nsg; ++g) {
      double sum = 0.0;
      for (int i = g * RB_PDM_SUPER; i < (g + 1) * RB_PDM_SUPER && i < ng; ++i) sum += (double)xs[i * BLOCK] * pi[i * BLOCK];
      ss[g * BLOCK] = sum;
      total += sum;
    }
    scale = total;
    age = 0;
  }

  __device__ __forceinline__ void load(const SsaRunParams& p, rb_u32 traj, bool valid) {
    const int S = ng - 1;
    for (int s = 0; s < S; ++s) xs[s * BLOCK] = valid ? p.x[(size_t)s * p.ldn + traj] : 0;
    xs[S * BLOCK] = valid ? 1 : 0;
    rebuild();
  }
  __device__ __forceinline__ void store(const SsaRunParams& p, rb_u32 traj) {
    const int S = ng - 1;
    for (int s = 0; s < S; ++s) p.x[(size_t)s * p.ldn + traj] = xs[s * BLOCK];
  }

  __device__ __forceinline__ double propensities(const SsaRunParams&) {
    // Periodic rebuild, taken by all lanes of the (converged part of the) warp together; a lane whose total has
    // collapsed by many orders of magnitude since its last rebuild (the state is probably absorbing and what is
    // left is rounding residue) asks for one at once.
    ++age;
    const bool stale = age >= RB_PDM_REFRESH || (total < scale * 0x1.0p-30 && age > 1u);
    if (__any_sync(__activemask(), stale)) rebuild();
    return total;
  }

  // The reaction whose interval of the cumulative propensity contains `chosen`: super-group, group, reaction.
  __device__ __forceinline__ int select(const SsaRunParams&, double chosen) const {
    double base = 0.0;
    int sg = 0;
    for (int g = 0; g + 1 < nsg; ++g) {  // the last super-group needs no test
      const double next = base + ss[g * BLOCK];
      if (sg == g && !(chosen < next)) {
        base = next;
        sg = g + 1;
      }
    }
    int grp = sg * RB_PDM_SUPER;
    const int last = min(ng, grp + RB_PDM_SUPER) - 1;
#pragma unroll
    for (int q = 0; q < RB_PDM_SUPER - 1; ++q) {
      const int i = sg * RB_PDM_SUPER + q;
      if (i < last && grp == i) {
        const double next = base + (double)xs[i * BLOCK] * pi[i * BLOCK];
        if (!(chosen < next)) {
          base = next;
          grp = i + 1;
        }
      }
    }
    const double xi = (double)xs[grp * BLOCK];
    const rb_u32 e1 = __ldg(own_ptr + grp + 1);
    int pick = n_reactions;  // nothing matches (rounding residue of the sums): no event this pass
    for (rb_u32 e = __ldg(own_ptr + grp); e < e1; ++e) {
      const double2 w = __ldg(own + e);
      const rb_u64 meta = (rb_u64)__double_as_longlong(w.y);
      const rb_u32 partner = (rb_u32)(meta >> 32);
      double a = xi * w.x;
      if (partner != RB_PDM_NONE) a *= partner == (rb_u32)grp ? xi - 1.0 : (double)xs[partner * BLOCK];
      base += a;
      if (chosen < base) {
        pick = (int)(rb_u32)meta;
        break;
      }
    }
    return pick;
  }

  __device__ __forceinline__ int none() const { return n_reactions; }

  __device__ __forceinline__ void apply(const SsaRunParams& p, int pick, rb_u32& nev) {
    if (pick >= n_reactions) return;
    ++nev;
    const uint4 j = __ldg(reinterpret_cast<const uint4*>(p.gtab) + 2 * pick + 1);
    const rb_u32 idx[4] = {j.x & 0xffffu, j.x >> 16, j.y & 0xffffu, j.y >> 16};
    const int diff[4] = {(int)(short)(j.z & 0xffffu), (int)(short)(j.z >> 16), (int)(short)(j.w & 0xffffu),
                         (int)(short)(j.w >> 16)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (diff[q] == 0) continue;
      const int s = (int)idx[q];
      const double d = (double)diff[q];
      const int x_old = xs[s * BLOCK];
      double pi_s = pi[s * BLOCK];
      const double g_old = (double)x_old * pi_s;
      xs[s * BLOCK] = x_old + diff[q];
      const rb_u32 e1 = __ldg(col_ptr + s + 1);
      for (rb_u32 e = __ldg(col_ptr + s); e < e1; ++e) {
        const double2 w = __ldg(col + e);
        const int i = (int)(rb_u32)__double_as_longlong(w.y);
        const double kd = w.x * d;
        if (i == s) {
          pi_s += kd;  // 2A: the species is its own partner
        } else {
          pi[i * BLOCK] += kd;
          const double delta = (double)xs[i * BLOCK] * kd;
          ss[(i / RB_PDM_SUPER) * BLOCK] += delta;
          total += delta;
        }
      }
      pi[s * BLOCK] = pi_s;
      const double delta = (double)(x_old + diff[q]) * pi_s - g_old;
      ss[(s / RB_PDM_SUPER) * BLOCK] += delta;
      total += delta;
    }
  }

  __device__ __forceinline__ void record(const SsaRunParams& p, int* dst, rb_u32 stride) const {
    const rb_u32* save = p.gtab + n_reactions * RB_GTAB_WORDS_PER_REACTION;
    for (rb_u32 j = 0; j < p.n_save; ++j) dst[(size_t)j * stride] = xs[__ldg(save + j) * BLOCK];
  }
};

__global__ void __launch_bounds__(RB_PDM_BLOCK) rb_ssa_pdm_kernel(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbPdmNet net;
  rb_ssa_loop<RbPdmNet, RB_MODE_STATIC>(net, p, rb_smem);
}
__global__ void __launch_bounds__(RB_PDM_BLOCK) rb_ssa_pdm_kernel_dyn(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbPdmNet net;
  rb_ssa_loop<RbPdmNet, RB_MODE_SPARSE>(net, p, rb_smem);
}
__global__ void __launch_bounds__(RB_PDM_BLOCK) rb_ssa_pdm_kernel_dns(const __grid_constant__ SsaRunParams p) {
  extern __shared__ __align__(16) int rb_smem[];
  RbPdmNet net;
  rb_ssa_loop<RbPdmNet, RB_MODE_DENSE>(net, p, rb_smem);
}

typedef void (*RbPdmKernel)(const SsaRunParams);
static RbPdmKernel pdm_kernel(int mode) {
  return mode == RB_MODE_STATIC ? rb_ssa_pdm_kernel : mode == RB_MODE_SPARSE ? rb_ssa_pdm_kernel_dyn : rb_ssa_pdm_kernel_dns;
}

cudaError_t rb_pdm_occupancy(int mode, size_t smem_bytes, int* ctas_per_sm) {
  auto kernel = pdm_kernel(mode);
  cudaError_t err = rb_raise_smem_limit(reinterpret_cast<const void*>(kernel), smem_bytes);
  if (err != cudaSuccess) return err;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kernel, RB_PDM_BLOCK, smem_bytes);
}

cudaError_t rb_pdm_launch(int mode, const SsaRunParams& p, unsigned grid, size_t smem_bytes, cudaStream_t stream) {
  auto kernel = pdm_kernel(mode);
  cudaError_t err = rb_raise_smem_limit(reinterpret_cast<const void*>(kernel), smem_bytes);
  if (err != cudaSuccess) return err;
  kernel<<<grid, RB_PDM_BLOCK, smem_bytes, stream>>>(p);
  return cudaGetLastError();
}
