// pdm.cpp -- lowering of a mass-action network into its partial-propensity form (pdm.hpp).
#include "pdm.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "ssa_params.h"

int rb_pdm_lower(const rebop_network& net, RbPdmLowered* out, std::string* why) {
  auto no = [&](const std::string& msg) {
    if (why) *why = "the partial-propensity kernel (REBOP_KERNEL_PDM) needs " + msg;
    return REBOP_ERR_LIMIT;
  };
  const uint32_t S = net.n_species, R = (uint32_t)net.rx.size();
  if (S > 0xfffeu) return no("at most 65534 species");
  if (R >= (1u << 30)) return no("fewer than 2^30 reactions");
  std::vector<double> c(S + 1, 0.0);
  std::vector<std::vector<std::pair<uint32_t, double>>> row(S + 1);  // (j, K_ij)
  std::vector<std::vector<RbPdmGroup::Own>> own(S + 1);
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
    // counts must not go negative: every consumed species has to be a reactant of at least that order, so that the
    // propensity is 0 before the count would cross 0
    for (size_t sp = 0; sp < rx.diff.size(); ++sp) {
      if (rx.diff[sp] >= 0) continue;
      int64_t ord = 0;
      for (size_t j = 0; j < rx.term_idx.size(); ++j)
        if (rx.term_idx[j] == sp) ord += rx.term_exp[j];
      if (ord < -rx.diff[sp]) return no("reactions that consume no more of a species than their reactant order (counts could go negative)");
    }
    for (int64_t d : rx.diff)
      if (d < -32767 || d > 32767) return no("stoichiometric differences within int16");
    std::vector<uint32_t> sp;  // reactant species, one entry per unit of order
    for (size_t j = 0; j < rx.term_idx.size(); ++j)
      for (uint32_t e = 0; e < rx.term_exp[j]; ++e) sp.push_back(rx.term_idx[j]);
    if (sp.empty()) {
      c[S] += rx.k;
      own[S].push_back({rx.k, r, RB_PDM_NONE});
    } else if (sp.size() == 1) {
      c[sp[0]] += rx.k;
      own[sp[0]].push_back({rx.k, r, RB_PDM_NONE});
    } else {
      const uint32_t i = sp[0], j = sp[1];
      if (i == j) c[i] -= rx.k;  // k (x - 1) = -k + k x
      bool merged = false;
      for (auto& e : row[i])
        if (e.first == j) { e.second += rx.k; merged = true; }
      if (!merged) row[i].push_back({j, rx.k});
      own[i].push_back({rx.k, r, j});
    }
  }
  out->groups.clear();
  out->consts.clear();
  for (uint32_t i = 0; i <= S; ++i) {
    if (own[i].empty()) continue;
    RbPdmGroup g;
    g.species = i;
    g.const_first = (uint32_t)out->consts.size();
    out->consts.push_back(c[i]);
    for (const auto& e : row[i]) {
      g.partners.push_back(e.first);
      out->consts.push_back(e.second);
    }
    g.own = own[i];
    out->groups.push_back(g);
  }
  if (out->consts.size() > RB_MAX_K)
    return no("at most " + std::to_string(RB_MAX_K) + " derived constants (owner species + distinct reactant pairs)");
  const unsigned ng = (unsigned)out->groups.size();
  out->group_size = 1;
  unsigned max_ck = RB_PDM_MAX_CHECKPOINTS;
  if (const char* env = std::getenv("REBOP_B200_PDM_CK")) max_ck = (unsigned)std::max(1, std::atoi(env));  // development knob
  while ((ng + out->group_size - 1) / out->group_size > max_ck) ++out->group_size;
  out->n_checkpoints = ng ? (ng + out->group_size - 1) / out->group_size : 1;

  // device image
  const unsigned nb = out->n_checkpoints;
  uint32_t header[4] = {nb, 0, 0, R};
  size_t off = 2;  // header: 16 bytes
  header[1] = (uint32_t)off; off += (nb + 1 + 1) / 2;
  off += off & 1;  // 16-byte alignment of the entries
  header[2] = (uint32_t)off; off += 2 * (size_t)R;
  out->image.assign(off + 2, 0);
  uint32_t* bp = reinterpret_cast<uint32_t*>(out->image.data() + header[1]);
  uint64_t* ew = out->image.data() + header[2];
  // Inside a block the walk defines the sub-intervals itself (the unrolled pass only fixes the block sums), so the
  // reactions of a block may come in any order: the ones likely to carry most of the block's propensity first, so
  // that most lanes leave the walk after its first step.  The weight is a static guess -- rate constant, times a
  // typical count for the partner of a second-order reaction -- and only affects speed.
  struct Entry { double k; uint32_t i, j, reaction, kind; double weight; };
  uint32_t ne = 0;
  for (unsigned b0 = 0; b0 < ng; b0 += out->group_size) {
    bp[b0 / out->group_size] = ne;
    std::vector<Entry> block;
    for (unsigned g = b0; g < ng && g < b0 + out->group_size; ++g) {
      const RbPdmGroup& grp = out->groups[g];
      for (const RbPdmGroup::Own& e : grp.own) {
        const uint32_t kind = e.partner == RB_PDM_NONE ? 0u : e.partner == grp.species ? 2u : 1u;
        block.push_back({e.k, grp.species, kind == 1u ? e.partner : grp.species, e.reaction, kind, e.k * (kind ? 100.0 : 1.0)});
      }
    }
    std::stable_sort(block.begin(), block.end(), [](const Entry& a, const Entry& b) { return a.weight > b.weight; });
    for (const Entry& e : block) {
      std::memcpy(ew + 2 * ne, &e.k, 8);
      ew[2 * ne + 1] = (uint64_t)(e.i | (e.j << 16)) | ((uint64_t)(e.reaction | (e.kind << 30)) << 32);
      ++ne;
    }
  }
  for (unsigned b2 = ng ? (ng - 1) / out->group_size + 1 : 0; b2 <= nb; ++b2) bp[b2] = ne;
  std::memcpy(out->image.data(), header, sizeof header);
  return REBOP_OK;
}
