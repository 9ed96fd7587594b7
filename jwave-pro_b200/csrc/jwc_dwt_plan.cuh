// jwc_dwt_plan.cuh -- how a decimated transform (FWT pyramid / WPT full tree) is cut into fused passes.
//
// A pass takes every node of length h at tree depth l0 and carries it k levels down inside shared memory:
// tile = T consecutive samples of one node + the halo the k decimating steps need:
//   forward (analysis):  right halo (L-2)(2^k - 1) input samples
//   inverse (synthesis): left halo of HL_jj <= L-2 coefficients in every depth-jj child array (it does not grow)
// FWT: only the low-pass chain continues, each D_{l0+jj} tile is shipped as soon as it exists.
// WPT: all 2^jj children stay in shared memory, the 2^k leaves are shipped at the end.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace jwc {

#ifndef JWC_HD
#ifdef __CUDACC__
#define JWC_HD __host__ __device__
#else
#define JWC_HD
#endif
#endif

// Rows per work item: outputs (forward) / output pairs (inverse) one thread computes with a sliding register window.
// ODD counts only: the lanes of a warp then hit distinct banks (LDS.128 / STS.128 with a lane stride of R x 16 bytes).
// kDwtR is the largest count and the padding of every node array in shared memory (an item at the end of a node reads
// up to 2(R-1) doubles past it).  Filters longer than 10 taps use at most 5 rows (registers).
constexpr int kDwtR = 7;
JWC_HD inline constexpr int dwt_rmax(int L) { return L > 10 ? 5 : kDwtR; }

// Rows per item of one level: as many as the build offers, unless that would leave more than 3/4 of the threads
// without an item (deep, small levels) -- then 3, then 1.  (Round 2 measured the alternative "fewest rounds of items x
// (R + overhead)", which picks 3 rows earlier: FWT Daubechies8 8 % slower, the fixed cost per item is what matters.)
JWC_HD inline int dwt_pick_r(int L, int len, int parents, int nt) {
  const int RM = dwt_rmax(L);
  if (4 * parents * ((len + RM - 1) / RM) >= nt) return RM;
  if (4 * parents * ((len + 2) / 3) >= nt) return 3;
  return 1;
}

enum { DWT_BULK = 0, DWT_SCALAR = 2 };

struct DwtPass {
  int l0 = 0, k = 0, T = 0, cap = 0, threads = 256, mode = DWT_SCALAR;
  size_t smem = 0;
};

struct DwtPlanInput {
  int64_t n;
  int levels, L;
  bool tree, inverse, aligned16;
  int smem_budget, tile_override, group_override, threads_override;
  int k0_override = 0;   // > 0: depth of the first pass (l0 = 0) is forced (experiments)
  int fixed_override = 0;   // > 0: per-pass fixed cost of the pyramid model in 0.01 ps per sample (experiments)
};

// ---- forward geometry -------------------------------------------------------------------------------------------------
inline int64_t dwt_fwd_halo(int L, int k, int jj) { return (int64_t)(L - 2) * (((int64_t)1 << (k - jj)) - 1); }
inline int64_t dwt_fwd_len(int L, int k, int jj, int64_t tlen) { return (tlen >> jj) + dwt_fwd_halo(L, k, jj); }
inline int64_t dwt_node_stride(int64_t len) { return len + (len & 1) + 2 * kDwtR; }
inline int64_t dwt_fwd_cap(bool tree, int L, int k, int64_t tlen) {
  int64_t cap = 0;
  for (int jj = 0; jj <= k; jj++) {
    const int64_t nodes = (jj == 0) ? 1 : (tree ? ((int64_t)1 << jj) : 2);
    cap = std::max(cap, nodes * dwt_node_stride(dwt_fwd_len(L, k, jj, tlen)));
  }
  return cap + (cap & 1);
}

// ---- inverse geometry: left halo of the depth-jj arrays (rounded up to even so bulk copies stay 16-byte aligned)
inline int64_t dwt_inv_halo(int L, int jj) {
  int64_t h = 0;
  for (int q = 1; q <= jj; q++) {
    h = (h + 1) / 2 + (L / 2 - 1);
    h += h & 1;
  }
  return h;
}
inline int64_t dwt_inv_len(int L, int jj, int64_t tlen) { return (tlen >> jj) + dwt_inv_halo(L, jj); }
inline int64_t dwt_inv_cap(bool tree, int L, int k, int64_t tlen) {
  int64_t cap = 0;
  for (int jj = 0; jj <= k; jj++) {
    const int64_t nodes = (jj == 0) ? 1 : (tree ? ((int64_t)1 << jj) : 2);
    cap = std::max(cap, nodes * dwt_node_stride(dwt_inv_len(L, jj, tlen)));
  }
  return cap + (cap & 1);
}

// outputs (forward) / output pairs (inverse) of level jj per parent node, as the kernels count them
inline int64_t dwt_level_len(const DwtPlanInput& in, int k, int jj, int64_t tlen) {
  if (!in.inverse) return dwt_fwd_len(in.L, k, jj, tlen);
  return (dwt_inv_halo(in.L, jj - 1) >> 1) + (tlen >> jj);
}
// item-rounds x rows per item of level jj with `thr` threads: what the level costs in units of one output row
inline double dwt_level_cost(const DwtPlanInput& in, int k, int jj, int64_t tlen, int thr) {
  const int64_t parents = in.tree ? ((int64_t)1 << (jj - 1)) : 1;
  const int64_t len = dwt_level_len(in, k, jj, tlen);
  const int R = dwt_pick_r(in.L, (int)len, (int)parents, thr);
  const int64_t items = parents * ((len + R - 1) / R);
  return (double)((items + thr - 1) / thr) * (R + 2);
}

inline bool dwt_make_pass(const DwtPlanInput& in, int l0, int k, DwtPass* out, double* est) {
  const int64_t h = in.n >> l0;                 // node length at the top of the pass
  if (k < 1 || ((int64_t)1 << k) > h) return false;
  const int64_t budget = in.smem_budget / 8 - 16 - 2 * 64;   // mbarriers + shared-memory tap copy
  auto cap_of = [&](int64_t t) { return in.inverse ? dwt_inv_cap(in.tree, in.L, k, t) : dwt_fwd_cap(in.tree, in.L, k, t); };
  auto smem_doubles = [&](int64_t t) { return 2 * cap_of(t); };
  // T: a power of two, 2^k <= T <= h, as large as the budget allows
  int64_t T = h;
  if (in.tile_override > 0)
    while (T > in.tile_override && T > ((int64_t)1 << k)) T >>= 1;
  while (T > ((int64_t)1 << k) && smem_doubles(T) > budget) T >>= 1;
  if (smem_doubles(T) > budget) return false;
  const int64_t H = in.inverse ? dwt_inv_halo(in.L, k) : dwt_fwd_halo(in.L, k, 0);
  if (T < h && !in.inverse && 2 * H > T && in.tile_override <= 0) return false;   // halo would dominate the tile
  out->l0 = l0; out->k = k; out->T = (int)T; out->cap = (int)cap_of(T);
  out->smem = (size_t)smem_doubles(T) * 8 + 2 * 64 * 8 + 128;
  // bulk (TMA) copies need even piece lengths: forward T >= 2, inverse (T >> k) even
  const bool even_ok = in.inverse ? ((T >> k) % 2 == 0) : (T >= 2);
  out->mode = (in.aligned16 && even_ok) ? DWT_BULK : DWT_SCALAR;
  // threads: 128 (measured best on B200 for FWT db8 and WPT sym8: more resident CTAs hide the per-level barriers)
  const int thr = in.threads_override > 0 ? in.threads_override : 128;
  double eff;
  {
    // useful output rows per thread vs. what the rounds of items actually cost (see dwt_pick_r)
    double useful = 0, issued = 0;
    for (int jj = 1; jj <= k; jj++) {
      const double parents = in.tree ? (double)((int64_t)1 << (jj - 1)) : 1.0;
      useful += parents * (double)dwt_level_len(in, k, jj, T) / thr;
      issued += dwt_level_cost(in, k, jj, T, thr);
    }
    eff = issued > 0 ? useful / issued : 1.0;
  }
  out->threads = thr;
  // time model per sample of the pass input (ps)
  double work = 0;   // output pairs computed per input sample, summed over the levels
  for (int jj = 1; jj <= k; jj++) {
    const double parents = in.tree ? (double)((int64_t)1 << (jj - 1)) : 1.0;
    const double len = in.inverse ? (double)dwt_inv_len(in.L, jj - 1, T) / 2.0 : (double)dwt_fwd_len(in.L, k, jj, T);
    work += parents * len / (double)T;
  }
  const double flops = 4.0 * in.L * work / std::max(eff, 0.3);
  double halo_rd;
  if (!in.inverse) halo_rd = (double)H / (double)T;
  else halo_rd = (in.tree ? (double)((int64_t)1 << k) * (double)H : (double)(k + 1) * (double)(in.L - 2)) / (double)T;
  const double bytes = 8.0 * (1.0 + halo_rd) + 8.0;
  const double tm = bytes / 5.5, tc = flops / 32.0;
  const double frac = in.tree ? 1.0 : 1.0 / (double)((int64_t)1 << l0);   // FWT passes shrink geometrically
  // every level is a block-wide barrier whose latency is only partly hidden by the other resident CTAs
  *est = frac * (std::max(tm, tc) + 0.35 * std::min(tm, tc) + 0.25 * k) + (in.tree ? 1.0 : (in.fixed_override > 0 ? 0.01 * in.fixed_override : 0.05));
  return true;
}

struct DwtPlan {
  std::vector<DwtPass> passes;   // in forward order (l0 ascending); the inverse executes them back to front
  bool ok = false;
};

// steps = number of levels actually performed (reference loops stop at h < 2)
inline DwtPlan dwt_plan(const DwtPlanInput& in, int steps) {
  DwtPlan plan;
  if (steps <= 0) return plan;
  std::vector<double> best(steps + 1, 1e300);
  std::vector<int> choice(steps + 1, 0);
  std::vector<DwtPass> pass_at(steps + 1);
  best[steps] = 0;
  for (int l0 = steps - 1; l0 >= 0; l0--) {
    const int kmax = in.group_override > 0 ? std::min(in.group_override, steps - l0) : steps - l0;
    for (int k = 1; k <= kmax; k++) {
      DwtPass p;
      double t;
      if (l0 == 0 && in.k0_override > 0 && k != std::min(in.k0_override, steps)) continue;
      if (!dwt_make_pass(in, l0, k, &p, &t)) continue;
      if (best[l0 + k] < 1e299 && t + best[l0 + k] < best[l0]) {
        best[l0] = t + best[l0 + k];
        choice[l0] = k;
        pass_at[l0] = p;
      }
    }
  }
  if (best[0] >= 1e299) return plan;
  for (int l0 = 0; l0 < steps; l0 += choice[l0]) plan.passes.push_back(pass_at[l0]);
  plan.ok = true;
  return plan;
}

}  // namespace jwc
