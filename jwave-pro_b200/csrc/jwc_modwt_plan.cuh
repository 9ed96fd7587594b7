// jwc_modwt_plan.cuh -- how a J-level MODWT is cut into fused passes.
//
// A pass handles levels j0+1 .. j0+k of every signal on tiles that stay in shared memory.  Level j0+jj convolves with
// stride 2^(j0+jj-1); seen on the 2^j0 phase subsequences n = i*2^j0 + ph of V_{j0} this is a stride-2^(jj-1)
// convolution in the decimated index i, so a tile is  P phases x (T2 + halo) decimated samples  with halo
// (L-1)(2^k - 1) decimated samples -- independent of j0.  That is what keeps the halo (and the redundant work on it)
// small for long filters / deep levels (Daubechies20, J = 8: 9945 samples in one pass, 585 in each of two passes).
// Pass boundaries cost one extra write + read of V (16 B/sample); the planner trades that against halo overhead with
// a two-roof (HBM, fp64) time model.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#ifndef JWC_HD
#ifdef __CUDACC__
#define JWC_HD __host__ __device__
#else
#define JWC_HD
#endif
#endif

namespace jwc {

#ifndef JWC_MODWT_R
#define JWC_MODWT_R 7
#endif
constexpr int kModwtR = JWC_MODWT_R;  // outputs per work item; odd => conflict-free LDS.64/STS.64 for every stride (see kernel)

enum { MODE_BULK = 0, MODE_VEC2 = 1, MODE_SCALAR = 2 };

struct ModwtPass {
  int j0 = 0, k = 0, logP = 0, T2 = 0, Hp = 0, mode = MODE_SCALAR, vcap = 0, threads = 256;
  size_t smem = 0;
};

struct ModwtPlan {
  std::vector<ModwtPass> passes;
  int generic_from = 0;  // levels > generic_from (1-based: generic_from+1 .. J) run on the per-level generic kernels; 0 = none... see below
  bool all_fused = true;
};

struct ModwtPlanInput {
  int64_t n;
  int J, L;
  bool aligned16;      // every base pointer 16-byte aligned
  int smem_budget;     // bytes per CTA the plan may use
  int tile_override, group_override, threads_override;
  int logp_override = 0, tile_deep_override = 0;   // phase-split passes (j0 > 0) only
  int plan_override = 0;   // > 0: pass depths as decimal digits, first pass first (422 = 4 + 2 + 2 levels); experiments
  bool inverse;        // inverse needs (V ping-pong + W double buffer), forward (V ping-pong + W staging)
};

// blockIdx / tiles without a division: the quotient of n < 2^31 by d is (mulhi(m, n) + n) >> l with l = ceil(log2 d),
// m = floor(2^32 (2^l - d) / d) + 1 (Granlund-Montgomery; the sum cannot overflow for n < 2^31).  The run-time
// division it replaces is ~22 dependent instructions (I2F, MUFU.RCP, F2I, fix-ups) at the very top of every CTA, in
// front of its first tile request (tests/test_plan_bounds.py::test_fastdiv checks it on the CPU).
struct FastDiv {
  unsigned m;
  int l;
};
inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f{1u, 0};
  while (((uint64_t)1 << f.l) < d) f.l++;
  f.m = (unsigned)(((((uint64_t)1 << f.l) - d) << 32) / d + 1);
  return f;
}
JWC_HD inline unsigned fastdiv(unsigned n, const FastDiv& f) {
#ifdef __CUDA_ARCH__
  return (__umulhi(f.m, n) + n) >> f.l;
#else
  return ((unsigned)(((uint64_t)f.m * n) >> 32) + n) >> f.l;
#endif
}

inline int64_t modwt_halo(int L, int k) { return (int64_t)(L - 1) * (((int64_t)1 << k) - 1); }

// The walk t -> t + 2^j0 (mod n) on a circular signal splits it into G = gcd(2^j0, n) interleaved cycles of n / G
// positions each (G = 2^j0 when 2^j0 divides n: the plain phase subsequences).  Along a cycle, every level of a pass
// that starts at depth j0 is a stride-2^(jj-1) filter in the cycle index, so the tile scheme works for ANY n; only the
// address of cycle index i changes from i 2^j0 to (i 2^j0) mod n.
inline int64_t modwt_cycles(int64_t n, int j0) {
  const int64_t S0 = (int64_t)1 << j0, low = n & (-n);
  return std::min(S0, low);
}

// shared-memory doubles for one CTA of a pass
inline int64_t modwt_smem_doubles(bool inverse, int P, int T2, int Hp, int k) {
  const int64_t pad = (int64_t)kModwtR * P * ((int64_t)1 << (k - 1));
  const int64_t vcap = (int64_t)P * (T2 + Hp) + pad;
  if (inverse) return 4 * (vcap + (vcap & 1));                       // V0 V1 W0 W1, all tile+halo sized
  if (P >= 4) return 2 * (vcap + (vcap & 1));                        // V0 V1; W goes straight from registers to HBM
  return 2 * (vcap + (vcap & 1)) + 2 * (int64_t)P * T2;              // V0 V1 + two W staging tiles
}

// work items of level jj (1-based inside the pass) for a full tile, see the kernels' item loops
inline int64_t modwt_items(bool inverse, int L, int k, int jj, int logP, int T2) {
  const int64_t P = (int64_t)1 << logP, s = P << (jj - 1);
  const int64_t halo = inverse ? (int64_t)(L - 1) * (((int64_t)1 << (jj - 1)) - 1)
                               : (int64_t)(L - 1) * (((int64_t)1 << k) - ((int64_t)1 << jj));
  const int64_t len = P * (T2 + halo);
  const int64_t rows = (len + s - 1) / s;
  return ((rows + kModwtR - 1) / kModwtR) * s;
}

// fraction of thread-rounds doing useful items when `threads` threads share the items of every level
inline double modwt_lane_efficiency(bool inverse, int L, int k, int logP, int T2, int threads) {
  double useful = 0, issued = 0;
  for (int jj = 1; jj <= k; jj++) {
    const int64_t it = modwt_items(inverse, L, k, jj, logP, T2);
    useful += (double)it;
    issued += (double)(((it + threads - 1) / threads) * threads);
  }
  return issued > 0 ? useful / issued : 1.0;
}

inline bool modwt_make_pass_p(const ModwtPlanInput& in, int j0, int k, int logP, ModwtPass* out, double* est_time) {
  const int64_t G = modwt_cycles(in.n, j0);
  const int P = 1 << logP;
  if (P > G) return false;                 // the phases of a CTA are consecutive cycles
  // odd n: one cycle, every row is a lone 8-byte gather / scatter (a quarter of each 32-byte sector used).  Worth it for
  // the fp64-bound long filters (Daubechies20 J = 8 on 99 999 samples: 48 ms fused against 129 ms on the per-level
  // kernels), not for the HBM-bound short ones (Daubechies4 J = 10: 43 ms against 41 ms)
  if (j0 > 0 && G == 1 && in.L <= 10) return false;
  const int64_t Nd = in.n / G;             // cycle length (= n >> j0 when 2^j0 divides n)
  int64_t H = modwt_halo(in.L, k);
  int mode;
  if (j0 == 0) mode = (in.aligned16 && (in.n % 2) == 0 && H + 1 <= in.n) ? MODE_BULK : MODE_SCALAR;
  else mode = (in.aligned16 && logP >= 1) ? MODE_VEC2 : MODE_SCALAR;
  int64_t Hp = H + ((mode == MODE_BULK) ? (H & 1) : 0);
  const int64_t budget = in.smem_budget / 8 - 8 - 128;   // mbarriers + shared-memory tap copy
  // largest even T2 that fits
  int64_t lo = 2, hi = std::max<int64_t>(2, Nd + (Nd & 1)), best = 0;
  const int tile_ov = (j0 > 0 && in.tile_deep_override > 0) ? in.tile_deep_override : in.tile_override;
  if (tile_ov > 0) hi = std::min<int64_t>(hi, tile_ov);
  if (modwt_smem_doubles(in.inverse, P, (int)hi, (int)Hp, k) <= budget) best = hi;
  else {
    while (lo <= hi) {
      int64_t mid = ((lo + hi) / 2) & ~(int64_t)1;
      if (mid < 2) mid = 2;
      if (modwt_smem_doubles(in.inverse, P, (int)mid, (int)Hp, k) <= budget) { best = mid; lo = mid + 2; }
      else hi = mid - 2;
    }
    if (best >= 128) best &= ~(int64_t)63;   // keep tiles 512-byte multiples
  }
  if (best < 2) return false;
  if (best < Nd && best < H && tile_ov <= 0) return false;  // halo would dominate the tile
  // Candidates: the largest tile that fits and smaller ones in steps of 64 (down to 5/8 of it), each with 128 or 256
  // threads.  Whole multiples of 128 threads only: a warp's scheduler is fixed by (warp index % 4), so 160- or
  // 192-thread CTAs put twice the work on one or two of the four schedulers of the SM (the fp64 pipe then idles at
  // 62-75 % of its rate however many CTAs are resident).  The tile length decides how well the items of every level
  // fill whole rounds of threads; the model below weighs that against the halo overhead of a shorter tile.
  double best_est = 1e300;
  bool found = false;
  const int64_t tmax = best;
  // (forward only.  The inverse keeps the largest tile and picks among 128..256 threads in steps of 32 by lane
  // efficiency, as in round 1: measured on Daubechies20 J = 8, the search below cost it 10 %.)
  // ... and only for the fp64-bound long filters: the HBM-bound short ones want the largest tile (Daubechies4 on
  // 100 000 samples: 2.42 ms with the largest tile and 256 threads, 2.62 ms with the searched 1984 x 128).
  const bool legacy = in.inverse || in.L <= 10;
  const int64_t tmin = (legacy || tile_ov > 0 || tmax >= Nd || tmax < 512)
                           ? tmax : std::max<int64_t>(H, (tmax * 5 / 8) & ~(int64_t)63);
  int inv_thr = 256;
  if (legacy) {
    double e0 = 0;
    const int Tf = (int)std::min<int64_t>(tmax, Nd);
    for (int cand = 256; cand >= 128; cand -= 32) {
      const double e = modwt_lane_efficiency(in.inverse, in.L, k, logP, Tf, cand);
      if (e > e0 + 0.02) { e0 = e; inv_thr = cand; }
    }
    // fp64-bound long filters: whole multiples of 128 threads only (see above) -- the 160-thread CTAs this search liked
    // for the phase-split inverse passes of Daubechies20 put two of their five warps on one scheduler.  With 128 the
    // model then prefers levels 5-7 + level 8 over 5-6 + 7-8 behind the first pass: 34.0 ms against 34.7 (and 36.4
    // with the 160-thread plan) on one box.
    if (in.inverse && in.L > 10) inv_thr = 128;

  }
  for (int64_t T2 = tmax; T2 >= tmin && T2 >= 2; T2 -= 64) {
    const int Tfull = (int)std::min<int64_t>(T2, Nd);
    for (int cand = 128; cand <= 256; cand += 128) {   // 128 first: smaller CTAs = more of them resident; 256 must win clearly
      if ((in.threads_override > 0 || legacy) && cand != 128) continue;
      const int thr = in.threads_override > 0 ? in.threads_override : (legacy ? inv_thr : cand);
      const double eff = modwt_lane_efficiency(in.inverse, in.L, k, logP, Tfull, thr);
      // time model per input sample (units: ps): memory and fp64 overlap only partly inside a CTA, short passes worst
      const double T = (double)Tfull;
      double avg_hrem = 0;
      for (int jj = 1; jj <= k; jj++)
        avg_hrem += in.inverse ? (double)(in.L - 1) * (double)(((int64_t)1 << (jj - 1)) - 1)
                               : (double)(in.L - 1) * (double)(((int64_t)1 << k) - ((int64_t)1 << jj));
      avg_hrem /= k;
      const double sector = (j0 == 0) ? 1.0 : (P >= 4 ? 1.0 : (P == 2 ? 1.6 : 2.5));   // strided 16 B / 8 B rows waste sectors
      const double bytes = sector * (8.0 * (k + 2) + 8.0 * (double)Hp / T * (in.inverse ? (k + 1) * 0.6 : 1.0));
      // the last tile of a signal is shorter: its CTA costs a full tile's latency for a fraction of the work
      const double tiles = std::ceil((double)Nd / T);
      const double edge = (tiles * T) / (double)Nd;
      const double flops = 4.0 * in.L * k * (1.0 + avg_hrem / T) / std::max(eff, 0.3);
      const double tm = bytes / 5.5, tc = flops / 32.0;
      double est = (std::max(tm, tc) + 0.35 * std::min(tm, tc)) * (0.5 + 0.5 * edge) + 2.0;
      // a CTA costs about 4 ns of GPU time whatever it does (launch, barriers, mbarrier round trips): measured on
      // windows of 512 samples, where a phase-split pass with 32-sample CTAs ran at 127 ps/sample
      const double cta_samples = (double)P * T;
      if (cta_samples < 1024.0) est += 4000.0 / cta_samples;
      if (cand == 256) est *= 1.04;
      if (!legacy) {   // too few resident warps cannot cover the load / barrier phases of the other CTAs of the SM
        const double smem_bytes = (double)modwt_smem_doubles(in.inverse, P, (int)T2, (int)Hp, k) * 8 + 2048;
        const int ctas = std::max(1, std::min(8, (int)(232448.0 / smem_bytes)));
        const int warps = ctas * thr / 32;
        if (warps < 12) est *= 1.0 + 0.05 * (12 - warps);
      }
      if (est < best_est - 1e-9) {
        best_est = est;
        found = true;
        out->T2 = (int)T2;
        out->threads = thr;
      }
    }
  }
  if (!found) return false;
  best = out->T2;
  out->j0 = j0; out->k = k; out->logP = logP; out->Hp = (int)Hp; out->mode = mode;
  const int64_t pad = (int64_t)kModwtR * P * ((int64_t)1 << (k - 1));
  int64_t vcap = (int64_t)P * (best + Hp) + pad;
  vcap += vcap & 1;
  out->vcap = (int)vcap;
  out->smem = (size_t)modwt_smem_doubles(in.inverse, P, (int)best, (int)Hp, k) * 8 + 1024 + 64;
  *est_time = best_est;
  return true;
}

inline bool modwt_make_pass(const ModwtPlanInput& in, int j0, int k, ModwtPass* out, double* est_time) {
  if (j0 == 0) return modwt_make_pass_p(in, j0, k, 0, out, est_time);
  bool ok = false;
  double bt = 1e300;
  for (int logP = std::min(j0, 2); logP >= 1; --logP) {   // 4 phases (32-byte rows) or 2 (16-byte rows)
    if (in.logp_override > 0 && logP != std::min(j0, in.logp_override)) continue;
    if (((int64_t)1 << logP) > modwt_cycles(in.n, j0)) continue;
    ModwtPass p;
    double t;
    if (modwt_make_pass_p(in, j0, k, logP, &p, &t) && t < bt) { bt = t; *out = p; ok = true; }
  }
  if (!ok) {   // last resort: single phase, scalar gathers
    ModwtPass p;
    double t;
    if (modwt_make_pass_p(in, j0, k, 0, &p, &t)) { bt = t; *out = p; ok = true; }
  }
  *est_time = bt;
  return ok;
}

// dynamic programme over pass boundaries; levels that cannot be fused (no tile fits the budget) fall to the generic
// per-level kernels (24 B/sample/level, no fusion).
inline ModwtPlan modwt_plan(const ModwtPlanInput& in) {
  const int J = in.J;
  std::vector<double> best(J + 1, 1e300);
  std::vector<int> choice(J + 1, 0);   // k > 0: fused pass of k levels; -1: everything from here generic
  std::vector<ModwtPass> pass_at(J + 1);
  best[J] = 0;
  for (int j0 = J - 1; j0 >= 0; j0--) {
    // per-level generic kernels: 24 B/sample/level of traffic at about half the copy rate, and their uncoalesced,
    // unfused inner loop reaches roughly a quarter of the fp64 peak for long filters
    const double gen = (2.0 * 24.0 / 5.5 + 4.0 * in.L / 32.0 * 4.0) * (J - j0);
    best[j0] = gen;
    choice[j0] = -1;
    const int kmax = in.group_override > 0 ? std::min(in.group_override, J - j0) : J - j0;
    for (int k = 1; k <= kmax; k++) {
      ModwtPass p;
      double t;
      if (!modwt_make_pass(in, j0, k, &p, &t)) continue;
      if (t + best[j0 + k] < best[j0]) {
        best[j0] = t + best[j0 + k];
        choice[j0] = k;
        pass_at[j0] = p;
      }
    }
  }
  if (in.plan_override > 0) {   // follow the digits as far as they go and as far as passes can be made
    int digits[12], nd = 0;
    for (int v = in.plan_override; v > 0 && nd < 12; v /= 10) digits[nd++] = v % 10;
    ModwtPlan forced;
    int j = 0;
    for (int d = nd - 1; d >= 0 && j < J; d--) {
      const int k = std::min(digits[d], J - j);
      ModwtPass p;
      double t;
      if (k < 1 || !modwt_make_pass(in, j, k, &p, &t)) break;
      forced.passes.push_back(p);
      j += k;
    }
    if (j == J) {
      forced.generic_from = J;
      forced.all_fused = true;
      return forced;
    }
  }
  ModwtPlan plan;
  int j0 = 0;
  while (j0 < J && choice[j0] > 0) {
    plan.passes.push_back(pass_at[j0]);
    j0 += choice[j0];
  }
  plan.generic_from = j0;          // levels j0+1 .. J (if any) are generic
  plan.all_fused = (j0 == J);
  return plan;
}

}  // namespace jwc
