// jwc_modwt_fast.cu -- fused multi-level MODWT forward on shared-memory tiles (sm_100a).
//
// Computes, for levels j0+1 .. j0+k of every signal (reference: transforms/MODWTTransform.java:290-304 with the direct
// convolution :677-690, paths relative to /root/reference/src/main/java/jwave/):
//     W_j[t] = sum_m h~[m] V_{j-1}[(t - m 2^(j-1)) mod N],   V_j[t] = sum_m g~[m] V_{j-1}[(t - m 2^(j-1)) mod N]
//
// Data layout / traffic: the input tile (T2 outputs + left halo (L-1)(2^k-1)) is read from HBM ONCE (1-D bulk TMA
// copy, the circular wrap is done by splitting the copy at the signal boundary), all k levels are computed in shared
// memory (V ping-pongs between two buffers), each W_j tile is staged in shared memory and leaves through an
// asynchronous bulk store, so HBM sees one read and k+1 writes per sample: 8(k+2) bytes.
//
// Inner loop ("column scheme"): at stride s the tile is a matrix M[row][col], position = row*s + col, and the
// a-trous filter is a vertical FIR down each column.  One work item = R consecutive rows of one column: it slides a
// window down the column, R + L - 1 LDS.64 for 2*R*L DFMA (taps are immediate constant-bank operands), so shared
// memory bandwidth (128 B/clk/SM) stays below the fp64 pipe.  Lanes map to consecutive columns (s >= 16) or to
// (column, row-block) pairs whose addresses are distinct mod 16 doubles because R is odd -> conflict-free for all s.
#include <cstdlib>

#include "jwc_internal.cuh"
#include "jwc_modwt_plan.cuh"
#include "jwc_tma.cuh"

namespace jwc {

namespace {

// JWC_DEBUG=1 prints every plan once per distinct shape to stderr
void debug_plan(const char* what, const ModwtPlan& plan, int64_t n, int levels, int L) {
  static const bool on = getenv("JWC_DEBUG") != nullptr;
  if (!on) return;
  fprintf(stderr, "[jwc] %s n=%lld J=%d L=%d:", what, (long long)n, levels, L);
  for (const ModwtPass& p : plan.passes)
    fprintf(stderr, " [j0=%d k=%d P=%d T2=%d Hp=%d mode=%d smem=%zu thr=%d]", p.j0, p.k, 1 << p.logP, p.T2, p.Hp, p.mode,
            p.smem, p.threads);
  if (!plan.all_fused) fprintf(stderr, " generic from level %d", plan.generic_from + 1);
  fprintf(stderr, "\n");
}

// CTAs resident on the whole device for this launch shape ~ one wave; that is how far ahead tiles are pulled into L2
int prefetch_distance(jwc_ctx* ctx, const DeviceSlot& dev, size_t smem, int threads) {
  (void)smem; (void)threads;
  if (ctx->tune.l2_prefetch < 0) return 0;
  if (ctx->tune.l2_prefetch > 0) return ctx->tune.l2_prefetch;
  return dev.sm_count;   // measured on B200 (C2 forward, 2 CTAs/SM): half a wave ahead 3.15 ms, a wave 3.19, two waves 3.32
}

// Element offset of row i (0 <= i < Nd) of a phase-split pass.  When 2^j0 divides N the rows of one phase are
// i 2^j0 apart; otherwise the walk t -> t + 2^j0 (mod N) closes after N / gcd(2^j0, N) positions and row i sits at
// (i (2^j0 mod N)) mod N (jwc_modwt_plan.cuh, modwt_cycles).  The quotient comes from one fp64 multiply (off by at
// most one, corrected) -- a 64-bit integer modulo is a ~100-instruction subroutine.
struct RowMap {
  int64_t S0, Sm, N;   // 2^j0, 2^j0 mod N, signal length
  double invN;
  int wrap;            // 0: offset = i * S0 (kernels instantiated with WRAP = false: the code of the round-1 paths)
  // RT: the WRAP instantiation serves every shape and tests `wrap` at run time (long-filter forward kernel, see there)
  template <bool WRAP, bool RT = false>
  __device__ __forceinline__ int64_t at(int64_t i, int64_t s0) const {   // s0 = 2^j0 as the caller already has it
    if (!WRAP || (RT && !wrap)) return i * s0;
    const uint64_t prod = (uint64_t)i * (uint64_t)Sm;   // < 2^62
    const uint64_t q = (uint64_t)((double)prod * invN);
    int64_t r = (int64_t)(prod - q * (uint64_t)N);
    while (r < 0) r += N;
    while (r >= N) r -= N;
    return r;
  }
};

struct FwdPassArgs {
  const double* in;   // V_{j0}
  double* coeffs;     // coefficient block base
  double* vout;       // V_{j0+k}
  int64_t in_sig, coeff_sig, vout_sig;
  int64_t N, Nd;
  int j0, k, logP, T2, Hp, tiles_i, groups, vcap, mode;
  FastDiv tiles_div;   // division by tiles_i
  int log_groups;      // groups is a power of two (cycles and phases per CTA are)
  int pf_dist;   // L2 prefetch distance in CTAs (0 = off): the tile of CTA blockIdx + pf_dist is pulled into L2
  unsigned nblocks;
  RowMap rm;     // address of a decimated row (cycle index) of a phase-split pass
};

// Long filters: 2L taps do not fit the uniform-register file (63 x 32 bit), and ptxas then feeds every DFMA through
// LDC + R2UR moves.  For L > kUniformTapsMax the taps are read from a shared-memory copy (broadcast LDS.64) into a
// rolling window of R live taps per filter, loaded exactly when the sliding window first needs them.
__device__ __forceinline__ const double* const_taps(const FilterPair& f) {
  // index the compiler cannot prove uniform => every tap is one LDC.64 into an ordinary register (see jwc_dwt_fast.cu)
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return reinterpret_cast<const double*>(&f) + z;   // f0 at [0, 64), f1 at [64, 128)
}
constexpr int kUniformTapsMax = 10;
constexpr int kTapDoubles = 2 * JWC_MAX_TAPS;   // shared-memory copy of FilterPair: f0 at [0..64), f1 at [64..128)

template <int L, int R, bool WITH_W>
__device__ __forceinline__ void fwd_item(const double* __restrict__ top, int s, const FilterPair& f,
                                         const double* __restrict__ taps, double (&av)[R], double (&aw)[R]) {
  constexpr bool ST = (L > kUniformTapsMax);
#pragma unroll
  for (int q = 0; q < R; q++) { av[q] = 0.0; aw[q] = 0.0; }
  double tg[L], th[L];
  // rows R-1 down to -(L-1): every accumulator meets its taps in ascending m, the reference's order
  const double* p = top;
#pragma unroll
  for (int i = R - 1; i >= -(L - 1); --i) {
    if (ST) {
      const int mn = R - 1 - i;   // the one tap index this row uses for the first time
      if (mn < L) {
        tg[mn] = taps[mn];
        if (WITH_W) th[mn] = taps[JWC_MAX_TAPS + mn];
      }
    }
    const double x = *p;
    p -= s;
#pragma unroll
    for (int q = 0; q < R; q++) {
      const int m = q - i;
      if (m >= 0 && m < L) {
        av[q] = fma(x, ST ? tg[m] : f.f0[m], av[q]);
        if (WITH_W) aw[q] = fma(x, ST ? th[m] : f.f1[m], aw[q]);
      }
    }
  }
}

// row index modulo the decimated signal length: one conditional add/subtract covers every tile whose halo is shorter
// than the signal; only the pathological multi-wrap shapes pay for the 64-bit modulo
__device__ __noinline__ int64_t wrap_row_slow(int64_t i, int64_t nd) {
  i %= nd;
  return i < 0 ? i + nd : i;
}
__device__ __forceinline__ int64_t wrap_row(int64_t i, int64_t nd) {
  if (i < 0) {
    i += nd;
    if (i < 0) i = wrap_row_slow(i, nd);   // pathological multi-wrap shapes only (halo longer than the signal)
  } else if (i >= nd) {
    i -= nd;
    if (i >= nd) i = wrap_row_slow(i, nd);
  }
  return i;
}

#ifndef JWC_FWD_MINB
#define JWC_FWD_MINB 2
#endif
#ifndef JWC_INV_MINB
#define JWC_INV_MINB 2
#endif
#ifndef JWC_FWD_MAXT
#define JWC_FWD_MAXT 256        // launch bounds of the forward kernel (experiments: 512 x 2 for the short filters)
#define JWC_FWD_MINB_SHORT 3
#endif
template <int L, int R, bool WRAP>
__global__ void __launch_bounds__(JWC_FWD_MAXT, (L > 10 ? JWC_FWD_MINB : JWC_FWD_MINB_SHORT)) modwt_fwd_pass_kernel(const __grid_constant__ FwdPassArgs a,
                                                             const __grid_constant__ FilterPair f) {
  // shared memory is addressed as smem[int offset] everywhere: keeps the accesses plain LDS/STS with register offsets
  // Long filters run ONE instantiation (WRAP = true) for every shape and test RowMap::wrap at run time: with both
  // store paths in the item loop ptxas settles on 64 registers and spreads the 46 shared loads of an item evenly between
  // its 560 DFMAs, where the WRAP = false code takes 124 registers and front-loads 29 of them -- measured on
  // Daubechies20 J = 8: forward 27.5 -> 26.2 ms.  (The inverse kernel and the short filters lose 3-4 % with the
  // same trick and keep their compile-time split.)
  constexpr bool RT = (L > kUniformTapsMax);
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int P = 1 << a.logP;
  // passes on >= 4 phases (all strides >= 4 doubles = one 32-byte sector per row) store W straight from registers;
  // only the single-phase pass needs the shared-memory staging tiles to make its stride-1/2 levels coalesced
  const bool direct_w = (a.logP >= 2);
  const int wtile = direct_w ? 0 : P * a.T2;
  const int oV1 = a.vcap, oW0 = 2 * a.vcap, oW1 = 2 * a.vcap + wtile;
  const int oT = oW1 + wtile;                      // tap copy (kTapDoubles), then the mbarrier
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + oT + kTapDoubles);
  if (L > kUniformTapsMax) {
    for (int t = threadIdx.x; t < JWC_MAX_TAPS; t += blockDim.x) {
      smem[oT + t] = f.f0[t];
      smem[oT + JWC_MAX_TAPS + t] = f.f1[t];
    }
  }

  // CTA index taken apart without divisions (nblocks < 2^31): multiply-high by tiles_i's magic number, shifts for the groups
  unsigned bid = fastdiv(blockIdx.x, a.tiles_div);
  const int ti = (int)(blockIdx.x - bid * (unsigned)a.tiles_i);
  const int pg = (int)(bid & (unsigned)(a.groups - 1));
  const int64_t b = (int64_t)(bid >> a.log_groups);
  const int64_t i0 = (int64_t)ti * a.T2;
  const int tlen2 = (int)((a.Nd - i0 < a.T2) ? (a.Nd - i0) : a.T2);
  const int ph0 = pg << a.logP;
  const int64_t S0 = (int64_t)1 << a.j0;
  const double* in_b = a.in + b * a.in_sig;
  double* co_b = a.coeffs + b * a.coeff_sig;
  double* vo_b = a.vout + b * a.vout_sig;
  const int rows_in = a.Hp + tlen2;

  // ---- load V_{j0}: decimated rows i0-Hp .. i0+tlen2-1 (circular), P phases each ---------------------------
  if (a.mode == MODE_BULK) {
    if (tid == 0) {
      ptx::mbar_init(bar, 1);
      ptx::fence_mbar_init();
      const int64_t g0 = i0 - a.Hp;  // even; Hp <= N guaranteed by the planner
      ptx::mbar_expect_tx(bar, (uint32_t)rows_in * 8u);
      if (g0 >= 0) {
        ptx::bulk_g2s(smem, in_b + g0, (uint32_t)rows_in * 8u, bar);
      } else {
        const uint32_t head = (uint32_t)(-g0);
        ptx::bulk_g2s(smem, in_b + (a.N + g0), head * 8u, bar);
        ptx::bulk_g2s(smem + head, in_b, ((uint32_t)rows_in - head) * 8u, bar);
      }
      if (a.pf_dist > 0 && blockIdx.x + (unsigned)a.pf_dist < a.nblocks) {
        // roughly one wave ahead: the CTA that will own tile blockIdx + pf_dist finds its input in L2
        const unsigned nb = blockIdx.x + (unsigned)a.pf_dist;
        const unsigned q2 = fastdiv(nb, a.tiles_div);
        const int ti2 = (int)(nb - q2 * (unsigned)a.tiles_i);
        const int64_t b2 = (int64_t)q2;   // groups == 1 in bulk mode
        const int64_t i2 = (int64_t)ti2 * a.T2;
        const int64_t l2 = (a.Nd - i2 < a.T2) ? (a.Nd - i2) : a.T2;
        ptx::bulk_prefetch_l2(a.in + b2 * a.in_sig + i2, (uint32_t)l2 * 8u);
      }
      ptx::mbar_wait(bar, 0);   // one thread sleeps on the mbarrier, the CTA sleeps on the hardware barrier
    }
    __syncthreads();
    ptx::mbar_wait(bar, 0);     // already complete: every thread takes its own acquire on the TMA-written tile
  } else if (a.mode == MODE_VEC2) {
    const int hp2 = P >> 1;  // 16-byte chunks per row
    const int chunks = rows_in * hp2;
    for (int q = tid; q < chunks; q += nt) {
      const int r = q >> (a.logP - 1), pp = (q & (hp2 - 1)) * 2;
      const int64_t i = wrap_row(i0 - a.Hp + r, a.Nd);
      ptx::cp_async16(smem + r * P + pp, in_b + a.rm.template at<WRAP, RT>(i, S0) + ph0 + pp);
    }
    ptx::cp_async_commit_wait_all();
    __syncthreads();
  } else {
    const int total = rows_in * P;
    for (int e = tid; e < total; e += nt) {
      const int r = e >> a.logP, p = e & (P - 1);
      const int64_t i = wrap_row(i0 - a.Hp + r, a.Nd);
      smem[e] = in_b[a.rm.template at<WRAP, RT>(i, S0) + ph0 + p];
    }
    __syncthreads();
  }

  // ---- k fused levels ---------------------------------------------------------------------------------------------
  const double* ctaps = const_taps(f);
  const int eW = P * a.Hp;  // first virtual position whose W / final V is an output of this tile
  for (int jj = 1; jj <= a.k; jj++) {
    const int sh = a.logP + jj - 1;
    const int s = 1 << sh;
    const int hrem = (L - 1) * ((1 << a.k) - (1 << jj));
    const int e0 = P * (a.Hp - hrem);
    const int len = P * (hrem + tlen2);
    const int oin = (jj & 1) ? 0 : oV1;
    const int oout = (jj & 1) ? oV1 : 0;
    const int owst = ((jj & 1) ? oW0 : oW1) - eW;   // W staging indexed by virtual position
    const int rows = (len + s - 1) >> sh;
    const int nrb = (rows + R - 1) / R;
    const int items = nrb << sh;
    const int span = (R - 1) << sh;
    // direct_w: W row of this level at the phase origin (gw) and, when 2^j0 divides N, at the tile origin (gwt)
    double* gw = co_b + (int64_t)(a.j0 + jj - 1) * a.N + ph0;
    const bool plain = !WRAP || (RT && !a.rm.wrap);
    double* gwt = gw + (plain ? i0 * S0 : 0);
    const int64_t gstep = ((int64_t)s >> a.logP) * S0;          // global distance of two rows of an item (2^j0 | N)
  #pragma unroll 1
  for (int w = tid; w < items; w += nt) {
      const int rb = w >> sh, c = w & (s - 1);
      const int rel0 = ((rb * R) << sh) + c;
      const int ef = e0 + rel0;
      const bool full = rel0 + span < len;
      double av[R], aw[R];
      if (ef + span < eW) {           // item entirely inside the halo: only V is needed by the next level
        fwd_item<L, R, false>(smem + oin + ef + span, s, f, ctaps, av, aw);
        if (full) {
#pragma unroll
          for (int q = 0; q < R; q++) smem[oout + ef + (q << sh)] = av[q];
        } else {
#pragma unroll
          for (int q = 0; q < R; q++)
            if (rel0 + (q << sh) < len) smem[oout + ef + (q << sh)] = av[q];
        }
      } else {
        fwd_item<L, R, true>(smem + oin + ef + span, s, f, ctaps, av, aw);
        if (direct_w) {
          // virtual position e -> row r = (e - eW) >> logP, phase p = (e - eW) & (P - 1); rows of one item are s/P apart
          const int v0 = ef - eW;   // may be negative for an item that starts in the halo
          if (plain && full && v0 >= 0) {
            // the common case, no predicates: 32-bit element offset from the tile origin (tile-local, < N < 2^31)
            double* gp = gwt + ((unsigned)(v0 >> a.logP) * (unsigned)S0 + (unsigned)(v0 & (P - 1)));
            int so = oout + ef;
#pragma unroll
            for (int q = 0; q < R; q++) {
              smem[so] = av[q];
              *gp = aw[q];
              so += s;
              gp += gstep;
            }
          } else if (plain) {
            double* g0 = gwt + ((int64_t)(v0 >> a.logP)) * S0 + (v0 & (P - 1));
#pragma unroll
            for (int q = 0; q < R; q++) {
              const int e = ef + (q << sh);
              if (full || rel0 + (q << sh) < len) {
                smem[oout + e] = av[q];
                if (e >= eW) g0[q * gstep] = aw[q];
              }
            }
          } else {   // rows addressed along the cycles of the circular signal
#pragma unroll
            for (int q = 0; q < R; q++) {
              const int e = ef + (q << sh);
              if (full || rel0 + (q << sh) < len) {
                smem[oout + e] = av[q];
                if (e >= eW) gw[a.rm.template at<true>(i0 + ((e - eW) >> a.logP), S0) + ((e - eW) & (P - 1))] = aw[q];
              }
            }
          }
        } else if (full && ef >= eW) {
#pragma unroll
          for (int q = 0; q < R; q++) {
            smem[oout + ef + (q << sh)] = av[q];
            smem[owst + ef + (q << sh)] = aw[q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < R; q++) {
            const int e = ef + (q << sh);
            if (rel0 + (q << sh) < len) {
              smem[oout + e] = av[q];
              if (e >= eW) smem[owst + e] = aw[q];
            }
          }
        }
      }
    }
    // ---- W_{j0+jj} tile leaves; the last level also ships V ----------------------------------------------------------
    const int64_t wrow = (int64_t)(a.j0 + jj - 1) * a.N;
    const double* wst = smem + owst + eW;
    const double* vres = smem + oout + eW;
    if (a.mode == MODE_BULK) {
      ptx::fence_proxy_async();
      if (tid == 0) ptx::bulk_wait_read<0>();  // the store issued one level ago has finished reading its staging tile
      __syncthreads();
      if (tid == 0) {
        ptx::bulk_s2g(co_b + wrow + i0, wst, (uint32_t)tlen2 * 8u);
        if (jj == a.k) ptx::bulk_s2g(vo_b + i0, vres, (uint32_t)tlen2 * 8u);
        ptx::bulk_commit();
      }
    } else {
      __syncthreads();
      const int nv = (jj == a.k) ? 2 : 1;
      for (int v = direct_w ? 1 : 0; v < nv; v++) {
        const double* src = v ? vres : wst;
        double* dst = v ? vo_b : (co_b + wrow);
        if (a.mode == MODE_VEC2) {
          const int hp2 = P >> 1, chunks = tlen2 * hp2;
          for (int q = tid; q < chunks; q += nt) {
            const int r = q >> (a.logP - 1), pp = (q & (hp2 - 1)) * 2;
            const double2 val = *reinterpret_cast<const double2*>(src + r * P + pp);
            *reinterpret_cast<double2*>(dst + a.rm.template at<WRAP, RT>(i0 + r, S0) + ph0 + pp) = val;
          }
        } else {
          const int total = tlen2 * P;
          for (int e = tid; e < total; e += nt) {
            const int r = e >> a.logP, p = e & (P - 1);
            dst[a.rm.template at<WRAP, RT>(i0 + r, S0) + ph0 + p] = src[e];
          }
        }
      }
    }
  }
  if (a.mode == MODE_BULK && tid == 0) ptx::bulk_wait_read<0>();
}

template <int L>
int launch_fwd_pass(jwc_ctx* ctx, cudaStream_t st, const FwdPassArgs& a, const FilterPair& f, int threads, size_t smem,
                    int64_t nblocks) {
  if (a.rm.wrap || L > kUniformTapsMax) {   // 2^j0 does not divide N (rows addressed along the cycles of the circular
                                            // signal), and every shape of the long filters (see the kernel)
    auto kern = modwt_fwd_pass_kernel<L, kModwtR, true>;
    JWC_CUDA_CHECK(allow_max_dynamic_smem(kern));
    kern<<<(unsigned)nblocks, threads, smem, st>>>(a, f);
  } else if constexpr (L <= kUniformTapsMax) {
    auto kern = modwt_fwd_pass_kernel<L, kModwtR, false>;
    JWC_CUDA_CHECK(allow_max_dynamic_smem(kern));
    kern<<<(unsigned)nblocks, threads, smem, st>>>(a, f);
  }
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

int dispatch_fwd_pass(jwc_ctx* ctx, cudaStream_t st, const FwdPassArgs& a, const FilterPair& f, int L, int threads,
                      size_t smem, int64_t nblocks) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_fwd_pass<LL>(ctx, st, a, f, threads, smem, nblocks);
    JWC_ALL_L(JWC_CASE)
#undef JWC_CASE
    default: return JWC_ERR_UNSUPPORTED;
  }
}

}  // namespace

int fast_modwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_x, double* d_coeffs,
                       int64_t batch, int64_t n, int levels, const FilterPair& f, int L, int64_t x_sig) {
  if (x_sig <= 0) x_sig = n;
  if (L < 2 || L > 40 || (L & 1)) return JWC_ERR_UNSUPPORTED;
  if (levels > 30 || n >= ((int64_t)1 << 31)) return JWC_ERR_UNSUPPORTED;
  ModwtPlanInput pin{};
  pin.n = n; pin.J = levels; pin.L = L;
  pin.aligned16 = ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_coeffs)) & 15) == 0 && (x_sig & 1) == 0;
  // measured on B200 (C2): forward best with 2 large CTAs per SM (113 KB), inverse best with 3 (75 KB)
  // (fp64-bound long filters prefer 3 CTAs per SM: db20 J8 34.9 ms at 75 KB vs 37.3 ms at 113 KB)
  pin.smem_budget = ctx->tune.modwt_smem > 0 ? ctx->tune.modwt_smem : (L <= 10 ? 113000 : 75776);
  if (pin.smem_budget > dev.max_smem_optin) pin.smem_budget = dev.max_smem_optin;
  pin.tile_override = ctx->tune.modwt_tile; pin.group_override = ctx->tune.modwt_group;
  pin.threads_override = ctx->tune.modwt_threads_fwd > 0 ? ctx->tune.modwt_threads_fwd : ctx->tune.modwt_threads;
  pin.logp_override = ctx->tune.modwt_logp; pin.tile_deep_override = ctx->tune.modwt_tile_deep;
  pin.plan_override = ctx->tune.modwt_plan_fwd;
  pin.inverse = false;
  const ModwtPlan plan = modwt_plan(pin);
  if (plan.passes.empty()) return JWC_ERR_UNSUPPORTED;
  debug_plan("modwt forward", plan, n, levels, L);

  Scratch ws(ctx, dev, st);
  const int64_t cs = (int64_t)(levels + 1) * n;
  const int npass = (int)plan.passes.size();
  const bool need_scratch = !(plan.all_fused && npass == 1);
  double* vbuf[2] = {nullptr, nullptr};
  if (need_scratch) {
    vbuf[0] = ws.get((size_t)batch * n);
    if (npass >= 2 || !plan.all_fused) vbuf[1] = ws.get((size_t)batch * n);
    if (!vbuf[0] || !vbuf[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* vin = d_x;
  int64_t vin_sig = x_sig;
  for (int pi = 0; pi < npass; pi++) {
    const ModwtPass& p = plan.passes[pi];
    const bool last = plan.all_fused && pi == npass - 1;
    FwdPassArgs a{};
    a.in = vin; a.in_sig = vin_sig;
    a.coeffs = d_coeffs; a.coeff_sig = cs;
    if (last) { a.vout = d_coeffs + (int64_t)levels * n; a.vout_sig = cs; }
    else { a.vout = vbuf[pi & 1]; a.vout_sig = n; }
    const int64_t cycles = modwt_cycles(n, p.j0);   // = 2^j0 when that divides n
    a.N = n; a.Nd = n / cycles;
    a.j0 = p.j0; a.k = p.k; a.logP = p.logP; a.T2 = p.T2; a.Hp = p.Hp;
    a.tiles_i = (int)((a.Nd + p.T2 - 1) / p.T2);
    a.groups = (int)(cycles >> p.logP);
    a.tiles_div = make_fastdiv((unsigned)a.tiles_i);
    a.log_groups = 0;
    while ((1 << a.log_groups) < a.groups) a.log_groups++;
    a.rm.S0 = (int64_t)1 << p.j0; a.rm.N = n; a.rm.Sm = a.rm.S0 % n; a.rm.invN = 1.0 / (double)n;
    a.rm.wrap = (n % a.rm.S0) != 0 || (ctx->tune.modwt_force_wrap > 0 && p.j0 > 0);
    a.vcap = p.vcap; a.mode = p.mode;
    const int64_t nblocks = (int64_t)a.tiles_i * a.groups * batch;
    if (nblocks > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
    a.nblocks = (unsigned)nblocks;
    a.pf_dist = (p.mode == MODE_BULK) ? prefetch_distance(ctx, dev, p.smem, p.threads) : 0;
    int rc = dispatch_fwd_pass(ctx, st, a, f, L, p.threads, p.smem, nblocks);
    if (rc != JWC_OK) return rc;
    vin = a.vout; vin_sig = a.vout_sig;
  }
  if (!plan.all_fused) {
    // remaining levels on the per-level kernels, continuing from V_{generic_from}
    return generic_modwt_forward_from(ctx, dev, st, vin, vin_sig, plan.generic_from + 1, d_coeffs, batch, n, levels, f, L);
  }
  return JWC_OK;
}

// ==================================================================================================================
// inverse: levels j0+k down to j0+1 (reference: transforms/MODWTTransform.java:355-372 + adjoint convolution :703-716)
//     V_{j-1}[t] = sum_m g~[m] V_j[(t + m 2^(j-1)) mod N]  +  sum_m h~[m] W_j[(t + m 2^(j-1)) mod N]
// Same column scheme, mirrored: the halo is on the right, (L-1)(2^jj - 1) decimated samples for the level-jj inputs.
// The k+1 input tiles are read from HBM once (bulk TMA copies, W_{jj-1} prefetched into the second W buffer while level
// jj is computed), V ping-pongs in shared memory, one bulk store of the V_{j0} tile at the end: 8(k+2) B/sample.
// ==================================================================================================================
namespace {

struct InvPassArgs {
  const double* vin;     // V_{j0+k}
  const double* coeffs;  // coefficient block base (rows W_1 .. W_J, V_J)
  double* vout;          // V_{j0}
  int64_t vin_sig, coeff_sig, vout_sig;
  int64_t N, Nd;
  int j0, k, logP, T2, Hp, tiles_i, groups, vcap, mode;
  FastDiv tiles_div;   // division by tiles_i
  int log_groups;      // groups is a power of two (cycles and phases per CTA are)
  int pf_dist;
  RowMap rm;
  int top_barrier;   // 1 = round-1 form of the tile wait (one thread on the mbarrier, the rest on a block barrier)
  unsigned nblocks;
};

// -DJWC_INV_ONE_SUM=1: experiment of round 2 (profiles/r2_sweeps.txt, calls r7a / r7b), NOT the shipped form: 84 instead of
// 96 registers for the 40-tap filter, no gain with 128 or 256 threads, -4 % with 9-row items; the default keeps the
// reference's two sums.
#ifndef JWC_INV_ONE_SUM
#define JWC_INV_ONE_SUM 0
#endif
template <int L, int R>
__device__ __forceinline__ void inv_item(const double* __restrict__ pv, const double* __restrict__ pw, int s,
                                         const FilterPair& f, const double* __restrict__ taps, double (&out)[R]) {
  constexpr bool ST = (L > kUniformTapsMax);
  // ONE: both sums run through one accumulator per output (R instead of 2R live accumulators; the result differs from
  // the reference's "two sums added at the end" in the last bits only, far inside the 1e-12 parity bound)
  constexpr bool ONE = (JWC_INV_ONE_SUM != 0) && ST;
  double ag[R], ah[ONE ? 1 : R];
#pragma unroll
  for (int q = 0; q < R; q++) {
    ag[q] = 0.0;
    if constexpr (!ONE) ah[q] = 0.0;
  }
  double tg[L], th[L];
  // rows 0 .. R+L-2 ascending: accumulator q meets tap m = i - q in ascending m (reference order)
#pragma unroll
  for (int i = 0; i < R + L - 1; ++i) {
    if (ST && i < L) {   // tap i is first needed at row i (by accumulator 0)
      tg[i] = taps[i];
      th[i] = taps[JWC_MAX_TAPS + i];
    }
    const double xv = *pv, xw = *pw;
    pv += s;
    pw += s;
    if (ONE) {
#pragma unroll
      for (int q = 0; q < R; q++) {
        const int m = i - q;
        if (m >= 0 && m < L) ag[q] = fma(xv, tg[m], ag[q]);
      }
#pragma unroll
      for (int q = 0; q < R; q++) {
        const int m = i - q;
        if (m >= 0 && m < L) ag[q] = fma(xw, th[m], ag[q]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < R; q++) {
        const int m = i - q;
        if (m >= 0 && m < L) {
          ag[q] = fma(xv, ST ? tg[m] : f.f0[m], ag[q]);
          ah[q] = fma(xw, ST ? th[m] : f.f1[m], ah[q]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) {
    if constexpr (ONE) out[q] = ag[q];
    else out[q] = ag[q] + ah[q];   // :366-369: the two sums are added at the end
  }
}

// rows [i_start, i_start + rows) of the decimated index (circular), P phases each, into dst (virtual layout r*P + p)
template <bool WRAP>
__device__ __forceinline__ void inv_issue_load(const InvPassArgs& a, double* dst, const double* src_b, int64_t i_start,
                                               int rows, int ph0, uint64_t* bar, int tid, int nt) {
  const int P = 1 << a.logP;
  const int64_t S0 = (int64_t)1 << a.j0;
  if (a.mode == MODE_BULK) {
    if (tid == 0) {
      // caller has armed `bar` with expect_tx for the total; split at the wrap point (single wrap: rows <= N)
      const int64_t first = (a.N - i_start < rows) ? (a.N - i_start) : rows;
      ptx::bulk_g2s(dst, src_b + i_start, (uint32_t)first * 8u, bar);
      if (first < rows) ptx::bulk_g2s(dst + first, src_b, (uint32_t)(rows - first) * 8u, bar);
    }
  } else if (a.mode == MODE_VEC2) {
    const int hp2 = P >> 1, chunks = rows * hp2;
    for (int q = tid; q < chunks; q += nt) {
      const int r = q >> (a.logP - 1), pp = (q & (hp2 - 1)) * 2;
      const int64_t i = wrap_row(i_start + r, a.Nd);
      ptx::cp_async16(dst + r * P + pp, src_b + a.rm.template at<WRAP>(i, S0) + ph0 + pp);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  } else {
    const int total = rows * P;
    for (int e = tid; e < total; e += nt) {
      const int r = e >> a.logP, p = e & (P - 1);
      const int64_t i = wrap_row(i_start + r, a.Nd);
      dst[e] = src_b[a.rm.template at<WRAP>(i, S0) + ph0 + p];
    }
  }
}

// one level of a tile: every work item = R consecutive rows of one column (see the file header)
template <int L, int R>
__device__ __forceinline__ void inv_level(double* smem, int ovin, int owin, int ovout, int sh, int len, int rows,
                                          const FilterPair& f, const double* __restrict__ ctaps, int tid, int nt) {
  const int s = 1 << sh;
  const int nrb = (rows + R - 1) / R;
  const int items = nrb << sh;
  const int span = (R - 1) << sh;
#pragma unroll 1
  for (int w = tid; w < items; w += nt) {
    const int rb = w >> sh, c = w & (s - 1);
    const int rel0 = ((rb * R) << sh) + c;
    double o[R];
    inv_item<L, R>(smem + ovin + rel0, smem + owin + rel0, s, f, ctaps, o);
    if (rel0 + span < len) {
#pragma unroll
      for (int q = 0; q < R; q++) smem[ovout + rel0 + (q << sh)] = o[q];
    } else {
#pragma unroll
      for (int q = 0; q < R; q++)
        if (rel0 + (q << sh) < len) smem[ovout + rel0 + (q << sh)] = o[q];
    }
  }
}

template <int L, int R, bool WRAP>
__global__ void __launch_bounds__(256, (L > 10 ? JWC_INV_MINB : 3)) modwt_inv_pass_kernel(const __grid_constant__ InvPassArgs a,
                                                             const __grid_constant__ FilterPair f) {
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int P = 1 << a.logP;
  auto Vb = [&](int i) { return smem + (i & 1) * a.vcap; };          // always "smem + int": plain LDS/STS addressing
  auto Wb = [&](int i) { return smem + (2 + (i & 1)) * a.vcap; };
  const bool bulk = (a.mode == MODE_BULK);
  const int oT = 4 * a.vcap;                       // tap copy (kTapDoubles), then the two mbarriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oT + kTapDoubles);
  if (L > kUniformTapsMax) {
    for (int t = threadIdx.x; t < JWC_MAX_TAPS; t += blockDim.x) {
      smem[oT + t] = f.f0[t];
      smem[oT + JWC_MAX_TAPS + t] = f.f1[t];
    }
  }

  // CTA index taken apart without divisions (nblocks < 2^31): multiply-high by tiles_i's magic number, shifts for the groups
  unsigned bid = fastdiv(blockIdx.x, a.tiles_div);
  const int ti = (int)(blockIdx.x - bid * (unsigned)a.tiles_i);
  const int pg = (int)(bid & (unsigned)(a.groups - 1));
  const int64_t b = (int64_t)(bid >> a.log_groups);
  const int64_t i0 = (int64_t)ti * a.T2;
  const int tlen2 = (int)((a.Nd - i0 < a.T2) ? (a.Nd - i0) : a.T2);
  const int ph0 = pg << a.logP;
  const int64_t S0 = (int64_t)1 << a.j0;
  const double* vin_b = a.vin + b * a.vin_sig;
  const double* co_b = a.coeffs + b * a.coeff_sig;
  double* vo_b = a.vout + b * a.vout_sig;
  // halo (decimated) needed by the level-jj inputs; in bulk mode rounded up to an even row count
  auto halo = [&](int jj) { const int h = (L - 1) * ((1 << jj) - 1); return bulk ? h + (h & 1) : h; };

  if (bulk) {
    if (tid == 0) {
      ptx::mbar_init(&bars[0], 1);
      ptx::mbar_init(&bars[1], 1);
      ptx::fence_mbar_init();
      // V_{j0+k} and W_{j0+k} are requested right here, before the block barrier that publishes the mbarriers
      const int rows = tlen2 + halo(a.k);
      ptx::mbar_expect_tx(&bars[0], 2u * (uint32_t)rows * 8u);
      inv_issue_load<WRAP>(a, Vb(0), vin_b, i0, rows, ph0, &bars[0], tid, nt);
      inv_issue_load<WRAP>(a, Wb(0), co_b + (int64_t)(a.j0 + a.k - 1) * a.N, i0, rows, ph0, &bars[0], tid, nt);
    }
    __syncthreads();
  }
  if (bulk && a.pf_dist > 0 && tid <= a.k && blockIdx.x + (unsigned)a.pf_dist < a.nblocks) {
    // one wave ahead: thread t pulls the W_{j0+t} tile (t = 1..k) resp. the V tile (t = 0) of CTA blockIdx + pf_dist into L2
    const unsigned nb = blockIdx.x + (unsigned)a.pf_dist;
    const unsigned q2 = fastdiv(nb, a.tiles_div);
    const int ti2 = (int)(nb - q2 * (unsigned)a.tiles_i);
    const int64_t b2 = (int64_t)q2;
    const int64_t i2 = (int64_t)ti2 * a.T2;
    const int64_t l2 = (a.Nd - i2 < a.T2) ? (a.Nd - i2) : a.T2;
    const double* src = (tid == 0) ? (a.vin + b2 * a.vin_sig) : (a.coeffs + b2 * a.coeff_sig + (int64_t)(a.j0 + tid - 1) * a.N);
    ptx::bulk_prefetch_l2(src + i2, (uint32_t)l2 * 8u);
  }
  // prologue: V_{j0+k} and W_{j0+k} (bulk mode: already requested above)
  if (!bulk) {
    const int rows = tlen2 + halo(a.k);
    inv_issue_load<WRAP>(a, Vb(0), vin_b, i0, rows, ph0, &bars[0], tid, nt);
    inv_issue_load<WRAP>(a, Wb(0), co_b + (int64_t)(a.j0 + a.k - 1) * a.N, i0, rows, ph0, &bars[0], tid, nt);
  }
  const double* ctaps = const_taps(f);
  for (int jj = a.k, u = 0; jj >= 1; --jj, ++u) {
    const int wb = u & 1;
    if (jj > 1) {  // prefetch W_{j0+jj-1} into the other W buffer (last read two barriers ago)
      const int rows = tlen2 + halo(jj - 1);
      if (bulk && tid == 0) ptx::mbar_expect_tx(&bars[wb ^ 1], (uint32_t)rows * 8u);
      inv_issue_load<WRAP>(a, Wb(wb ^ 1), co_b + (int64_t)(a.j0 + jj - 2) * a.N, i0, rows, ph0, &bars[wb ^ 1], tid, nt);
    }
    if (bulk) {
      const uint32_t par = (uint32_t)((u >> 1) & 1);
      if (a.top_barrier >= 16) {
        ptx::mbar_wait_hint(&bars[wb], par, (uint32_t)a.top_barrier);   // experiment: suspend-time hint in ns
      } else if (!a.top_barrier) {
        // every thread takes its own acquire on the mbarrier: the block barrier at the end of the previous level has
        // already ordered the shared-memory traffic, so the top of a level needs no second one (measured on C2:
        // 3.50 -> 3.33 ms against one sleeper on the mbarrier + a block barrier for the rest)
        ptx::mbar_wait(&bars[wb], par);
      } else {
        if (tid == 0) ptx::mbar_wait(&bars[wb], par);   // one sleeper on the mbarrier, the rest on the hardware barrier
        __syncthreads();
        ptx::mbar_wait(&bars[wb], par);                 // complete already: per-thread acquire of the TMA-written tiles
      }
    } else {
      if (a.mode == MODE_VEC2) {
        if (jj > 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
    }
    const int sh = a.logP + jj - 1;
    const int s = 1 << sh;
    const int hout = (L - 1) * ((1 << (jj - 1)) - 1);   // halo the NEXT level still needs
    const int len = P * (tlen2 + hout);
    const int ovin = (u & 1) * a.vcap, owin = (2 + wb) * a.vcap, ovout = ((u + 1) & 1) * a.vcap;
    const int rows = (len + s - 1) >> sh;
    inv_level<L, R>(smem, ovin, owin, ovout, sh, len, rows, f, ctaps, tid, nt);
    if (bulk && jj == 1) ptx::fence_proxy_async();
    __syncthreads();
  }
  // ---- V_{j0} tile leaves ---------------------------------------------------------------------------------------------
  const double* res = Vb(a.k);
  if (bulk) {
    if (tid == 0) {
      ptx::bulk_s2g(vo_b + i0, res, (uint32_t)tlen2 * 8u);
      ptx::bulk_commit();
      ptx::bulk_wait_read<0>();
    }
  } else if (a.mode == MODE_VEC2) {
    const int hp2 = P >> 1, chunks = tlen2 * hp2;
    for (int q = tid; q < chunks; q += nt) {
      const int r = q >> (a.logP - 1), pp = (q & (hp2 - 1)) * 2;
      *reinterpret_cast<double2*>(vo_b + a.rm.template at<WRAP>(i0 + r, S0) + ph0 + pp) = *reinterpret_cast<const double2*>(res + r * P + pp);
    }
  } else {
    const int total = tlen2 * P;
    for (int e = tid; e < total; e += nt) {
      const int r = e >> a.logP, p = e & (P - 1);
      vo_b[a.rm.template at<WRAP>(i0 + r, S0) + ph0 + p] = res[e];
    }
  }
}

template <int L>
int launch_inv_pass(jwc_ctx* ctx, cudaStream_t st, const InvPassArgs& a, const FilterPair& f, int threads, size_t smem,
                    int64_t nblocks) {
  if (a.rm.wrap) {
    auto kern = modwt_inv_pass_kernel<L, kModwtR, true>;
    JWC_CUDA_CHECK(allow_max_dynamic_smem(kern));
    kern<<<(unsigned)nblocks, threads, smem, st>>>(a, f);
  } else {
    auto kern = modwt_inv_pass_kernel<L, kModwtR, false>;
    JWC_CUDA_CHECK(allow_max_dynamic_smem(kern));
    kern<<<(unsigned)nblocks, threads, smem, st>>>(a, f);
  }
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

int dispatch_inv_pass(jwc_ctx* ctx, cudaStream_t st, const InvPassArgs& a, const FilterPair& f, int L, int threads,
                      size_t smem, int64_t nblocks) {
  switch (L) {
#define JWC_CASE(LL) case LL: return launch_inv_pass<LL>(ctx, st, a, f, threads, smem, nblocks);
    JWC_ALL_L(JWC_CASE)
#undef JWC_CASE
    default: return JWC_ERR_UNSUPPORTED;
  }
}

}  // namespace

int fast_modwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_coeffs, double* d_x,
                       int64_t batch, int64_t n, int levels, const FilterPair& f, int L) {
  if (L < 2 || L > 40 || (L & 1)) return JWC_ERR_UNSUPPORTED;
  if (levels > 30 || n >= ((int64_t)1 << 31)) return JWC_ERR_UNSUPPORTED;
  ModwtPlanInput pin{};
  pin.n = n; pin.J = levels; pin.L = L;
  pin.aligned16 = ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_coeffs)) & 15) == 0;
  pin.smem_budget = ctx->tune.modwt_smem > 0 ? ctx->tune.modwt_smem : 75776;
  if (pin.smem_budget > dev.max_smem_optin) pin.smem_budget = dev.max_smem_optin;
  pin.tile_override = ctx->tune.modwt_tile; pin.group_override = ctx->tune.modwt_group;
  pin.threads_override = ctx->tune.modwt_threads;
  pin.logp_override = ctx->tune.modwt_logp; pin.tile_deep_override = ctx->tune.modwt_tile_deep;
  pin.plan_override = ctx->tune.modwt_plan_inv;
  pin.inverse = true;
  const ModwtPlan plan = modwt_plan(pin);
  // the inverse starts at the deepest level: it can only be fused if the whole chain is (no generic head)
  if (plan.passes.empty() || !plan.all_fused) return JWC_ERR_UNSUPPORTED;
  debug_plan("modwt inverse", plan, n, levels, L);

  Scratch ws(ctx, dev, st);
  const int64_t cs = (int64_t)(levels + 1) * n;
  const int npass = (int)plan.passes.size();
  double* vbuf[2] = {nullptr, nullptr};
  if (npass >= 2) {
    vbuf[0] = ws.get((size_t)batch * n);
    if (npass >= 3) vbuf[1] = ws.get((size_t)batch * n);
    if (!vbuf[0] || (npass >= 3 && !vbuf[1])) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* vin = d_coeffs + (int64_t)levels * n;
  int64_t vin_sig = cs;
  for (int pi = npass - 1, step = 0; pi >= 0; --pi, ++step) {   // deepest pass first
    const ModwtPass& p = plan.passes[pi];
    InvPassArgs a{};
    a.vin = vin; a.vin_sig = vin_sig;
    a.coeffs = d_coeffs; a.coeff_sig = cs;
    if (pi == 0) { a.vout = d_x; a.vout_sig = n; }
    else { a.vout = vbuf[step & 1]; a.vout_sig = n; }
    const int64_t cycles = modwt_cycles(n, p.j0);   // = 2^j0 when that divides n
    a.N = n; a.Nd = n / cycles;
    a.j0 = p.j0; a.k = p.k; a.logP = p.logP; a.T2 = p.T2; a.Hp = p.Hp;
    a.tiles_i = (int)((a.Nd + p.T2 - 1) / p.T2);
    a.groups = (int)(cycles >> p.logP);
    a.tiles_div = make_fastdiv((unsigned)a.tiles_i);
    a.log_groups = 0;
    while ((1 << a.log_groups) < a.groups) a.log_groups++;
    a.rm.S0 = (int64_t)1 << p.j0; a.rm.N = n; a.rm.Sm = a.rm.S0 % n; a.rm.invN = 1.0 / (double)n;
    a.rm.wrap = (n % a.rm.S0) != 0 || (ctx->tune.modwt_force_wrap > 0 && p.j0 > 0);
    a.vcap = p.vcap; a.mode = p.mode;
    const int64_t nblocks = (int64_t)a.tiles_i * a.groups * batch;
    if (nblocks > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
    a.nblocks = (unsigned)nblocks;
    a.pf_dist = (p.mode == MODE_BULK && ctx->tune.l2_prefetch > 0) ? ctx->tune.l2_prefetch : 0;   // off by default: measured slower
    a.top_barrier = ctx->tune.top_barrier;
    int rc = dispatch_inv_pass(ctx, st, a, f, L, p.threads, p.smem, nblocks);
    if (rc != JWC_OK) return rc;
    vin = a.vout; vin_sig = a.vout_sig;
  }
  return JWC_OK;
}

}  // namespace jwc
