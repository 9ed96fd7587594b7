// jwc_modwt_small.cu -- forward MODWT of SHORT signals (analysis windows): the whole signal lives in shared memory.
//
// The tile kernels of jwc_modwt_fast.cu carry a halo of (L-1)(2^J - 1) samples; for a 512-sample window at J = 8 that
// halo is 1785 samples (db4) -- 3.5x the signal, all of it recomputed.  A short signal needs no halo at all: the CTA
// keeps V_(j-1) of its signals in shared memory, indexes it circularly (any n, not only 2^p) and writes W_j straight
// to its row; HBM traffic is the algorithmic minimum (n read, (J+1) n written) and the arithmetic is exactly
// 2 L J n multiply-adds.  128 threads per CTA, S = 4, 2 or 1 signals per CTA (32..128 threads per signal).
//
// Arithmetic: MODWTTransform.java:290-304 + circularConvolve :677-690,
//   W_j[t] = sum_m h[m] V_(j-1)[(t - m 2^(j-1)) mod n],  V_j likewise with g (m ascending, FMA).
#include <algorithm>

#include "jwc_internal.cuh"
#include "jwc_tma.cuh"

namespace jwc {

namespace {

constexpr int kThreads = 128;

struct SmallArgs {
  const double* x;
  double* coeffs;
  int64_t x_sig;     // distance between consecutive input signals (windows may overlap)
  int64_t batch;
  int n, J, L;
  int per_cta;       // signals per CTA
  double* abs_parts; // fused magnitude: sum |coefficient| over everything this CTA writes goes to abs_parts[blockIdx.x]
};

constexpr int kChain = 5;   // outputs per work item of the register-blocked path; ODD: chains start 5 strides apart, so a
                            // warp's shared loads fall into distinct banks (4 gave 4-way conflicts, ncu: 94 M conflicts for
                            // 16 M wavefronts); 5 rather than 7 keeps the short classes of the coarse levels >= 80 % full

// LT = compile-time filter length (taps become constant-bank operands, the chain loop unrolls), 0 = run-time length.
// Register-blocked path (LT > 0, n divisible by the stride): a thread produces the kChain outputs
// t0, t0 + s, ..., t0 + (kChain-1) s, which share their inputs -- kChain + L - 1 shared-memory loads instead of
// kChain * L.  Other shapes (any n, any L) take the one-output-at-a-time loop.
// ABS: the magnitude sum of compressions/CompressorMagnitude.java:78-90 (sum of |c| over all coefficients) is taken in
// the store epilogue -- every W_j[t] and the final V_J[t] adds its absolute value to a per-thread sum as it leaves the
// registers -- so the thresholding that follows needs one select pass only (no extra read of the coefficients).
template <int LT, bool ABS = false>
__global__ void __launch_bounds__(kThreads) modwt_small_fwd_kernel(const __grid_constant__ SmallArgs a,
                                                                   const __grid_constant__ FilterPair f) {
  extern __shared__ double sm[];
  double asum = 0.0;
  const int tps = kThreads / a.per_cta;            // threads per signal
  const int s_local = threadIdx.x / tps, r = threadIdx.x % tps;
  const int64_t sig = (int64_t)blockIdx.x * a.per_cta + s_local;
  const bool live = sig < a.batch;
  const int n = a.n;
  double* cur = sm + (size_t)s_local * 2 * n;
  double* nxt = cur + n;
  if (live) {
    const double* x = a.x + sig * a.x_sig;
    for (int t = r; t < n; t += tps) ptx::cp_async8(cur + t, x + t);   // all pieces in flight together
  }
  ptx::cp_async_commit_wait_all();
  __syncthreads();
  double* co = a.coeffs + (live ? sig : 0) * (int64_t)(a.J + 1) * n;
  for (int j = 1; j <= a.J; j++) {
    // stride 2^(j-1) mod n without overflow (n < 2^31, j <= 30)
    const int step = (int)((((int64_t)1) << (j - 1)) % n);
    const bool last = (j == a.J);
    if (live) {
      double* w_row = co + (int64_t)(j - 1) * n;
      double* v_row = co + (int64_t)a.J * n;
      bool done = false;
      if constexpr (LT > 0) {
        if (step > 0 && n % step == 0) {
          // residue class a0 (mod step) holds n / step outputs a0 + b step; chains of kChain consecutive b, the last
          // one of a class may be partial (its surplus outputs are computed from wrapped inputs and dropped)
          const int per_class = n / step, chains = (per_class + kChain - 1) / kChain;
          const int items = step * chains;
          int back = ((LT - 1) * step) % n;          // inputs start (L-1) strides before the first output
          // step = 2^(j-1) mod n is a power of two whenever 2^(j-1) < n: shift / mask instead of two integer
          // divisions per item (a ~20-instruction sequence each, next to 2 * kChain * L DFMAs)
          const bool p2 = (step & (step - 1)) == 0;
          const int lg = 31 - __clz(step);
          for (int it = r; it < items; it += tps) {
            const int a0 = p2 ? (it & (step - 1)) : it % step, b0 = (p2 ? (it >> lg) : it / step) * kChain;
            const int t0 = a0 + b0 * step;
            int idx = t0 - back;
            if (idx < 0) idx += n;
            double aw[kChain], av[kChain];
#pragma unroll
            for (int q = 0; q < kChain; q++) aw[q] = av[q] = 0.0;
#pragma unroll
            for (int u = -(LT - 1); u <= kChain - 1; u++) {
              const double v = cur[idx];
#pragma unroll
              for (int q = 0; q < kChain; q++) {
                const int m = q - u;                 // output t0 + q s takes input t0 + u s with tap m
                if (m >= 0 && m < LT) {
                  aw[q] = fma(v, f.f1[m], aw[q]);
                  av[q] = fma(v, f.f0[m], av[q]);
                }
              }
              idx += step;
              if (idx >= n) idx -= n;
            }
#pragma unroll
            for (int q = 0; q < kChain; q++) {
              if (b0 + q < per_class) {
                const int t = t0 + q * step;
                w_row[t] = aw[q];
                if (ABS) asum += fabs(aw[q]);
                if (last) { v_row[t] = av[q]; if (ABS) asum += fabs(av[q]); }
                else nxt[t] = av[q];
              }
            }
          }
          done = true;
        }
      }
      if (!done) {
        for (int t = r; t < n; t += tps) {
          double sw = 0.0, sv = 0.0;
          int idx = t;
          for (int m = 0; m < a.L; m++) {
            const double v = cur[idx];
            sw = fma(v, f.f1[m], sw);
            sv = fma(v, f.f0[m], sv);
            idx -= step;
            if (idx < 0) idx += n;
          }
          w_row[t] = sw;
          if (ABS) asum += fabs(sw);
          if (last) { v_row[t] = sv; if (ABS) asum += fabs(sv); }
          else nxt[t] = sv;
        }
      }
    }
    __syncthreads();
    double* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (ABS) {   // fixed-shape reduction: lanes, then the four warps -> abs_parts[blockIdx.x] (no atomics: deterministic)
    __shared__ double red[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) asum += __shfl_down_sync(0xffffffffu, asum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = asum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < kThreads / 32; w++) tot += red[w];
      a.abs_parts[blockIdx.x] = tot;
    }
  }
}

// Inverse: V_(j-1)[t] = sum_m g[m] V_j[(t + m s) mod n] + sum_m h[m] W_j[(t + m s) mod n]  (MODWTTransform.java:355-372,
// :703-716), j = J .. 1.  V ping-pongs in shared memory, W_j streams into one of two more buffers (cp.async) while the
// level before it computes.
template <int LT>
__global__ void __launch_bounds__(kThreads) modwt_small_inv_kernel(const __grid_constant__ SmallArgs a,
                                                                   const __grid_constant__ FilterPair f) {
  extern __shared__ double sm[];
  const int tps = kThreads / a.per_cta;
  const int s_local = threadIdx.x / tps, r = threadIdx.x % tps;
  const int64_t sig = (int64_t)blockIdx.x * a.per_cta + s_local;
  const bool live = sig < a.batch;
  const int n = a.n;
  double* cur = sm + (size_t)s_local * 4 * n;
  double* nxt = cur + n;
  double* wbuf[2] = {nxt + n, nxt + 2 * n};
  const double* co = a.x + (live ? sig : 0) * (int64_t)(a.J + 1) * n;   // a.x = coefficient array here
  double* out = a.coeffs + (live ? sig : 0) * a.x_sig;                    // a.coeffs = reconstructed signals, stride x_sig
  // rows are 16-byte aligned when n is even and the coefficient block is: 16-byte cp.async (half the LDGSTS count)
  const bool v16 = ((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
  auto load_row = [&](double* dst, const double* src) {
    if (v16) for (int t = 2 * r; t < n; t += 2 * tps) ptx::cp_async16(dst + t, src + t);
    else for (int t = r; t < n; t += tps) ptx::cp_async8(dst + t, src + t);
  };
  if (live) {
    load_row(cur, co + (int64_t)a.J * n);
    load_row(wbuf[0], co + (int64_t)(a.J - 1) * n);
  }
  ptx::cp_async_commit();
  int wi = 0;
  for (int j = a.J; j >= 1; j--, wi ^= 1) {
    // W of the NEXT level streams in while this level computes
    if (live && j > 1) load_row(wbuf[wi ^ 1], co + (int64_t)(j - 2) * n);
    ptx::cp_async_commit();
    ptx::cp_async_wait<1>();
    __syncthreads();
    const double* wb = wbuf[wi];
    const int step = (int)((((int64_t)1) << (j - 1)) % n);
    const bool last = (j == 1);
    if (live) {
      bool done = false;
      if constexpr (LT > 0) {
        if (step > 0 && n % step == 0) {
          const int per_class = n / step, chains = (per_class + kChain - 1) / kChain;
          const int items = step * chains;
          const bool p2 = (step & (step - 1)) == 0;   // see the forward kernel
          const int lg = 31 - __clz(step);
          for (int it = r; it < items; it += tps) {
            const int a0 = p2 ? (it & (step - 1)) : it % step, b0 = (p2 ? (it >> lg) : it / step) * kChain;
            const int t0 = a0 + b0 * step;
            int idx = t0;
            double acc[kChain];
#pragma unroll
            for (int q = 0; q < kChain; q++) acc[q] = 0.0;
#pragma unroll
            for (int u = 0; u <= kChain + LT - 2; u++) {
              const double v = cur[idx], w = wb[idx];
#pragma unroll
              for (int q = 0; q < kChain; q++) {
                const int m = u - q;                 // output t0 + q s takes input t0 + u s with tap m
                if (m >= 0 && m < LT) acc[q] = fma(w, f.f1[m], fma(v, f.f0[m], acc[q]));
              }
              idx += step;
              if (idx >= n) idx -= n;
            }
#pragma unroll
            for (int q = 0; q < kChain; q++) {
              if (b0 + q < per_class) {
                const int t = t0 + q * step;
                if (last) out[t] = acc[q];
                else nxt[t] = acc[q];
              }
            }
          }
          done = true;
        }
      }
      if (!done) {
        for (int t = r; t < n; t += tps) {
          double sa = 0.0, sd = 0.0;
          int idx = t;
          for (int m = 0; m < a.L; m++) {
            sa = fma(cur[idx], f.f0[m], sa);
            sd = fma(wb[idx], f.f1[m], sd);
            idx += step;
            if (idx >= n) idx -= n;
          }
          if (last) out[t] = sa + sd;
          else nxt[t] = sa + sd;
        }
      }
    }
    __syncthreads();
    double* tmp = cur; cur = nxt; nxt = tmp;
  }
}

}  // namespace

// Returns JWC_ERR_UNSUPPORTED when the shape is not one this kernel is meant for.
int small_modwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_x, double* d_coeffs,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L, int64_t x_sig, AbsSum* abs) {
  (void)dev;
  if (x_sig <= 0) x_sig = n;
  const int mode = ctx->tune.modwt_small;
  if (mode < 0 || n > 2048 || n < 1 || levels > 30 || L < 1) return JWC_ERR_UNSUPPORTED;
  // worth it once the tile kernels' halo is comparable to the signal itself
  const int64_t halo = (int64_t)(L - 1) * ((((int64_t)1) << levels) - 1);
  if (mode == 0 && 3 * halo <= n) return JWC_ERR_UNSUPPORTED;   // measured: Haar J8 on 512 (halo 255) small 1.98 ms vs 2.46; db4 J3 (halo 49) 1.83 vs 1.16
  SmallArgs a{};
  a.x = d_x; a.coeffs = d_coeffs; a.x_sig = x_sig; a.batch = batch; a.n = (int)n; a.J = levels; a.L = L;
  a.per_cta = n <= 512 ? 4 : (n <= 1024 ? 2 : 1);
  if (ctx->tune.small_per_cta == 1 || ctx->tune.small_per_cta == 2 || ctx->tune.small_per_cta == 4)
    a.per_cta = std::min(a.per_cta, ctx->tune.small_per_cta);
  const int64_t ctas = (batch + a.per_cta - 1) / a.per_cta;
  if (ctas > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
  const size_t smem = (size_t)a.per_cta * 2 * (size_t)n * sizeof(double);
  if (abs) {   // magnitude sum fused into the stores: one partial per CTA
    abs->parts = abs->ws->get((size_t)ctas);
    if (!abs->parts) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    abs->nparts = ctas;
    abs->fused = true;
    a.abs_parts = abs->parts;
  }
  switch (L) {
#define JWC_SCASE(LL)                                                                          \
  case LL:                                                                                     \
    if (abs) modwt_small_fwd_kernel<LL, true><<<(unsigned)ctas, kThreads, smem, st>>>(a, f);   \
    else modwt_small_fwd_kernel<LL, false><<<(unsigned)ctas, kThreads, smem, st>>>(a, f);      \
    break;
    JWC_SCASE(2) JWC_SCASE(4) JWC_SCASE(6) JWC_SCASE(8) JWC_SCASE(10) JWC_SCASE(12) JWC_SCASE(14) JWC_SCASE(16)
    JWC_SCASE(18) JWC_SCASE(20)
#undef JWC_SCASE
    default:
      if (abs) modwt_small_fwd_kernel<0, true><<<(unsigned)ctas, kThreads, smem, st>>>(a, f);
      else modwt_small_fwd_kernel<0, false><<<(unsigned)ctas, kThreads, smem, st>>>(a, f);
      break;
  }
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

}  // namespace jwc

namespace jwc {

int small_modwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_coeffs, double* d_x,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L) {
  (void)dev;
  const int mode = ctx->tune.modwt_small;
  if (mode < 0 || n > 2048 || n < 1 || levels > 30 || L < 1) return JWC_ERR_UNSUPPORTED;
  const int64_t halo = (int64_t)(L - 1) * ((((int64_t)1) << levels) - 1);
  if (mode == 0 && 3 * halo <= n) return JWC_ERR_UNSUPPORTED;
  SmallArgs a{};
  a.x = d_coeffs; a.coeffs = d_x; a.x_sig = n; a.batch = batch; a.n = (int)n; a.J = levels; a.L = L;
  // two signals per CTA (64 threads each) even for the shortest windows: the inverse keeps four buffers per signal, so
  // four signals per CTA means 64 KB and three CTAs (12 warps) per SM; measured on 512-sample windows: 3.72 -> 3.32 ms
  a.per_cta = n <= 1024 ? 2 : 1;
  if (ctx->tune.small_per_cta == 1 || ctx->tune.small_per_cta == 2 || ctx->tune.small_per_cta == 4)
    a.per_cta = std::min(a.per_cta, ctx->tune.small_per_cta);
  const int64_t ctas = (batch + a.per_cta - 1) / a.per_cta;
  if (ctas > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
  const size_t smem = (size_t)a.per_cta * 4 * (size_t)n * sizeof(double);   // 64 KB: above the 48 KB default
  switch (L) {
#define JWC_SCASE(LL)                                                                                               \
  case LL:                                                                                                          \
    JWC_CUDA_CHECK(allow_max_dynamic_smem(modwt_small_inv_kernel<LL>)); \
    modwt_small_inv_kernel<LL><<<(unsigned)ctas, kThreads, smem, st>>>(a, f);                                         \
    break;
    JWC_SCASE(2) JWC_SCASE(4) JWC_SCASE(6) JWC_SCASE(8) JWC_SCASE(10) JWC_SCASE(12) JWC_SCASE(14) JWC_SCASE(16)
    JWC_SCASE(18) JWC_SCASE(20)
#undef JWC_SCASE
    default:
      JWC_CUDA_CHECK(allow_max_dynamic_smem(modwt_small_inv_kernel<0>));
      modwt_small_inv_kernel<0><<<(unsigned)ctas, kThreads, smem, st>>>(a, f);
      break;
  }
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

}  // namespace jwc
