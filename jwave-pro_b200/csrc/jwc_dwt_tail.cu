// jwc_dwt_tail.cu -- the deep end of the FWT pyramid for SHORT signals: one warp per signal, all remaining levels in
// one launch.
//
// The tile kernels of jwc_dwt_fast.cu give every (signal, tile) its own CTA.  For a large batch of short signals
// (rows of an image, analysis windows) the levels below ~256 samples are then a launch of one tiny CTA per signal --
// 131 072 CTAs that each move a few hundred bytes and spend their time in block barriers (measured: 0.5 - 0.8 ms per
// such pass on B200 while touching < 1 % of the data).  Here a warp keeps its signal's current approximation in its
// own slice of shared memory and walks the remaining levels with __syncwarp only; 4 signals per CTA.
//
// Arithmetic: Wavelet.java:236-260 (analysis) and :277-303 (synthesis, gather form), level loops
// FastWaveletTransform.java:85-99 / :133-151.  The block length h is a power of two, so `mod h` is a mask and blocks
// shorter than the filter wrap several times without special cases.
#include "jwc_internal.cuh"
#include "jwc_tma.cuh"

namespace jwc {

namespace {

constexpr int kWarps = 4;

struct TailArgs {
  const double* src;   // forward: A at the first tail level, signal stride src_sig; inverse: the coefficient array
  double* dst;         // forward: the coefficient array (stride n); inverse: A at the top of the tail (stride dst_sig)
  int64_t src_sig, dst_sig;
  int64_t n;           // length of the whole signal = stride of the coefficient array
  int64_t batch;
  int h0;              // block length at the top of the tail (power of two, <= kDwtTailLen)
  int nlev;            // levels done here
  int L;
};

// LT = compile-time filter length (taps become constant-bank operands, tap loops unroll); 0 = run-time length
template <int LT>
__global__ void __launch_bounds__(kWarps * 32) tail_fwd_kernel(const __grid_constant__ TailArgs a,
                                                                const __grid_constant__ FilterPair f) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t sig = (int64_t)blockIdx.x * kWarps + warp;
  if (sig >= a.batch) return;   // warps are independent: no block-wide barrier below
  double* cur = sm + (size_t)warp * ((a.h0 + a.h0 / 2 + 1) & ~1);   // even slice size: 16-byte loads stay aligned
  double* nxt = cur + a.h0;
  const double* x = a.src + sig * a.src_sig;
  for (int t = lane; t < a.h0; t += 32) ptx::cp_async8(cur + t, x + t);   // all pieces in flight together
  ptx::cp_async_commit_wait_all();
  __syncwarp();
  double* out = a.dst + sig * a.n;
  int h = a.h0;
  for (int lev = 0; lev < a.nlev; lev++) {
    const int half = h >> 1, mask = h - 1;
    const bool last = (lev == a.nlev - 1);
    for (int i = lane; i < half; i += 32) {
      double lo = 0.0, hi = 0.0;
      if constexpr (LT > 0) {
#pragma unroll
        for (int p = 0; p < LT / 2; p++) {   // x[2i+2p], x[2i+2p+1] as one 16-byte load (lanes 16 bytes apart)
          const double2 v = *reinterpret_cast<const double2*>(cur + ((2 * i + 2 * p) & mask));
          lo = fma(v.y, f.f0[2 * p + 1], fma(v.x, f.f0[2 * p], lo));
          hi = fma(v.y, f.f1[2 * p + 1], fma(v.x, f.f1[2 * p], hi));
        }
      } else {
        for (int j = 0; j < a.L; j++) {
          const double v = cur[(2 * i + j) & mask];
          lo = fma(v, f.f0[j], lo);
          hi = fma(v, f.f1[j], hi);
        }
      }
      out[half + i] = hi;
      if (last) out[i] = lo;
      else nxt[i] = lo;
    }
    __syncwarp();
    double* t = cur; cur = nxt; nxt = t;
    h = half;
  }
}

template <int LT>
__global__ void __launch_bounds__(kWarps * 32) tail_inv_kernel(const __grid_constant__ TailArgs a,
                                                                const __grid_constant__ FilterPair f) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t sig = (int64_t)blockIdx.x * kWarps + warp;
  if (sig >= a.batch) return;
  // the coefficients of all tail levels are one contiguous prefix [A_deep | D_deep | ... | D_top] of the signal's
  // array: fetch it once (one global-memory latency for the whole tail), then ping-pong the approximations
  double* coef = sm + (size_t)warp * 2 * a.h0;
  double* cur = coef + a.h0;
  double* nxt = cur + a.h0 / 2;
  const double* c = a.src + sig * a.n;
  double* dst = a.dst + sig * a.dst_sig;
  for (int t = lane; t < a.h0; t += 32) ptx::cp_async8(coef + t, c + t);
  ptx::cp_async_commit_wait_all();
  __syncwarp();
  int half = a.h0 >> a.nlev;                    // length of the deepest approximation (>= 1)
  const double* lo = coef;
  for (int lev = 0; lev < a.nlev; lev++) {
    const int h = half << 1, mask = half - 1;
    const bool last = (lev == a.nlev - 1);
    const double* hi = coef + half;
    if constexpr (LT > 0) {
      // one output PAIR (2q, 2q+1) per step: both use the coefficient rows q - m, m < L/2
      for (int q = lane; q < half; q += 32) {
        double e = 0.0, o = 0.0;
#pragma unroll
        for (int m = 0; m < LT / 2; m++) {
          const int i = (q - m) & mask;
          const double cl = lo[i], ch = hi[i];
          e = fma(ch, f.f1[2 * m], fma(cl, f.f0[2 * m], e));
          o = fma(ch, f.f1[2 * m + 1], fma(cl, f.f0[2 * m + 1], o));
        }
        if (last) { dst[2 * q] = e; dst[2 * q + 1] = o; }
        else { nxt[2 * q] = e; nxt[2 * q + 1] = o; }
      }
    } else {
      for (int k = lane; k < h; k += 32) {
        double acc = 0.0;
        for (int j = k & 1; j < a.L; j += 2) {
          const int i = ((k - j) >> 1) & mask;    // (2i + j) mod h == k
          acc = fma(hi[i], f.f1[j], fma(lo[i], f.f0[j], acc));
        }
        if (last) dst[k] = acc;
        else nxt[k] = acc;
      }
    }
    __syncwarp();
    lo = nxt;
    double* t = cur; cur = nxt; nxt = t;
    half = h;
  }
}

int launch_tail(jwc_ctx* ctx, cudaStream_t st, const TailArgs& a, const FilterPair& f, bool inverse) {
  const int64_t ctas = (a.batch + kWarps - 1) / kWarps;
  if (ctas <= 0) return JWC_OK;
  if (ctas > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
  const size_t smem = (size_t)kWarps * (inverse ? 2 * a.h0 : ((a.h0 + a.h0 / 2 + 1) & ~1)) * sizeof(double);
  switch ((a.L & 1) ? 0 : a.L) {
#define JWC_TCASE(LL)                                                                          \
  case LL:                                                                                     \
    if (inverse) tail_inv_kernel<LL><<<(unsigned)ctas, kWarps * 32, smem, st>>>(a, f);         \
    else         tail_fwd_kernel<LL><<<(unsigned)ctas, kWarps * 32, smem, st>>>(a, f);         \
    break;
    JWC_TCASE(2) JWC_TCASE(4) JWC_TCASE(6) JWC_TCASE(8) JWC_TCASE(10) JWC_TCASE(12) JWC_TCASE(14) JWC_TCASE(16)
    JWC_TCASE(18) JWC_TCASE(20)
#undef JWC_TCASE
    default:
      if (inverse) tail_inv_kernel<0><<<(unsigned)ctas, kWarps * 32, smem, st>>>(a, f);
      else         tail_fwd_kernel<0><<<(unsigned)ctas, kWarps * 32, smem, st>>>(a, f);
      break;
  }
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

}  // namespace

// First level of the pyramid that the tail kernels take over for a signal of length n with `steps` levels to do:
// the first level whose block is <= kDwtTailLen samples.  -1: no tail (long signals keep their deep levels inside the
// last tile pass; nothing to gain there).
int dwt_tail_start(int64_t n, int steps) {
  if (n > kDwtTailMaxN) return -1;
  int l = 0;
  while ((n >> l) > kDwtTailLen) l++;
  return l < steps ? l : -1;
}

int dwt_tail_forward(jwc_ctx* ctx, cudaStream_t st, const double* src, int64_t src_sig, double* d_out, int64_t n,
                     int h0, int nlev, int64_t batch, const FilterPair& f, int L) {
  TailArgs a{};
  a.src = src; a.src_sig = src_sig; a.dst = d_out; a.dst_sig = n; a.n = n; a.batch = batch; a.h0 = h0; a.nlev = nlev;
  a.L = L;
  return launch_tail(ctx, st, a, f, false);
}

int dwt_tail_inverse(jwc_ctx* ctx, cudaStream_t st, const double* d_in, int64_t n, double* dst, int64_t dst_sig, int h0,
                     int nlev, int64_t batch, const FilterPair& f, int L) {
  TailArgs a{};
  a.src = d_in; a.src_sig = n; a.dst = dst; a.dst_sig = dst_sig; a.n = n; a.batch = batch; a.h0 = h0; a.nlev = nlev;
  a.L = L;
  return launch_tail(ctx, st, a, f, true);
}

}  // namespace jwc
