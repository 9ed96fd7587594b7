// jwc_dwt_whole.cu -- FWT of signals of 1024..4096 samples (image rows, analysis windows): the whole signal in shared
// memory, every level in place, one launch.
//
// For short signals the tile kernels of jwc_dwt_fast.cu need a second launch with one tiny CTA per signal for the deep
// levels and their inverse reads the detail blocks level by level.  Here a CTA owns one signal: it loads the n
// samples (forward) or the n coefficients (inverse) once, coalesced, runs all levels IN PLACE in shared memory -- the
// pyramid layout [A_J | D_J | ... | D_1] is exactly what in-place analysis of a shrinking prefix produces, and what
// in-place synthesis of a growing prefix consumes -- and stores the n results once, coalesced.  HBM traffic is the
// algorithmic 16 B/sample; no halo, no scratch.
//
// Work split inside a level: a thread owns R consecutive outputs (analysis: R low/high pairs from 2R + L - 2 inputs,
// read as R + L/2 - 1 double2; synthesis: R output pairs from R + L/2 - 1 rows of each half).  R is odd, so the
// threads' windows start R double2 (or R doubles) apart and a warp's shared loads fall into distinct banks.  Results
// stay in registers across the barrier that separates the level's reads from its in-place writes.
//
// Arithmetic: Wavelet.java:236-260 / :277-303 (fused multiply-add), loops FastWaveletTransform.java:85-99 / :133-151.
#include "jwc_internal.cuh"
#include "jwc_tma.cuh"

namespace jwc {

namespace {

struct WholeArgs {
  const double* src;
  const double* prefix;   // inverse only: A at the deepest level of this launch when it comes from another buffer
  double* dst;
  int64_t src_sig, dst_sig, prefix_sig;
  int prefix_len;
  int n, steps;
  int vec;   // 1: rows are 16-byte aligned (double2 copies in and out)
  int tree;  // 1: wavelet packets -- every block of h samples is transformed at each level, not only the first
};

// asynchronous copies: every thread's pieces are in flight together (a plain load/store loop pays one global-memory
// latency per iteration)
__device__ __forceinline__ void load_row(double* sm, const double* x, int n, int vec) {
  if (vec) {
    for (int t = 2 * threadIdx.x; t < n; t += 2 * blockDim.x) ptx::cp_async16(sm + t, x + t);
  } else {
    for (int t = threadIdx.x; t < n; t += blockDim.x) ptx::cp_async8(sm + t, x + t);
  }
  ptx::cp_async_commit_wait_all();
}
// the same, but samples [0, plen) come from `pre` (the approximation rebuilt by the tail kernel)
__device__ __forceinline__ void load_row_split(double* sm, const double* x, const double* pre, int plen, int n, int vec) {
  if (vec) {
    for (int t = 2 * threadIdx.x; t < n; t += 2 * blockDim.x) ptx::cp_async16(sm + t, (t < plen ? pre : x) + t);
  } else {
    for (int t = threadIdx.x; t < n; t += blockDim.x) ptx::cp_async8(sm + t, (t < plen ? pre : x) + t);
  }
  ptx::cp_async_commit_wait_all();
}
__device__ __forceinline__ void store_row(double* y, const double* sm, int n, int vec) {
  if (vec) {
    for (int t = 2 * threadIdx.x; t < n; t += 2 * blockDim.x)
      *reinterpret_cast<double2*>(y + t) = *reinterpret_cast<const double2*>(sm + t);
  } else {
    for (int t = threadIdx.x; t < n; t += blockDim.x) y[t] = sm[t];
  }
}

// One analysis level in place: every block of h samples (the first `blocks` of them) becomes [lo | hi] of itself.
// A work item = R consecutive outputs of one block; item -> (block, run) so that a level's items are dense in the
// thread index whatever the block length.
template <int L, int R>
__device__ __forceinline__ void fwd_level(double* sm, int h, int blocks, const FilterPair& f) {
  const int half = h >> 1, mask = h - 1;
  const int rpb = (half + R - 1) / R;                 // runs per block
  const int item = threadIdx.x;
  const bool active = item < blocks * rpb;
  const int p = item / rpb, i0 = (item - p * rpb) * R;
  double* blk = sm + p * h;
  double lo[R], hi[R];
  if (active) {
    constexpr int W = R + L / 2 - 1;
    double2 win[W];
#pragma unroll
    for (int t = 0; t < W; t++) win[t] = *reinterpret_cast<const double2*>(blk + ((2 * i0 + 2 * t) & mask));
#pragma unroll
    for (int q = 0; q < R; q++) lo[q] = hi[q] = 0.0;
#pragma unroll
    for (int m = 0; m < L; m++)
#pragma unroll
      for (int q = 0; q < R; q++) {
        const double v = (m & 1) ? win[q + m / 2].y : win[q + m / 2].x;   // x[2(i0+q) + m]
        lo[q] = fma(v, f.f0[m], lo[q]);
        hi[q] = fma(v, f.f1[m], hi[q]);
      }
  }
  __syncthreads();   // every read of this level is done: write [lo | hi] over the block
  if (active) {
#pragma unroll
    for (int q = 0; q < R; q++)
      if (i0 + q < half) {
        blk[i0 + q] = lo[q];
        blk[half + i0 + q] = hi[q];
      }
  }
  __syncthreads();
}

template <int L, int R>
__device__ __forceinline__ void inv_level(double* sm, int h, int blocks, const FilterPair& f) {
  constexpr int M = L / 2;
  const int half = h >> 1, mask = half - 1;
  const int rpb = (half + R - 1) / R;
  const int item = threadIdx.x;
  const bool active = item < blocks * rpb;
  const int p = item / rpb, q0 = (item - p * rpb) * R;
  double* blk = sm + p * h;
  double e[R], o[R];
  if (active) {
    constexpr int W = R + M - 1;
    double wl[W], wh[W];
#pragma unroll
    for (int t = 0; t < W; t++) {
      const int i = (q0 - (M - 1) + t) & mask;   // two's-complement mask = mod half (also when L/2 > half)
      wl[t] = blk[i];
      wh[t] = blk[half + i];
    }
#pragma unroll
    for (int u = 0; u < R; u++) e[u] = o[u] = 0.0;
#pragma unroll
    for (int m = M - 1; m >= 0; m--)
#pragma unroll
      for (int u = 0; u < R; u++) {
        const double cl = wl[u - m + M - 1], ch = wh[u - m + M - 1];
        e[u] = fma(ch, f.f1[2 * m], fma(cl, f.f0[2 * m], e[u]));
        o[u] = fma(ch, f.f1[2 * m + 1], fma(cl, f.f0[2 * m + 1], o[u]));
      }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int u = 0; u < R; u++)
      if (q0 + u < half) {
        blk[2 * (q0 + u)] = e[u];
        blk[2 * (q0 + u) + 1] = o[u];
      }
  }
  __syncthreads();
}

template <int L, int R>
__global__ void __launch_bounds__(R == 7 ? 320 : 416, R == 7 ? 4 : 3) whole_fwd_kernel(const __grid_constant__ WholeArgs a,
                                                        const __grid_constant__ FilterPair f) {
  extern __shared__ double sm[];
  const int n = a.n;
  load_row(sm, a.src + (int64_t)blockIdx.x * a.src_sig, n, a.vec);
  __syncthreads();
  int h = n;
  for (int lev = 0; lev < a.steps; lev++, h >>= 1) fwd_level<L, R>(sm, h, a.tree ? n / h : 1, f);
  store_row(a.dst + (int64_t)blockIdx.x * a.dst_sig, sm, n, a.vec);
}

// R = 7 launches at most 320 threads, R = 5 at most 416: four resp. three CTAs of 32 KB per SM
template <int L, int R>
__global__ void __launch_bounds__(R == 7 ? 320 : 416, R == 7 ? 4 : 3) whole_inv_kernel(const __grid_constant__ WholeArgs a,
                                                        const __grid_constant__ FilterPair f) {
  extern __shared__ double sm[];
  const int n = a.n;
  if (a.prefix)
    load_row_split(sm, a.src + (int64_t)blockIdx.x * a.src_sig, a.prefix + (int64_t)blockIdx.x * a.prefix_sig,
                   a.prefix_len, n, a.vec);
  else
    load_row(sm, a.src + (int64_t)blockIdx.x * a.src_sig, n, a.vec);
  __syncthreads();
  int h = n >> (a.steps - 1);
  for (int lev = 0; lev < a.steps; lev++, h <<= 1) inv_level<L, R>(sm, h, a.tree ? n / h : 1, f);
  store_row(a.dst + (int64_t)blockIdx.x * a.dst_sig, sm, n, a.vec);
}

// ---------------------------------------------------------------------------------------------------------------------
// The same in-place synthesis for LONG signals, tile by tile (pyramid inverse passes of k <= 3 levels).
//
// A tile of T = 4096 output samples needs A_(l0+k)[g0/2^k ..), D_(l0+k)[g0/2^k ..), ..., D_(l0+1)[g0/2 ..): laid out
// as [A | D_k | ... | D_1] in shared memory that IS the in-place pyramid of a T-sample signal, so the whole-signal
// level code runs on it unchanged.  Its periodic wrap is wrong only for the first c0 = 2 (L/2-1) (2^k - 1) outputs
// (c_(s-1) = 2 (c_s + L/2 - 1)); the tile therefore starts `pad` >= c0 samples early and throws that prefix away
// (2.5 % recomputation for L = 16, k = 3).  The loader wraps the segment reads mod the level lengths, which makes the
// first tile of a signal read the signal's tail -- the transform's own periodicity.
// ---------------------------------------------------------------------------------------------------------------------
struct TileArgs {
  const double* ain;    // A at depth l0 + k, h >> k samples per signal
  const double* din;    // the coefficient array (pyramid layout of the whole signal)
  double* out;          // A at depth l0, h samples per signal
  int64_t ain_sig, din_sig, out_sig;
  int64_t N;            // length of the whole signal (D_j lives at [N >> j, N >> (j-1)) of din)
  int64_t h;            // N >> l0
  int l0, k;
  int T, pad, valid;    // tile length, discarded prefix, outputs kept per tile (T - pad)
  int tiles;
};

template <int L, int R>
__global__ void __launch_bounds__(448) tile_inv_kernel(const __grid_constant__ TileArgs a,
                                                       const __grid_constant__ FilterPair f) {
  extern __shared__ double sm[];
  const int64_t b = blockIdx.x / a.tiles;
  const int t = blockIdx.x - (int)(b * a.tiles);
  const int64_t o0 = (int64_t)t * a.valid;       // first output this tile keeps
  const int64_t g0 = o0 - a.pad;                 // first output it computes (multiple of 2^k; negative for tile 0)
  const int T = a.T;
  // gather the k + 1 segments; every level length is a power of two, so wrapping is a mask
  {
    const double* ain = a.ain + b * a.ain_sig;
    const double* din = a.din + b * a.din_sig;
    const int seg = T >> a.k;
    const int64_t len = a.h >> a.k, start = g0 >> a.k;   // arithmetic shift: floor, also for negative g0
    for (int i = threadIdx.x; i < seg; i += blockDim.x) {
      const int64_t src = (start + i) & (len - 1);
      ptx::cp_async8(sm + i, ain + src);
      ptx::cp_async8(sm + seg + i, din + (a.N >> (a.l0 + a.k)) + src);
    }
    for (int s = a.k - 1; s >= 1; s--) {
      const int segs = T >> s;                            // D_(l0+s): T >> s samples at [T >> s, T >> (s-1))
      const int64_t lens = a.h >> s, starts = g0 >> s;
      const double* d = din + (a.N >> (a.l0 + s));
      for (int i = 2 * threadIdx.x; i < segs; i += 2 * blockDim.x)
        ptx::cp_async16(sm + segs + i, d + ((starts + i) & (lens - 1)));   // starts even, lens even: pairs stay together
    }
    ptx::cp_async_commit_wait_all();
  }
  __syncthreads();
  int h = T >> (a.k - 1);
  for (int lev = 0; lev < a.k; lev++, h <<= 1) inv_level<L, R>(sm, h, 1, f);
  double* out = a.out + b * a.out_sig + o0;
  int64_t keep = a.h - o0;
  if (keep > a.valid) keep = a.valid;
  for (int i = 2 * threadIdx.x; i < keep; i += 2 * blockDim.x)
    *reinterpret_cast<double2*>(out + i) = *reinterpret_cast<const double2*>(sm + a.pad + i);
}

template <int L>
int launch_tile_inv(jwc_ctx* ctx, cudaStream_t st, TileArgs a, const FilterPair& f, int64_t batch) {
  constexpr int R = (L <= 10) ? 7 : 5;
  a.T = 4096;
  const int c0 = 2 * (L / 2 - 1) * ((1 << a.k) - 1);
  a.pad = (c0 + 7) & ~7;                    // multiple of 2^k (k <= 3) and of two doubles
  a.valid = a.T - a.pad;
  a.tiles = (int)((a.h + a.valid - 1) / a.valid);
  const int64_t ctas = (int64_t)a.tiles * batch;
  if (ctas > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
  const int threads = ((a.T / 2 + R - 1) / R + 31) / 32 * 32;
  const size_t smem = (size_t)a.T * sizeof(double);
  tile_inv_kernel<L, R><<<(unsigned)ctas, threads, smem, st>>>(a, f);
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

template <int L>
int launch_whole(jwc_ctx* ctx, cudaStream_t st, const WholeArgs& a, const FilterPair& f, int64_t batch, bool inverse) {
  constexpr int R = (L <= 10) ? 7 : 5;
  int items = 0;   // one thread per work item of the busiest level
  for (int lev = 0, h = a.n; lev < a.steps; lev++, h >>= 1) {
    const int it = (a.tree ? a.n / h : 1) * ((h / 2 + R - 1) / R);
    if (it > items) items = it;
  }
  const int threads = (items + 31) / 32 * 32;
  if (threads > (R == 7 ? 320 : 416)) return JWC_ERR_UNSUPPORTED;
  const size_t smem = (size_t)a.n * sizeof(double);
  if (inverse) whole_inv_kernel<L, R><<<(unsigned)batch, threads, smem, st>>>(a, f);
  else         whole_fwd_kernel<L, R><<<(unsigned)batch, threads, smem, st>>>(a, f);
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

}  // namespace

// Levels the whole-signal kernel takes for a signal of length n with `steps` levels to do: all of them when the
// deepest block keeps >= 128 samples, otherwise the levels whose block is longer than kDwtTailLen (the rest goes to the
// warp-per-signal tail).  Every level costs the CTA two barriers -- fine for big levels, a loss for the deep end of a
// full-depth transform (measured, n = 4096: all 12 levels here 2.13 ms, 3 levels here + 9 in the tail ~1.5 ms).
// 0: this kernel does not apply.
int whole_dwt_levels(const jwc_ctx* ctx, int64_t n, int steps, int L, bool tree) {
  if (tree) {   // packets: every level is as big as the first, so all levels or nothing (launch_whole checks the fit)
    if (ctx->tune.dwt_whole < 0 || n < 64 || n > 4096 || (n & (n - 1)) || steps < 1 || L < 2 || L > 20 || (L & 1)) return 0;
    // measured on 131 072 rows of 4096 (this kernel vs the tile kernels): Haar 6 levels 2.82 vs 4.44 ms, but db4 3 levels
    // 2.26 vs 1.93 ms and sym8 3 levels 3.07 vs 2.69 ms -- a single fused tile pass beats it, several passes do not.
    // Default: short filters, deep transforms; anything else only when forced (dwt_whole = 1).
    if (ctx->tune.dwt_whole == 0 && !(L <= 4 && steps >= 4)) return 0;
    return steps;
  }
  if (ctx->tune.dwt_whole < 0 || n <= kDwtTailLen || n > 4096 || (n & (n - 1)) || steps < 1) return 0;
  if (L < 2 || L > 20 || (L & 1)) return 0;
  if (ctx->tune.dwt_whole > 0 || (n >> steps) >= 128) return steps;
  int top = 0;
  while ((n >> top) > kDwtTailLen) top++;
  return top < steps ? top : steps;
}

// `steps` levels of the FWT (pyramid) of batch signals of length n in (512, 4096] (see whole_dwt_levels).  Inverse with
// d_prefix != nullptr: the deepest approximation (n >> steps samples per signal, stride prefix_sig) is read from
// d_prefix instead of the head of d_in.
int whole_dwt(jwc_ctx* ctx, cudaStream_t st, const double* d_in, double* d_out, int64_t batch, int64_t n, int steps,
              const FilterPair& f, int L, int64_t ld, bool inverse, const double* d_prefix, int64_t prefix_sig, bool tree) {
  if (n < 64 || n > 4096 || (n & (n - 1)) || steps < 1 || ((int64_t)1 << steps) > n) return JWC_ERR_UNSUPPORTED;
  if (L < 2 || L > 20 || (L & 1) || batch > 0x7fffffffLL) return JWC_ERR_UNSUPPORTED;
  WholeArgs a{};
  a.src = d_in; a.dst = d_out; a.src_sig = ld; a.dst_sig = ld; a.n = (int)n; a.steps = steps; a.tree = tree ? 1 : 0;
  a.prefix = (inverse && !tree) ? d_prefix : nullptr; a.prefix_sig = prefix_sig; a.prefix_len = (int)(n >> steps);
  a.vec = (((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_prefix)) & 15) == 0 &&
           (ld & 1) == 0 && (prefix_sig & 1) == 0 && (a.prefix_len & 1) == 0) ? 1 : 0;
  switch (L) {
#define JWC_WCASE(LL) case LL: return launch_whole<LL>(ctx, st, a, f, batch, inverse);
    JWC_WCASE(2) JWC_WCASE(4) JWC_WCASE(6) JWC_WCASE(8) JWC_WCASE(10) JWC_WCASE(12) JWC_WCASE(14) JWC_WCASE(16)
    JWC_WCASE(18) JWC_WCASE(20)
#undef JWC_WCASE
    default: return JWC_ERR_UNSUPPORTED;
  }
}

// One pyramid-inverse pass of k <= 3 levels on long signals (h = N >> l0 >= 16384) with the tiled in-place kernel.
// Needs 16-byte aligned rows everywhere (the caller checks the strides it chose); JWC_ERR_UNSUPPORTED otherwise.
int tile_dwt_inverse_pass(jwc_ctx* ctx, cudaStream_t st, const double* ain, int64_t ain_sig, const double* din,
                          int64_t din_sig, double* out, int64_t out_sig, int64_t N, int l0, int k, int64_t batch,
                          const FilterPair& f, int L) {
  // Opt-in only (dwt_tile_inv = 1): measured on C3 (1024 x 2^20, 20 levels) it LOSES to the TMA tile kernels of
  // jwc_dwt_fast.cu -- Haar inverse 3.24 vs 3.11 ms, db8 5.66 vs 4.87 ms (element-wise cp.async gathers of four
  // segments and six block barriers per tile, against bulk copies with L2 prefetch).  Kept as the cross-check it is.
  if (ctx->tune.dwt_tile_inv <= 0 || L < 2 || L > 20 || (L & 1) || k < 1 || k > 3) return JWC_ERR_UNSUPPORTED;
  const int64_t h = N >> l0;
  if (h < 16384 || (h & (h - 1))) return JWC_ERR_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(ain) | reinterpret_cast<uintptr_t>(din) | reinterpret_cast<uintptr_t>(out)) & 15) ||
      ((ain_sig | din_sig | out_sig) & 1))
    return JWC_ERR_UNSUPPORTED;
  TileArgs a{};
  a.ain = ain; a.din = din; a.out = out; a.ain_sig = ain_sig; a.din_sig = din_sig; a.out_sig = out_sig;
  a.N = N; a.h = h; a.l0 = l0; a.k = k;
  switch (L) {
#define JWC_WCASE(LL) case LL: return launch_tile_inv<LL>(ctx, st, a, f, batch);
    JWC_WCASE(2) JWC_WCASE(4) JWC_WCASE(6) JWC_WCASE(8) JWC_WCASE(10) JWC_WCASE(12) JWC_WCASE(14) JWC_WCASE(16)
    JWC_WCASE(18) JWC_WCASE(20)
#undef JWC_WCASE
    default: return JWC_ERR_UNSUPPORTED;
  }
}

}  // namespace jwc
