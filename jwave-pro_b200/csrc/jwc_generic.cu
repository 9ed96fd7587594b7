// jwc_generic.cu -- shape-agnostic CUDA kernels, one transform level per launch.
//
// These cover EVERY shape the reference accepts (any n for MODWT, filter longer than the signal, blocks
// shorter than the filter, levels = 0 ...) and carry the EXACT mode (unfused mul/add in the reference's
// summation order => bit-identical to the JVM).  The fused tile kernels (jwc_modwt_fast.cu,
// jwc_dwt_fast.cu) take over for the large regular shapes; these remain the cross-check.
//
// Reference semantics (paths relative to /root/reference/src/main/java/jwave/):
//   MODWT level, forward   transforms/MODWTTransform.java:290-304 + circularConvolve :677-690
//   MODWT level, inverse   :355-372 + circularConvolveAdjoint :703-716 (two sums, then one add :366-369)
//   analysis step          transforms/wavelets/Wavelet.java:236-260
//   synthesis step         transforms/wavelets/Wavelet.java:277-303 (scatter-add; here in gather form, the
//                          contributions to one output are added in the same (i ascending, j ascending) order)
#include "jwc_internal.cuh"

namespace jwc {

namespace {

constexpr int kThreads = 256;

struct ModwtLevelArgs {
  const double* in_v;   // V_{j-1} (forward) / V_j (inverse)
  const double* in_w;   // inverse only: W_j
  double* out_v;        // V_j (forward) / V_{j-1} (inverse)
  double* out_w;        // forward only: W_j
  int64_t in_v_stride, in_w_stride, out_v_stride, out_w_stride;  // distance between signals, in doubles
  int64_t n, batch;
  int L;
  int64_t off[JWC_MAX_TAPS];  // (m * 2^(j-1)) mod n
};

template <bool EXACT>
__global__ void __launch_bounds__(kThreads) modwt_level_fwd_kernel(const __grid_constant__ ModwtLevelArgs a,
                                                                   const __grid_constant__ FilterPair f) {
  const int64_t total = a.n * a.batch;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
    const int64_t b = idx / a.n, t = idx - b * a.n;
    const double* x = a.in_v + b * a.in_v_stride;
    double sw = 0.0, sv = 0.0;
    for (int m = 0; m < a.L; m++) {
      int64_t i = t - a.off[m];
      if (i < 0) i += a.n;
      const double xv = x[i];
      sw = mac<EXACT>(sw, xv, f.f1[m]);
      sv = mac<EXACT>(sv, xv, f.f0[m]);
    }
    a.out_w[b * a.out_w_stride + t] = sw;
    a.out_v[b * a.out_v_stride + t] = sv;
  }
}

template <bool EXACT>
__global__ void __launch_bounds__(kThreads) modwt_level_inv_kernel(const __grid_constant__ ModwtLevelArgs a,
                                                                   const __grid_constant__ FilterPair f) {
  const int64_t total = a.n * a.batch;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
    const int64_t b = idx / a.n, t = idx - b * a.n;
    const double* v = a.in_v + b * a.in_v_stride;
    const double* w = a.in_w + b * a.in_w_stride;
    double sa = 0.0, sd = 0.0;
    for (int m = 0; m < a.L; m++) {
      int64_t i = t + a.off[m];
      if (i >= a.n) i -= a.n;
      sa = mac<EXACT>(sa, v[i], f.f0[m]);
      sd = mac<EXACT>(sd, w[i], f.f1[m]);
    }
    a.out_v[b * a.out_v_stride + t] = EXACT ? __dadd_rn(sa, sd) : sa + sd;
  }
}

struct DwtStepArgs {
  const double* src_lo;  // forward: the block to analyse; inverse: low-pass half of each block
  const double* src_hi;  // inverse only: high-pass half of each block
  double* dst_lo;        // forward: low-pass half; inverse: the synthesised block
  double* dst_hi;        // forward only
  int64_t src_lo_sig, src_hi_sig, dst_lo_sig, dst_hi_sig;  // signal strides
  int64_t src_lo_blk, src_hi_blk, dst_lo_blk, dst_hi_blk;  // block (packet) strides
  int64_t h;        // block length at this level
  int64_t blocks;   // blocks per signal (1 for FWT)
  int64_t batch;
  int L;
};

template <bool EXACT>
__global__ void __launch_bounds__(kThreads) dwt_step_fwd_kernel(const __grid_constant__ DwtStepArgs a,
                                                                const __grid_constant__ FilterPair f) {
  const int64_t half = a.h >> 1;
  const int64_t total = half * a.blocks * a.batch;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
    const int64_t i = idx % half;
    const int64_t pb = idx / half;
    const int64_t p = pb % a.blocks, b = pb / a.blocks;
    const double* x = a.src_lo + b * a.src_lo_sig + p * a.src_lo_blk;
    double lo = 0.0, hi = 0.0;
    for (int j = 0; j < a.L; j++) {
      int64_t k = 2 * i + j;
      if (k >= a.h) k %= a.h;
      const double xv = x[k];
      lo = mac<EXACT>(lo, xv, f.f0[j]);
      hi = mac<EXACT>(hi, xv, f.f1[j]);
    }
    a.dst_lo[b * a.dst_lo_sig + p * a.dst_lo_blk + i] = lo;
    a.dst_hi[b * a.dst_hi_sig + p * a.dst_hi_blk + i] = hi;
  }
}

template <bool EXACT>
__device__ __forceinline__ double synth_term(double acc, double clo, double s, double chi, double w) {
  if (EXACT) return __dadd_rn(acc, __dadd_rn(__dmul_rn(clo, s), __dmul_rn(chi, w)));  // Wavelet.java:294-296
  return fma(chi, w, fma(clo, s, acc));
}

template <bool EXACT>
__global__ void __launch_bounds__(kThreads) dwt_step_inv_kernel(const __grid_constant__ DwtStepArgs a,
                                                                const __grid_constant__ FilterPair f) {
  const int64_t h = a.h, half = h >> 1;
  const int64_t total = h * a.blocks * a.batch;
  const int L = a.L;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
    const int64_t k = idx % h;
    const int64_t pb = idx / h;
    const int64_t p = pb % a.blocks, b = pb / a.blocks;
    const double* lo = a.src_lo + b * a.src_lo_sig + p * a.src_lo_blk;
    const double* hi = a.src_hi + b * a.src_hi_sig + p * a.src_hi_blk;
    double acc = 0.0;
    if (h >= L) {
      // every i contributes at most one tap: first the unwrapped ones (j = k - 2i), then the wrapped (j = k + h - 2i)
      int64_t i0 = (k - L + 2) >> 1;  // ceil((k - L + 1) / 2), arithmetic shift handles negatives
      if (i0 < 0) i0 = 0;
      for (int64_t i = i0; 2 * i <= k; i++) {
        const int j = (int)(k - 2 * i);
        acc = synth_term<EXACT>(acc, lo[i], f.f0[j], hi[i], f.f1[j]);
      }
      for (int64_t i = (k + h - L + 2) >> 1; i < half; i++) {
        const int j = (int)(k + h - 2 * i);
        acc = synth_term<EXACT>(acc, lo[i], f.f0[j], hi[i], f.f1[j]);
      }
    } else {
      for (int64_t i = 0; i < half; i++) {
        int64_t j = (k - 2 * i) % h;
        if (j < 0) j += h;
        for (; j < L; j += h) acc = synth_term<EXACT>(acc, lo[i], f.f0[j], hi[i], f.f1[j]);
      }
    }
    a.dst_lo[b * a.dst_lo_sig + p * a.dst_lo_blk + k] = acc;
  }
}

__global__ void __launch_bounds__(kThreads) copy_rows_kernel(const double* __restrict__ src, double* __restrict__ dst,
                                                             int64_t src_stride, int64_t dst_stride, int64_t len,
                                                             int64_t batch) {
  const int64_t total = len * batch;
  for (int64_t idx = blockIdx.x * (int64_t)kThreads + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kThreads) {
    const int64_t b = idx / len, t = idx - b * len;
    dst[b * dst_stride + t] = src[b * src_stride + t];
  }
}

int grid_for(const DeviceSlot& dev, int64_t total) {
  int64_t blocks = (total + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)dev.sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int copy_rows(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* src, double* dst, int64_t src_stride,
              int64_t dst_stride, int64_t len, int64_t batch) {
  if (len <= 0 || batch <= 0) return JWC_OK;
  copy_rows_kernel<<<grid_for(dev, len * batch), kThreads, 0, st>>>(src, dst, src_stride, dst_stride, len, batch);
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

void fill_offsets(ModwtLevelArgs& a, int level, int L, int64_t n) {
  // (m * 2^(level-1)) mod n without overflow: stride mod n first
  const uint64_t s = ((uint64_t)1 << (level - 1)) % (uint64_t)n;
  for (int m = 0; m < L; m++) a.off[m] = (int64_t)((s * (uint64_t)m) % (uint64_t)n);  // n < 2^56 checked by the API
}

}  // namespace

int generic_modwt_forward_from(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_v, int64_t v_sig,
                               int first_level, double* d_coeffs, int64_t batch, int64_t n, int levels,
                               const FilterPair& f, int L, bool exact) {
  Scratch ws(ctx, dev, st);
  double* vbuf[2] = {nullptr, nullptr};
  const int todo = levels - first_level + 1;
  if (todo >= 2) {
    vbuf[0] = ws.get((size_t)batch * n);
    if (!vbuf[0]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  if (todo >= 3) {
    vbuf[1] = ws.get((size_t)batch * n);
    if (!vbuf[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const int64_t cs = (int64_t)(levels + 1) * n;
  const double* in = d_v;
  int64_t in_stride = v_sig;
  for (int j = first_level; j <= levels; j++) {
    ModwtLevelArgs a{};
    a.in_v = in;
    a.in_v_stride = in_stride;
    a.out_w = d_coeffs + (int64_t)(j - 1) * n;
    a.out_w_stride = cs;
    if (j == levels) {
      a.out_v = d_coeffs + (int64_t)levels * n;
      a.out_v_stride = cs;
    } else {
      a.out_v = vbuf[(j - first_level) & 1];
      a.out_v_stride = n;
    }
    a.n = n;
    a.batch = batch;
    a.L = L;
    fill_offsets(a, j, L, n);
    const int grid = grid_for(dev, n * batch);
    if (exact) modwt_level_fwd_kernel<true><<<grid, kThreads, 0, st>>>(a, f);
    else       modwt_level_fwd_kernel<false><<<grid, kThreads, 0, st>>>(a, f);
    count_launch(ctx);
    JWC_CUDA_CHECK(cudaGetLastError());
    in = a.out_v;
    in_stride = a.out_v_stride;
  }
  return JWC_OK;
}

int generic_modwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_x, double* d_coeffs,
                          int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool exact) {
  return generic_modwt_forward_from(ctx, dev, st, d_x, n, 1, d_coeffs, batch, n, levels, f, L, exact);
}

int generic_modwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_coeffs, double* d_x,
                          int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool exact) {
  Scratch ws(ctx, dev, st);
  double* vbuf[2] = {nullptr, nullptr};
  if (levels >= 2) {
    vbuf[0] = ws.get((size_t)batch * n);
    if (!vbuf[0]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  if (levels >= 3) {
    vbuf[1] = ws.get((size_t)batch * n);
    if (!vbuf[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const int64_t cs = (int64_t)(levels + 1) * n;
  const double* in_v = d_coeffs + (int64_t)levels * n;
  int64_t in_v_stride = cs;
  for (int j = levels; j >= 1; j--) {
    ModwtLevelArgs a{};
    a.in_v = in_v;
    a.in_v_stride = in_v_stride;
    a.in_w = d_coeffs + (int64_t)(j - 1) * n;
    a.in_w_stride = cs;
    if (j == 1) {
      a.out_v = d_x;
      a.out_v_stride = n;
    } else {
      a.out_v = vbuf[j & 1];
      a.out_v_stride = n;
    }
    a.n = n;
    a.batch = batch;
    a.L = L;
    fill_offsets(a, j, L, n);
    const int grid = grid_for(dev, n * batch);
    if (exact) modwt_level_inv_kernel<true><<<grid, kThreads, 0, st>>>(a, f);
    else       modwt_level_inv_kernel<false><<<grid, kThreads, 0, st>>>(a, f);
    count_launch(ctx);
    JWC_CUDA_CHECK(cudaGetLastError());
    in_v = a.out_v;
    in_v_stride = a.out_v_stride;
  }
  return JWC_OK;
}

// FWT: level l (0-based) analyses the prefix of length h = n >> l of A_l:  lo -> A_{l+1}, hi -> out[h/2 .. h) (final).
// WPT: level l analyses every block of length h; whole-array ping-pong.
int generic_dwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, bool exact,
                        int64_t ld) {
  if (ld <= 0) ld = n;
  auto sig_of = [&](const double* p) { return (p == d_in || p == d_out) ? ld : n; };   // scratch is dense
  int steps = 0;  // number of analysis steps actually performed (reference loop: while h >= 2 && l < level)
  for (int64_t h = n; h >= 2 && steps < levels; h >>= 1) steps++;
  if (steps == 0) return copy_rows(ctx, dev, st, d_in, d_out, ld, ld, n, batch);
  Scratch ws(ctx, dev, st);
  if (tree) {
    double* tmp = nullptr;
    if (steps >= 2) {
      tmp = ws.get((size_t)batch * n);
      if (!tmp) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    }
    const double* src = d_in;
    int64_t h = n;
    for (int l = 0; l < steps; l++, h >>= 1) {
      double* dst = (((steps - 1 - l) & 1) == 0) ? d_out : tmp;  // last step lands in d_out
      DwtStepArgs a{};
      a.src_lo = src; a.src_lo_sig = sig_of(src); a.src_lo_blk = h;
      a.dst_lo = dst; a.dst_lo_sig = sig_of(dst); a.dst_lo_blk = h;
      a.dst_hi = dst + (h >> 1); a.dst_hi_sig = sig_of(dst); a.dst_hi_blk = h;
      a.h = h; a.blocks = n / h; a.batch = batch; a.L = L;
      const int grid = grid_for(dev, (n >> 1) * batch);
      if (exact) dwt_step_fwd_kernel<true><<<grid, kThreads, 0, st>>>(a, f);
      else       dwt_step_fwd_kernel<false><<<grid, kThreads, 0, st>>>(a, f);
      count_launch(ctx);
      JWC_CUDA_CHECK(cudaGetLastError());
      src = dst;
    }
    return JWC_OK;
  }
  // pyramid: untouched tail of the reference's in-place array is simply D_1..D_l written at their final place;
  // approximations ping-pong through scratch (A_1 has n/2 samples, A_2 n/4, ...).
  double* abuf[2] = {nullptr, nullptr};
  if (steps >= 2) {
    abuf[0] = ws.get((size_t)batch * (n >> 1));
    if (!abuf[0]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  if (steps >= 3) {
    abuf[1] = ws.get((size_t)batch * (n >> 2));
    if (!abuf[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* src = d_in;
  int64_t src_sig = ld;
  int64_t h = n;
  for (int l = 0; l < steps; l++, h >>= 1) {
    DwtStepArgs a{};
    a.src_lo = src; a.src_lo_sig = src_sig; a.src_lo_blk = 0;
    if (l == steps - 1) { a.dst_lo = d_out; a.dst_lo_sig = ld; }
    else { a.dst_lo = abuf[l & 1]; a.dst_lo_sig = h >> 1; }
    a.dst_hi = d_out + (h >> 1); a.dst_hi_sig = ld;
    a.h = h; a.blocks = 1; a.batch = batch; a.L = L;
    const int grid = grid_for(dev, (h >> 1) * batch);
    if (exact) dwt_step_fwd_kernel<true><<<grid, kThreads, 0, st>>>(a, f);
    else       dwt_step_fwd_kernel<false><<<grid, kThreads, 0, st>>>(a, f);
    count_launch(ctx);
    JWC_CUDA_CHECK(cudaGetLastError());
    src = a.dst_lo;
    src_sig = a.dst_lo_sig;
  }
  return JWC_OK;
}

int generic_dwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, bool exact,
                        int64_t ld) {
  if (ld <= 0) ld = n;
  auto sig_of = [&](const double* p) { return (p == d_in || p == d_out) ? ld : n; };
  // reference: h starts at 2 << (log2 n - level) and doubles while h <= n  (FastWaveletTransform.java:137-151)
  int p = 0;
  while (((int64_t)1 << p) < n) p++;
  int64_t h0 = (int64_t)2 << (p - levels);
  int steps = 0;
  for (int64_t h = h0; h <= n && h >= 2; h <<= 1) steps++;
  if (steps == 0) return copy_rows(ctx, dev, st, d_in, d_out, ld, ld, n, batch);
  Scratch ws(ctx, dev, st);
  if (tree) {
    double* tmp = nullptr;
    if (steps >= 2) {
      tmp = ws.get((size_t)batch * n);
      if (!tmp) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    }
    const double* src = d_in;
    int64_t h = h0;
    for (int l = 0; l < steps; l++, h <<= 1) {
      double* dst = (((steps - 1 - l) & 1) == 0) ? d_out : tmp;
      DwtStepArgs a{};
      a.src_lo = src; a.src_lo_sig = sig_of(src); a.src_lo_blk = h;
      a.src_hi = src + (h >> 1); a.src_hi_sig = sig_of(src); a.src_hi_blk = h;
      a.dst_lo = dst; a.dst_lo_sig = sig_of(dst); a.dst_lo_blk = h;
      a.h = h; a.blocks = n / h; a.batch = batch; a.L = L;
      const int grid = grid_for(dev, n * batch);
      if (exact) dwt_step_inv_kernel<true><<<grid, kThreads, 0, st>>>(a, f);
      else       dwt_step_inv_kernel<false><<<grid, kThreads, 0, st>>>(a, f);
      count_launch(ctx);
      JWC_CUDA_CHECK(cudaGetLastError());
      src = dst;
    }
    return JWC_OK;
  }
  // pyramid: A_l of length h/2 (from d_in at the first step, scratch afterwards) + D_l = d_in[h/2 .. h) -> A_{l-1};
  // everything beyond the final prefix is copied through unchanged (the reference returns a copy of the input there,
  // but the final step always has h = n, so there is no untouched tail when steps > 0).
  double* abuf[2] = {nullptr, nullptr};
  if (steps >= 2) {
    abuf[0] = ws.get((size_t)batch * (n >> 1));
    if (!abuf[0]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  if (steps >= 3) {
    abuf[1] = ws.get((size_t)batch * (n >> 2));
    if (!abuf[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* src = d_in;
  int64_t src_sig = ld;
  int64_t h = h0;
  for (int l = 0; l < steps; l++, h <<= 1) {
    DwtStepArgs a{};
    a.src_lo = src; a.src_lo_sig = src_sig; a.src_lo_blk = 0;
    a.src_hi = d_in + (h >> 1); a.src_hi_sig = ld; a.src_hi_blk = 0;
    if (l == steps - 1) { a.dst_lo = d_out; a.dst_lo_sig = ld; }
    else {
      // ping-pong so that the larger buffer receives the larger result: remaining steps r = steps-1-l,
      // result length h = n >> r
      const int r = steps - 1 - l;
      a.dst_lo = abuf[(r - 1) & 1]; a.dst_lo_sig = h;
    }
    a.h = h; a.blocks = 1; a.batch = batch; a.L = L;
    const int grid = grid_for(dev, h * batch);
    if (exact) dwt_step_inv_kernel<true><<<grid, kThreads, 0, st>>>(a, f);
    else       dwt_step_inv_kernel<false><<<grid, kThreads, 0, st>>>(a, f);
    count_launch(ctx);
    JWC_CUDA_CHECK(cudaGetLastError());
    src = a.dst_lo;
    src_sig = a.dst_lo_sig;
  }
  return JWC_OK;
}

}  // namespace jwc
