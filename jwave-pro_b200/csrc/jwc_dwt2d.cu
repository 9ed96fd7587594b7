// jwc_dwt2d.cu -- column-direction analysis / synthesis steps for the 2-D FWT and WPT.
//
// Reference: transforms/BasicTransform.java:361-399 (2-D forward: every row through forward(row, lvlN), then every
// column through forward(col, lvlM)) and :436-474 (2-D reverse: columns first with lvlM, then rows with lvlN);
// transforms/ParallelTransform.java:70-91,222-271 computes the same thing with the rows spread over threads.  The
// reference gathers every column into a temporary array; here the row pass is the ordinary batched 1-D engine
// (batch*rows signals of length cols) and the column pass works IN PLACE on the row-major matrix: a thread owns one
// column of a 32-column strip, so every load and store of a warp is one contiguous 256-byte row segment, and walks a
// short run of output rows with the input rows it needs in registers (sliding window, 2 new rows per output pair).
// No transposes, no strided gathers; one launch per column level (the per-level traffic is h*cols reads + writes, a
// geometric series for the pyramid).
//
// Per-column arithmetic is Wavelet.java:236-260 (analysis) and :277-303 (synthesis) with `mod h` along the rows.
#include "jwc_internal.cuh"

namespace jwc {

namespace {

constexpr int kStrip = 32;   // columns per CTA (one warp = one 256-byte row segment)
constexpr int kGroups = 8;   // row groups per CTA
constexpr int kRun = 4;      // output rows (analysis) / output row pairs (synthesis) per thread

struct ColArgs {
  const double* src_lo;   // analysis: the block to analyse; synthesis: low-pass rows of each block
  const double* src_hi;   // synthesis only: high-pass rows of each block
  double* dst_lo;         // analysis: low-pass rows; synthesis: the synthesised block
  double* dst_hi;         // analysis only
  int64_t src_lo_mat, src_hi_mat, dst_lo_mat, dst_hi_mat;   // matrix strides, in doubles
  int64_t src_lo_blk, src_hi_blk, dst_lo_blk, dst_hi_blk;   // block (packet) strides, in doubles
  int64_t ld;       // doubles between consecutive rows (= cols; all buffers are dense row-major)
  int64_t cols;
  int64_t h;        // rows of one block at this level
  int64_t blocks;   // blocks per matrix (1 for the pyramid)
  int64_t tiles;    // row tiles per block
  int64_t strips;   // column strips
  int L;
};

struct Where {
  int64_t c, t, p, b;   // column, row tile, block, matrix
};
__device__ __forceinline__ Where locate(const ColArgs& a) {
  Where w;
  int64_t id = blockIdx.x;
  const int64_t s = id % a.strips;
  id /= a.strips;
  w.t = id % a.tiles;
  id /= a.tiles;
  w.p = id % a.blocks;
  w.b = id / a.blocks;
  w.c = s * kStrip + threadIdx.x;
  return w;
}

// ---- analysis, h >= 2, any L: lo[i] = sum_j x[(2i+j) mod h] s[j], hi likewise (j ascending, as the reference) ----
template <int L, bool EXACT>
__global__ void __launch_bounds__(kStrip* kGroups) col_ana_kernel(const __grid_constant__ ColArgs a,
                                                                  const __grid_constant__ FilterPair f) {
  const Where w = locate(a);
  const int64_t half = a.h >> 1;
  const int64_t i0 = (w.t * kGroups + threadIdx.y) * kRun;
  if (w.c >= a.cols || i0 >= half) return;
  const double* x = a.src_lo + w.b * a.src_lo_mat + w.p * a.src_lo_blk + w.c;
  constexpr int W = L + 2 * kRun - 2;
  double win[W];
  int64_t r = 2 * i0;   // < h
#pragma unroll
  for (int t = 0; t < W; t++) {
    win[t] = x[r * a.ld];
    if (++r == a.h) r = 0;   // single step, so blocks shorter than the filter simply go round several times
  }
  double* lo = a.dst_lo + w.b * a.dst_lo_mat + w.p * a.dst_lo_blk + w.c;
  double* hi = a.dst_hi + w.b * a.dst_hi_mat + w.p * a.dst_hi_blk + w.c;
#pragma unroll
  for (int q = 0; q < kRun; q++) {
    if (i0 + q < half) {
      double sl = 0.0, sh = 0.0;
#pragma unroll
      for (int j = 0; j < L; j++) {
        sl = mac<EXACT>(sl, win[2 * q + j], f.f0[j]);
        sh = mac<EXACT>(sh, win[2 * q + j], f.f1[j]);
      }
      lo[(i0 + q) * a.ld] = sl;
      hi[(i0 + q) * a.ld] = sh;
    }
  }
}

// ---- synthesis, fused multiply-add, needs h >= L (every coefficient row meets each output row at most once):
//      out[2q]   = sum_m lo[q-m] s[2m]   + hi[q-m] w[2m]
//      out[2q+1] = sum_m lo[q-m] s[2m+1] + hi[q-m] w[2m+1],   q-m taken mod h/2,  m < L/2
template <int L>
__global__ void __launch_bounds__(kStrip* kGroups) col_syn_kernel(const __grid_constant__ ColArgs a,
                                                                  const __grid_constant__ FilterPair f) {
  const Where w = locate(a);
  const int64_t half = a.h >> 1;
  const int64_t q0 = (w.t * kGroups + threadIdx.y) * kRun;
  if (w.c >= a.cols || q0 >= half) return;
  const double* lo = a.src_lo + w.b * a.src_lo_mat + w.p * a.src_lo_blk + w.c;
  const double* hi = a.src_hi + w.b * a.src_hi_mat + w.p * a.src_hi_blk + w.c;
  constexpr int M = L / 2;
  constexpr int W = kRun + M - 1;
  double wl[W], wh[W];
  int64_t i = q0 - (M - 1);
  if (i < 0) i += half;   // half >= M because h >= L
#pragma unroll
  for (int t = 0; t < W; t++) {
    wl[t] = lo[i * a.ld];
    wh[t] = hi[i * a.ld];
    if (++i == half) i = 0;
  }
  double* out = a.dst_lo + w.b * a.dst_lo_mat + w.p * a.dst_lo_blk + w.c;
#pragma unroll
  for (int u = 0; u < kRun; u++) {
    if (q0 + u < half) {
      double e = 0.0, o = 0.0;
#pragma unroll
      for (int m = M - 1; m >= 0; m--) {   // ascending coefficient row, like the reference's outer loop
        const double cl = wl[u - m + M - 1], ch = wh[u - m + M - 1];
        e = fma(ch, f.f1[2 * m], fma(cl, f.f0[2 * m], e));
        o = fma(ch, f.f1[2 * m + 1], fma(cl, f.f0[2 * m + 1], o));
      }
      out[(2 * (q0 + u)) * a.ld] = e;
      out[(2 * (q0 + u) + 1) * a.ld] = o;
    }
  }
}

// ---- shape-agnostic column steps (runtime L, blocks shorter than the filter, EXACT synthesis order) ----------------
template <bool EXACT>
__global__ void __launch_bounds__(kStrip* kGroups) col_ana_generic_kernel(const __grid_constant__ ColArgs a,
                                                                          const __grid_constant__ FilterPair f) {
  const Where w = locate(a);
  const int64_t half = a.h >> 1;
  if (w.c >= a.cols) return;
  const double* x = a.src_lo + w.b * a.src_lo_mat + w.p * a.src_lo_blk + w.c;
  double* lo = a.dst_lo + w.b * a.dst_lo_mat + w.p * a.dst_lo_blk + w.c;
  double* hi = a.dst_hi + w.b * a.dst_hi_mat + w.p * a.dst_hi_blk + w.c;
  const int64_t i0 = (w.t * kGroups + threadIdx.y) * kRun;
  for (int64_t i = i0; i < i0 + kRun && i < half; i++) {
    double sl = 0.0, sh = 0.0;
    int64_t r = 2 * i;
    for (int j = 0; j < a.L; j++) {
      const double xv = x[r * a.ld];
      sl = mac<EXACT>(sl, xv, f.f0[j]);
      sh = mac<EXACT>(sh, xv, f.f1[j]);
      if (++r == a.h) r = 0;
    }
    lo[i * a.ld] = sl;
    hi[i * a.ld] = sh;
  }
}

template <bool EXACT>
__device__ __forceinline__ double syn_term(double acc, double cl, double s, double ch, double w) {
  if (EXACT) return __dadd_rn(acc, __dadd_rn(__dmul_rn(cl, s), __dmul_rn(ch, w)));   // Wavelet.java:294-296
  return fma(ch, w, fma(cl, s, acc));
}

// gather form of the reference's scatter-add; the contributions to one output row are added in the reference's order
// (coefficient row i ascending, tap j ascending within a row)
template <bool EXACT>
__global__ void __launch_bounds__(kStrip* kGroups) col_syn_generic_kernel(const __grid_constant__ ColArgs a,
                                                                          const __grid_constant__ FilterPair f) {
  const Where w = locate(a);
  const int64_t h = a.h, half = h >> 1;
  if (w.c >= a.cols) return;
  const double* lo = a.src_lo + w.b * a.src_lo_mat + w.p * a.src_lo_blk + w.c;
  const double* hi = a.src_hi + w.b * a.src_hi_mat + w.p * a.src_hi_blk + w.c;
  double* out = a.dst_lo + w.b * a.dst_lo_mat + w.p * a.dst_lo_blk + w.c;
  const int L = a.L;
  const int64_t k0 = (w.t * kGroups + threadIdx.y) * (2 * kRun);
  for (int64_t k = k0; k < k0 + 2 * kRun && k < h; k++) {
    double acc = 0.0;
    if (h >= L) {
      int64_t i0 = (k - L + 2) >> 1;
      if (i0 < 0) i0 = 0;
      for (int64_t i = i0; 2 * i <= k; i++) {
        const int j = (int)(k - 2 * i);
        acc = syn_term<EXACT>(acc, lo[i * a.ld], f.f0[j], hi[i * a.ld], f.f1[j]);
      }
      for (int64_t i = (k + h - L + 2) >> 1; i < half; i++) {
        const int j = (int)(k + h - 2 * i);
        acc = syn_term<EXACT>(acc, lo[i * a.ld], f.f0[j], hi[i * a.ld], f.f1[j]);
      }
    } else {
      for (int64_t i = 0; i < half; i++) {
        int64_t j = (k - 2 * i) % h;
        if (j < 0) j += h;
        for (; j < L; j += h) acc = syn_term<EXACT>(acc, lo[i * a.ld], f.f0[j], hi[i * a.ld], f.f1[j]);
      }
    }
    out[k * a.ld] = acc;
  }
}

template <int L>
void launch_ana(const ColArgs& a, const FilterPair& f, dim3 grid, dim3 block, cudaStream_t st, bool exact) {
  if (exact) col_ana_kernel<L, true><<<grid, block, 0, st>>>(a, f);
  else       col_ana_kernel<L, false><<<grid, block, 0, st>>>(a, f);
}

#define JWC_EVEN_L(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) \
                      X(34) X(36) X(38) X(40)

int col_step(jwc_ctx* ctx, cudaStream_t st, ColArgs a, const FilterPair& f, int64_t batch, bool inverse, bool exact) {
  const int64_t half = a.h >> 1;
  a.strips = (a.cols + kStrip - 1) / kStrip;
  a.tiles = (half + kGroups * kRun - 1) / (kGroups * kRun);
  const int64_t ctas = a.strips * a.tiles * a.blocks * batch;
  if (ctas <= 0) return JWC_OK;
  if (ctas > 0x7fffffffLL) { set_error("2-D transform too large for one launch"); return JWC_ERR_UNSUPPORTED; }
  const dim3 grid((unsigned)ctas), block(kStrip, kGroups);
  bool done = false;
  if (!inverse) {
    switch (a.L) {
#define X(LL) case LL: launch_ana<LL>(a, f, grid, block, st, exact); done = true; break;
      JWC_EVEN_L(X)
#undef X
      default: break;
    }
    if (!done) {
      if (exact) col_ana_generic_kernel<true><<<grid, block, 0, st>>>(a, f);
      else       col_ana_generic_kernel<false><<<grid, block, 0, st>>>(a, f);
    }
  } else {
    if (!exact && a.h >= a.L) {
      switch (a.L) {
#define X(LL) case LL: col_syn_kernel<LL><<<grid, block, 0, st>>>(a, f); done = true; break;
        JWC_EVEN_L(X)
#undef X
        default: break;
      }
    }
    if (!done) {
      if (exact) col_syn_generic_kernel<true><<<grid, block, 0, st>>>(a, f);
      else       col_syn_generic_kernel<false><<<grid, block, 0, st>>>(a, f);
    }
  }
  count_launch(ctx);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

int count_steps(int64_t rows, int levels) {
  int steps = 0;   // the reference's loop: while h >= 2 && l < level
  for (int64_t h = rows; h >= 2 && steps < levels; h >>= 1) steps++;
  return steps;
}

}  // namespace

int dwt2d_column_steps(int64_t rows, int levels) { return count_steps(rows, levels); }

// Column pass of the 2-D forward transform: `levels` analysis steps along the rows of every column of the
// batch x rows x cols array d_src, result in d_out (same shape, must not overlap d_src).  levels >= 1 steps assumed.
int dwt2d_columns_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_src, double* d_out,
                          int64_t batch, int64_t rows, int64_t cols, int levels, const FilterPair& f, int L, bool tree,
                          bool exact) {
  (void)dev;
  const int steps = count_steps(rows, levels);
  const int64_t mat = rows * cols;
  Scratch ws(st);
  if (tree) {
    // every block of h rows -> [lo | hi] of itself; whole-array ping-pong, last step lands in d_out
    double* tmp = nullptr;
    if (steps >= 2) {
      tmp = ws.get((size_t)batch * mat);
      if (!tmp) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    }
    const double* src = d_src;
    int64_t h = rows;
    for (int l = 0; l < steps; l++, h >>= 1) {
      double* dst = (((steps - 1 - l) & 1) == 0) ? d_out : tmp;
      ColArgs a{};
      a.src_lo = src; a.src_lo_mat = mat; a.src_lo_blk = h * cols;
      a.dst_lo = dst; a.dst_lo_mat = mat; a.dst_lo_blk = h * cols;
      a.dst_hi = dst + (h >> 1) * cols; a.dst_hi_mat = mat; a.dst_hi_blk = h * cols;
      a.ld = cols; a.cols = cols; a.h = h; a.blocks = rows / h; a.L = L;
      const int rc = col_step(ctx, st, a, f, batch, false, exact);
      if (rc != JWC_OK) return rc;
      src = dst;
    }
    return JWC_OK;
  }
  // pyramid: detail rows go straight to their final place in d_out, approximations ping-pong through scratch
  double* abuf[2] = {nullptr, nullptr};
  if (steps >= 2) {
    abuf[0] = ws.get((size_t)batch * (mat >> 1));
    if (!abuf[0]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  if (steps >= 3) {
    abuf[1] = ws.get((size_t)batch * (mat >> 2));
    if (!abuf[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* src = d_src;
  int64_t src_mat = mat;
  int64_t h = rows;
  for (int l = 0; l < steps; l++, h >>= 1) {
    ColArgs a{};
    a.src_lo = src; a.src_lo_mat = src_mat;
    if (l == steps - 1) { a.dst_lo = d_out; a.dst_lo_mat = mat; }
    else { a.dst_lo = abuf[l & 1]; a.dst_lo_mat = (h >> 1) * cols; }
    a.dst_hi = d_out + (h >> 1) * cols; a.dst_hi_mat = mat;
    a.ld = cols; a.cols = cols; a.h = h; a.blocks = 1; a.L = L;
    const int rc = col_step(ctx, st, a, f, batch, false, exact);
    if (rc != JWC_OK) return rc;
    src = a.dst_lo;
    src_mat = a.dst_lo_mat;
  }
  // rows the column pass never touches do not exist: step 0 always covers all `rows` rows
  return JWC_OK;
}

// Column pass of the 2-D reverse transform (BasicTransform.java:444-456 runs it BEFORE the rows).
int dwt2d_columns_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_dst,
                          int64_t batch, int64_t rows, int64_t cols, int levels, const FilterPair& f, int L, bool tree,
                          bool exact) {
  const int steps = count_steps(rows, levels);
  const int64_t mat = rows * cols;
  Scratch ws(st);
  if (tree) {
    double* tmp = nullptr;
    if (steps >= 2) {
      tmp = ws.get((size_t)batch * mat);
      if (!tmp) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    }
    const double* src = d_in;
    int64_t h = rows >> (steps - 1);
    for (int l = 0; l < steps; l++, h <<= 1) {
      double* dst = (((steps - 1 - l) & 1) == 0) ? d_dst : tmp;
      ColArgs a{};
      a.src_lo = src; a.src_lo_mat = mat; a.src_lo_blk = h * cols;
      a.src_hi = src + (h >> 1) * cols; a.src_hi_mat = mat; a.src_hi_blk = h * cols;
      a.dst_lo = dst; a.dst_lo_mat = mat; a.dst_lo_blk = h * cols;
      a.ld = cols; a.cols = cols; a.h = h; a.blocks = rows / h; a.L = L;
      const int rc = col_step(ctx, st, a, f, batch, true, exact);
      if (rc != JWC_OK) return rc;
      src = dst;
    }
    return JWC_OK;
  }
  // pyramid: A_l from scratch (or the input for the first step), D_l from the input; the rows above the deepest
  // block that the steps never rebuild are the detail rows consumed later, so every row of d_dst is written by the
  // last step (h = rows).
  double* abuf[2] = {nullptr, nullptr};
  if (steps >= 2) {
    abuf[0] = ws.get((size_t)batch * (mat >> 1));
    if (!abuf[0]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  if (steps >= 3) {
    abuf[1] = ws.get((size_t)batch * (mat >> 2));
    if (!abuf[1]) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  }
  const double* alo = d_in;
  int64_t alo_mat = mat;
  int64_t h = rows >> (steps - 1);
  for (int l = 0; l < steps; l++, h <<= 1) {
    ColArgs a{};
    a.src_lo = alo; a.src_lo_mat = alo_mat;
    a.src_hi = d_in + (h >> 1) * cols; a.src_hi_mat = mat;
    if (l == steps - 1) { a.dst_lo = d_dst; a.dst_lo_mat = mat; }
    else { a.dst_lo = abuf[(steps - l) & 1]; a.dst_lo_mat = h * cols; }
    a.ld = cols; a.cols = cols; a.h = h; a.blocks = 1; a.L = L;
    const int rc = col_step(ctx, st, a, f, batch, true, exact);
    if (rc != JWC_OK) return rc;
    alo = a.dst_lo;
    alo_mat = a.dst_lo_mat;
  }
  (void)dev;
  return JWC_OK;
}

}  // namespace jwc
