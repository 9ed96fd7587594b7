// jwc_api.cu -- the extern "C" boundary of libjwavecuda.so (include/jwavecuda.h): context, memory helpers,
// argument validation, dispatch (fused tile kernels when the shape allows, generic kernels otherwise -- both CUDA,
// there is no CPU path), and the host-buffer pipeline that shards a batch over the context's devices.
#include <cstdarg>
#include <cstring>
#include <thread>

#include "jwc_internal.cuh"

namespace jwc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int ordinal) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != ordinal) ok = (cudaSetDevice(ordinal) == cudaSuccess);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

enum class Op { ModwtFwd, ModwtInv, FwtFwd, FwtInv, WptFwd, WptInv };

// 2-D calls reuse the 1-D plumbing: `n` is the row length (cols), `levels` the level along a row (lvlN); rows > 0
// switches the unit of work from one signal to one rows x n matrix with lvl_m levels down the columns.
struct Dim2 {
  int64_t rows = 0;
  int lvl_m = 0;
  // 3-D (BasicTransform.java:487-640): a unit is a space [depth][rows][n]; every [rows][n] matrix goes through the 2-D
  // transform, then every line along the first axis through the 1-D transform with lvl_depth levels
  int64_t depth = 0;
  int lvl_depth = 0;
  bool aed = false;   // Ancient-Egyptian decomposition of arbitrary-length signals (levels ignored: full depth per block)
  int64_t hop = 0;    // > 0: forward MODWT of overlapping windows of ONE series; window b starts at b * hop
};

}  // namespace
cudaError_t pool_alloc(const DeviceSlot& dev, void** p, size_t bytes, cudaStream_t st) {
  return dev.pool ? cudaMallocFromPoolAsync(p, bytes ? bytes : 1, dev.pool, st) : cudaMallocAsync(p, bytes ? bytes : 1, st);
}
namespace {

bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }
int ilog2(int64_t n) {
  int p = 0;
  while (((int64_t)1 << (p + 1)) <= n) p++;
  return p;
}

int validate(Op op, const void* in, const void* out, int64_t batch, int64_t n, int levels, const double* f0,
             const double* f1, int L, const Dim2& d2 = Dim2()) {
  if (d2.depth != 0) {
    JWC_REQUIRE(d2.rows != 0, "a space needs its matrix dimensions");
    JWC_REQUIRE(is_pow2(d2.depth), "given space depth is not 2^p (got %lld)", (long long)d2.depth);
    JWC_REQUIRE(d2.lvl_depth >= 0 && d2.lvl_depth <= ilog2(d2.depth),
                "given level %d is out of range for given array of length %lld", d2.lvl_depth, (long long)d2.depth);
    JWC_REQUIRE(d2.depth < ((int64_t)1 << 40) / (n > 0 ? n : 1) / (d2.rows > 0 ? d2.rows : 1), "space %lld x %lld x %lld too large",
                (long long)d2.depth, (long long)d2.rows, (long long)n);
  }
  if (d2.rows != 0) {
    JWC_REQUIRE(op != Op::ModwtFwd && op != Op::ModwtInv, "no 2-D MODWT");
    JWC_REQUIRE(is_pow2(d2.rows), "given matrix height is not 2^p (got %lld)", (long long)d2.rows);
    JWC_REQUIRE(d2.lvl_m >= 0 && d2.lvl_m <= ilog2(d2.rows),
                "given level %d is out of range for given array of length %lld", d2.lvl_m, (long long)d2.rows);
    JWC_REQUIRE(d2.rows < ((int64_t)1 << 40) / (n > 0 ? n : 1), "matrix %lld x %lld too large", (long long)d2.rows,
                (long long)n);
  }
  JWC_REQUIRE(in != nullptr && out != nullptr, "input/output pointer is NULL");
  JWC_REQUIRE(f0 != nullptr && f1 != nullptr, "filter pointer is NULL");
  JWC_REQUIRE(batch >= 0, "batch must be >= 0 (got %lld)", (long long)batch);
  JWC_REQUIRE(n >= 1, "signal length must be >= 1 (got %lld)", (long long)n);
  JWC_REQUIRE(n < ((int64_t)1 << 40), "signal length %lld too large", (long long)n);
  JWC_REQUIRE(L >= 1 && L <= JWC_MAX_TAPS, "filter length %d outside 1..%d", L, JWC_MAX_TAPS);
  if (d2.hop != 0) {
    JWC_REQUIRE(op == Op::ModwtFwd, "sliding windows exist for the forward MODWT only");
    JWC_REQUIRE(d2.hop >= 1, "window hop must be >= 1 (got %lld)", (long long)d2.hop);
  }
  if (d2.aed) {
    JWC_REQUIRE(op != Op::ModwtFwd && op != Op::ModwtInv, "no Ancient-Egyptian MODWT");
    JWC_REQUIRE(n < ((int64_t)1 << 31), "signal length %lld too large", (long long)n);
  } else if (op == Op::ModwtFwd || op == Op::ModwtInv) {
    JWC_REQUIRE(levels >= 1 && levels <= 40, "MODWT level %d out of range", levels);
  } else {
    JWC_REQUIRE(is_pow2(n), "given array length is not 2^p (got %lld)", (long long)n);
    JWC_REQUIRE(levels >= 0 && levels <= ilog2(n), "given level %d is out of range for given array of length %lld",
                levels, (long long)n);
  }
  return JWC_OK;
}

void load_filters(FilterPair& fp, const double* f0, const double* f1, int L) {
  memset(&fp, 0, sizeof(fp));
  memcpy(fp.f0, f0, sizeof(double) * (size_t)L);
  memcpy(fp.f1, f1, sizeof(double) * (size_t)L);
}

// One transform on device-resident buffers of slot `dev`, enqueued on `st`.
int run_device(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, Op op, const double* d_in, double* d_out,
               int64_t batch, int64_t n, int levels, const FilterPair& fp, int L, unsigned flags,
               const Dim2& d2 = Dim2()) {
  if (batch == 0) return JWC_OK;
  const bool exact = (flags & JWC_FLAG_EXACT) != 0;
  if (d2.aed) {
    // transforms/AncientEgyptianDecomposition.java:97-181 + tools/MathToolKit.java:57-84: n = sum of descending powers
    // of two; every block [off, off + 2^p) of every signal goes through the wrapped transform at full depth (p levels)
    // on its own.  Here each block is ONE batched launch sequence over all signals (row stride n), no copies; a block
    // of length 1 is the identity.
    const bool tree = (op == Op::WptFwd || op == Op::WptInv);
    const bool fwd = (op == Op::FwtFwd || op == Op::WptFwd);
    const bool generic = exact || (flags & JWC_FLAG_FORCE_GENERIC) != 0 || ctx->tune.force_generic != 0;
    int64_t off = 0;
    for (int p = 30; p >= 0; p--) {
      const int64_t len = (int64_t)1 << p;
      if (!(n & len)) continue;
      const double* src = d_in + off;
      double* dst = d_out + off;
      int rc = JWC_ERR_UNSUPPORTED;
      if (!generic && p > 0)
        rc = fwd ? fast_dwt_forward(ctx, dev, st, src, dst, batch, len, p, fp, L, tree, n)
                 : fast_dwt_inverse(ctx, dev, st, src, dst, batch, len, p, fp, L, tree, n);
      if (rc == JWC_ERR_UNSUPPORTED)
        rc = fwd ? generic_dwt_forward(ctx, dev, st, src, dst, batch, len, p, fp, L, tree, exact, n)
                 : generic_dwt_inverse(ctx, dev, st, src, dst, batch, len, p, fp, L, tree, exact, n);
      if (rc != JWC_OK) return rc;
      off += len;
    }
    return JWC_OK;
  }
  if (d2.depth != 0) {
    // BasicTransform.java:509-565 forward(double[][][], lvlP, lvlQ, lvlR): the 2-D transform of every [rows][n] matrix,
    // then every line along the first axis with lvlR levels; :602-640 reverse keeps that order (2-D reverse first).
    // The first-axis pass is the 2-D column pass on the view [batch][depth][rows * n] -- in place on the row-major
    // space, a warp's access is one contiguous row segment.
    const bool tree = (op == Op::WptFwd || op == Op::WptInv);
    const bool fwd = (op == Op::FwtFwd || op == Op::WptFwd);
    Dim2 mat = d2;
    mat.depth = 0;
    mat.lvl_depth = 0;
    if (dwt2d_column_steps(d2.depth, d2.lvl_depth) == 0)
      return run_device(ctx, dev, st, op, d_in, d_out, batch * d2.depth, n, levels, fp, L, flags, mat);
    Scratch ws(ctx, dev, st);
    double* mid = ws.get((size_t)(batch * d2.depth * d2.rows * n));
    if (!mid) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    const int rc = run_device(ctx, dev, st, op, d_in, mid, batch * d2.depth, n, levels, fp, L, flags, mat);
    if (rc != JWC_OK) return rc;
    return fwd ? dwt2d_columns_forward(ctx, dev, st, mid, d_out, batch, d2.depth, d2.rows * n, d2.lvl_depth, fp, L, tree, exact)
               : dwt2d_columns_inverse(ctx, dev, st, mid, d_out, batch, d2.depth, d2.rows * n, d2.lvl_depth, fp, L, tree, exact);
  }
  if (d2.rows != 0) {
    // BasicTransform.java:361-399 forward: rows (lvlN) then columns (lvlM); :436-474 reverse: columns, then rows.
    const bool tree = (op == Op::WptFwd || op == Op::WptInv);
    const bool fwd = (op == Op::FwtFwd || op == Op::WptFwd);
    const int64_t rows = d2.rows, cols = n;
    const int col_steps = dwt2d_column_steps(rows, d2.lvl_m);
    const int row_steps = dwt2d_column_steps(cols, levels);
    if (col_steps == 0) return run_device(ctx, dev, st, op, d_in, d_out, batch * rows, cols, levels, fp, L, flags);
    if (row_steps == 0)
      return fwd ? dwt2d_columns_forward(ctx, dev, st, d_in, d_out, batch, rows, cols, d2.lvl_m, fp, L, tree, exact)
                 : dwt2d_columns_inverse(ctx, dev, st, d_in, d_out, batch, rows, cols, d2.lvl_m, fp, L, tree, exact);
    Scratch ws(ctx, dev, st);
    double* mid = ws.get((size_t)(batch * rows * cols));
    if (!mid) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
    if (fwd) {
      const int rc = run_device(ctx, dev, st, op, d_in, mid, batch * rows, cols, levels, fp, L, flags);
      if (rc != JWC_OK) return rc;
      return dwt2d_columns_forward(ctx, dev, st, mid, d_out, batch, rows, cols, d2.lvl_m, fp, L, tree, exact);
    }
    const int rc = dwt2d_columns_inverse(ctx, dev, st, d_in, mid, batch, rows, cols, d2.lvl_m, fp, L, tree, exact);
    if (rc != JWC_OK) return rc;
    return run_device(ctx, dev, st, op, mid, d_out, batch * rows, cols, levels, fp, L, flags);
  }
  const bool generic = exact || (flags & JWC_FLAG_FORCE_GENERIC) != 0 || ctx->tune.force_generic != 0;
  int rc = JWC_ERR_UNSUPPORTED;
  if (!generic) {
    switch (op) {
      case Op::ModwtFwd:
        rc = small_modwt_forward(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, d2.hop);
        if (rc == JWC_ERR_UNSUPPORTED) rc = fast_modwt_forward(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, d2.hop);
        break;
      case Op::ModwtInv:
        rc = small_modwt_inverse(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L);
        if (rc == JWC_ERR_UNSUPPORTED) rc = fast_modwt_inverse(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L);
        break;
      case Op::FwtFwd: rc = fast_dwt_forward(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, false); break;
      case Op::FwtInv: rc = fast_dwt_inverse(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, false); break;
      case Op::WptFwd: rc = fast_dwt_forward(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, true); break;
      case Op::WptInv: rc = fast_dwt_inverse(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, true); break;
    }
    if (rc != JWC_ERR_UNSUPPORTED) return rc;
  }
  switch (op) {
    case Op::ModwtFwd:
      return generic_modwt_forward_from(ctx, dev, st, d_in, d2.hop > 0 ? d2.hop : n, 1, d_out, batch, n, levels, fp, L, exact);
    case Op::ModwtInv: return generic_modwt_inverse(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, exact);
    case Op::FwtFwd: return generic_dwt_forward(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, false, exact);
    case Op::FwtInv: return generic_dwt_inverse(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, false, exact);
    case Op::WptFwd: return generic_dwt_forward(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, true, exact);
    case Op::WptInv: return generic_dwt_inverse(ctx, dev, st, d_in, d_out, batch, n, levels, fp, L, true, exact);
  }
  return JWC_ERR_INVALID;
}

void io_sizes(Op op, int64_t n, int levels, int64_t* in_per, int64_t* out_per, const Dim2& d2 = Dim2()) {
  *in_per = n;
  *out_per = n;
  if (d2.rows != 0) *in_per = *out_per = n * d2.rows * (d2.depth != 0 ? d2.depth : 1);
  // sliding windows: *in_per stays n (one window); consecutive windows start d2.hop apart (see in_step below)
  if (op == Op::ModwtFwd) *out_per = (int64_t)(levels + 1) * n;
  if (op == Op::ModwtInv) *in_per = (int64_t)(levels + 1) * n;
}

// Host-buffer pipeline on ONE slot: chunks of signals, double-buffered device staging, H2D / kernels / D2H on three
// streams chained by events.  Called on its own host thread per slot when the context spans several devices.
// Borrow a stream triple of `slot` for one host-buffer call (device must be current); give it back when the call is
// drained.  Returns false when stream creation fails.
void lane_destroy(Lane& l) {
  for (cudaStream_t* st : {&l.compute, &l.copy_in, &l.copy_out})
    if (*st) { cudaStreamDestroy(*st); *st = nullptr; }
}

bool lane_acquire(jwc_ctx* ctx, int slot, Lane* lane) {
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto& idle = ctx->slots[slot].idle_lanes;
    if (!idle.empty()) {
      *lane = idle.back();
      idle.pop_back();
      return true;
    }
  }
  Lane l;
  if (cudaStreamCreateWithFlags(&l.compute, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&l.copy_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&l.copy_out, cudaStreamNonBlocking) != cudaSuccess) {
    (void)cudaGetLastError();
    lane_destroy(l);
    return false;
  }
  *lane = l;
  return true;
}
void lane_release(jwc_ctx* ctx, int slot, const Lane& lane) {
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->slots[slot].idle_lanes.push_back(lane);
}

int run_host_slot(jwc_ctx* ctx, int slot, Op op, const double* in, double* out, int64_t batch, int64_t n, int levels,
                  const FilterPair& fp, int L, unsigned flags, const Dim2& d2 = Dim2()) {
  if (batch == 0) return JWC_OK;
  const DeviceSlot& dev = ctx->slots[slot];
  DeviceGuard guard(dev.ordinal);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", dev.ordinal); return JWC_ERR_CUDA; }
  Lane lane;
  if (!lane_acquire(ctx, slot, &lane)) { set_error("cannot create streams on device %d", dev.ordinal); return JWC_ERR_CUDA; }
  int64_t in_per, out_per;
  io_sizes(op, n, levels, &in_per, &out_per, d2);
  // sliding windows: unit b of the input starts at b * hop and is in_per long, so a chunk of nb units spans
  // (nb - 1) * hop + in_per input samples; everywhere else units are dense (in_step == in_per)
  const int64_t in_step = d2.hop > 0 ? d2.hop : in_per;
  auto in_span = [&](int64_t nb) { return (nb - 1) * in_step + in_per; };
  const int64_t per_sig_bytes = (in_step + out_per) * (int64_t)sizeof(double);
  int64_t chunk_mb = ctx->tune.h2d_chunk_mb > 0 ? ctx->tune.h2d_chunk_mb : 128;   // in + out bytes per chunk
  int64_t chunk = (chunk_mb << 20) / per_sig_bytes;
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  constexpr int kMaxBuf = 4;
  const int64_t nchunks = (batch + chunk - 1) / chunk;
  int nbuf = ctx->tune.h2d_buffers > 0 ? ctx->tune.h2d_buffers : 3;   // staging depth of the copy/compute pipeline
  if (nbuf > kMaxBuf) nbuf = kMaxBuf;
  if (nbuf > nchunks) nbuf = (int)nchunks;

  cudaStream_t sc = lane.compute, si = lane.copy_in, so = lane.copy_out;
  // Copy pacing.  A copy engine drains one stream's queued copies before it looks at another stream (measured: with a
  // forward and an inverse call in flight, the forward's 33 MB H2D sat behind ALL of the inverse's 235 MB H2D chunks,
  // serialising the two calls; piece-wise copies and stream priorities did not change that).  So while other host
  // calls are in flight, a call keeps at most ONE copy queued in the direction it is heavy in: the next one is issued
  // only when the previous one has finished, which lets the other call's small copy slip in between.
  const bool pace_in = in_per >= out_per, pace_out = out_per > in_per;
  struct InFlight {
    std::atomic<int>& n;
    explicit InFlight(std::atomic<int>& a) : n(a) { n.fetch_add(1); }
    ~InFlight() { n.fetch_sub(1); }
  } in_flight(ctx->host_calls);
  double* d_in[kMaxBuf] = {};
  double* d_out[kMaxBuf] = {};
  cudaEvent_t ev_in[kMaxBuf] = {}, ev_k[kMaxBuf] = {}, ev_out[kMaxBuf] = {}, ev_alloc = nullptr;
  int rc = JWC_OK;
  auto fail = [&](cudaError_t e, const char* what) {
    set_error("%s failed: %s", what, cudaGetErrorString(e));
    rc = JWC_ERR_CUDA;
  };
  cudaError_t e;
  Scratch staging(ctx, dev, sc);   // kept in the lane's arena between calls; released when this call has drained
  for (int i = 0; i < nbuf && rc == JWC_OK; i++) {
    d_in[i] = staging.get((size_t)in_span(chunk));
    d_out[i] = staging.get((size_t)(chunk * out_per));
    if (!d_in[i] || !d_out[i]) {
      set_error("device staging allocation of %lld MiB failed", (long long)((chunk * (in_per + out_per) * 8) >> 20));
      rc = JWC_ERR_NOMEM;
    }
    if (rc == JWC_OK && ((e = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming)) != cudaSuccess ||
                         (e = cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming)) != cudaSuccess ||
                         (e = cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming)) != cudaSuccess))
      fail(e, "cudaEventCreate");
  }
  if (rc == JWC_OK && (e = cudaEventCreateWithFlags(&ev_alloc, cudaEventDisableTiming)) != cudaSuccess)
    fail(e, "cudaEventCreate");
  if (rc == JWC_OK) {
    // staging was allocated in stream order on sc: the copy streams must not touch it earlier
    if ((e = cudaEventRecord(ev_alloc, sc)) != cudaSuccess || (e = cudaStreamWaitEvent(si, ev_alloc, 0)) != cudaSuccess ||
        (e = cudaStreamWaitEvent(so, ev_alloc, 0)) != cudaSuccess)
      fail(e, "stream ordering");
  }
  int64_t c = 0;
  for (int64_t b0 = 0; b0 < batch && rc == JWC_OK; b0 += chunk, c++) {
    const int64_t nb = (batch - b0 < chunk) ? batch - b0 : chunk;
    const int k = (int)(c % nbuf);
    if (c >= nbuf) {
      // d_in[k] is free once the kernels of chunk c-nbuf ran; d_out[k] once its D2H finished
      if ((e = cudaStreamWaitEvent(si, ev_k[k], 0)) != cudaSuccess) { fail(e, "cudaStreamWaitEvent"); break; }
      if ((e = cudaStreamWaitEvent(sc, ev_out[k], 0)) != cudaSuccess) { fail(e, "cudaStreamWaitEvent"); break; }
    }
    if (c > 0 && pace_in && ctx->host_calls.load() > 1) cudaEventSynchronize(ev_in[(c - 1) % nbuf]);
    if ((e = cudaMemcpyAsync(d_in[k], in + b0 * in_step, (size_t)in_span(nb) * sizeof(double), cudaMemcpyHostToDevice,
                             si)) != cudaSuccess) { fail(e, "cudaMemcpyAsync(H2D)"); break; }
    if ((e = cudaEventRecord(ev_in[k], si)) != cudaSuccess) { fail(e, "cudaEventRecord"); break; }
    if ((e = cudaStreamWaitEvent(sc, ev_in[k], 0)) != cudaSuccess) { fail(e, "cudaStreamWaitEvent"); break; }
    rc = run_device(ctx, dev, sc, op, d_in[k], d_out[k], nb, n, levels, fp, L, flags, d2);
    if (rc != JWC_OK) break;
    if ((e = cudaEventRecord(ev_k[k], sc)) != cudaSuccess) { fail(e, "cudaEventRecord"); break; }
    if ((e = cudaStreamWaitEvent(so, ev_k[k], 0)) != cudaSuccess) { fail(e, "cudaStreamWaitEvent"); break; }
    if (c > 0 && pace_out && ctx->host_calls.load() > 1) cudaEventSynchronize(ev_out[(c - 1) % nbuf]);
    if ((e = cudaMemcpyAsync(out + b0 * out_per, d_out[k], (size_t)(nb * out_per) * sizeof(double),
                             cudaMemcpyDeviceToHost, so)) != cudaSuccess) { fail(e, "cudaMemcpyAsync(D2H)"); break; }
    if ((e = cudaEventRecord(ev_out[k], so)) != cudaSuccess) { fail(e, "cudaEventRecord"); break; }
  }
  // drain, then release staging in stream order on sc
  cudaError_t e1 = cudaStreamSynchronize(si), e2 = cudaStreamSynchronize(sc), e3 = cudaStreamSynchronize(so);
  if (rc == JWC_OK) {
    if (e1 != cudaSuccess) fail(e1, "cudaStreamSynchronize(copy_in)");
    else if (e2 != cudaSuccess) fail(e2, "cudaStreamSynchronize(compute)");
    else if (e3 != cudaSuccess) fail(e3, "cudaStreamSynchronize(copy_out)");
  }
  for (int i = 0; i < kMaxBuf; i++) {
    if (ev_in[i]) cudaEventDestroy(ev_in[i]);
    if (ev_k[i]) cudaEventDestroy(ev_k[i]);
    if (ev_out[i]) cudaEventDestroy(ev_out[i]);
  }
  if (ev_alloc) cudaEventDestroy(ev_alloc);
  staging.release();               // before the lane goes back: its next user finds the staging blocks free
  lane_release(ctx, slot, lane);
  return rc;
}

int run_host(jwc_ctx* ctx, Op op, const double* in, double* out, int64_t batch, int64_t n, int levels,
             const double* f0, const double* f1, int L, unsigned flags, const Dim2& d2 = Dim2()) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  int rc = validate(op, in, out, batch, n, levels, f0, f1, L, d2);
  if (rc != JWC_OK) return rc;
  FilterPair fp;
  load_filters(fp, f0, f1, L);
  int64_t in_per, out_per;
  io_sizes(op, n, levels, &in_per, &out_per, d2);
  const int nd = (int)ctx->slots.size();
  if (nd == 1 || batch < 2) return run_host_slot(ctx, 0, op, in, out, batch, n, levels, fp, L, flags, d2);
  // shard by signal: contiguous blocks, no data-path collective
  std::vector<std::thread> th;
  std::vector<int> rcs(nd, JWC_OK);
  std::vector<std::string> errs(nd);
  for (int s = 0; s < nd; s++) {
    const int64_t b0 = batch * s / nd, b1 = batch * (s + 1) / nd;
    th.emplace_back([=, &rcs, &errs, &fp]() {
      rcs[s] = run_host_slot(ctx, s, op, in + b0 * (d2.hop > 0 ? d2.hop : in_per), out + b0 * out_per, b1 - b0, n, levels,
                             fp, L, flags, d2);
      if (rcs[s] != JWC_OK) errs[s] = g_err;
    });
  }
  for (auto& t : th) t.join();
  for (int s = 0; s < nd; s++)
    if (rcs[s] != JWC_OK) {
      set_error("device slot %d: %s", s, errs[s].c_str());
      return rcs[s];
    }
  return JWC_OK;
}

int run_dev(jwc_ctx* ctx, int slot, void* stream, Op op, const double* d_in, double* d_out, int64_t batch, int64_t n,
            int levels, const double* f0, const double* f1, int L, unsigned flags, const Dim2& d2 = Dim2()) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), "device slot %d out of range", slot);
  int rc = validate(op, d_in, d_out, batch, n, levels, f0, f1, L, d2);
  if (rc != JWC_OK) return rc;
  FilterPair fp;
  load_filters(fp, f0, f1, L);
  const DeviceSlot& dev = ctx->slots[slot];
  DeviceGuard guard(dev.ordinal);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", dev.ordinal); return JWC_ERR_CUDA; }
  cudaStream_t st = stream ? (cudaStream_t)stream : dev.stream;
  return run_device(ctx, dev, st, op, d_in, d_out, batch, n, levels, fp, L, flags, d2);
}


// ---------------------------------------------------------------------------------------------------------------------
// One long series split over the context's devices (second row).
// Device slot p owns the contiguous chunk [n*p/P, n*(p+1)/P).  MODWT output t depends on inputs t-H .. t (forward,
// H = (L-1)(2^J - 1)) resp. t .. t+H (inverse), so one halo exchange per transform is enough: every device extends
// its chunk with the neighbour's H boundary samples (cudaMemcpyPeerAsync over NVLink, ring order, circular), runs the
// ordinary kernels on the extended chunk (their circular wrap only contaminates the H positions that are thrown away)
// and keeps the chunk part.  No collective.
// ---------------------------------------------------------------------------------------------------------------------
int modwt_split(jwc_ctx* ctx, bool inverse, const double* const* d_in, double* const* d_out, int64_t n, int levels,
                const double* g, const double* h, int L, unsigned flags) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(d_in != nullptr && d_out != nullptr && g != nullptr && h != nullptr, "NULL pointer");
  JWC_REQUIRE(levels >= 1 && levels <= 30, "MODWT level %d out of range", levels);
  JWC_REQUIRE(L >= 1 && L <= JWC_MAX_TAPS, "filter length %d outside 1..%d", L, JWC_MAX_TAPS);
  const int P = (int)ctx->slots.size();
  const int64_t H0 = (int64_t)(L - 1) * (((int64_t)1 << levels) - 1);
  const int64_t H = H0 + (H0 & 1);   // even, so the extended chunk keeps the 16-byte alignment of the bulk copies
  for (int p = 0; p < P; p++) {
    const int64_t len = n * (p + 1) / P - n * p / P;
    JWC_REQUIRE(d_in[p] != nullptr && d_out[p] != nullptr, "chunk pointer %d is NULL", p);
    JWC_REQUIRE(len >= H && len >= 1, "series of %lld samples is too short to split over %d devices (halo %lld)",
                (long long)n, P, (long long)H);
  }
  FilterPair fp;
  load_filters(fp, g, h, L);
  const int rows = levels + 1;
  int rc = JWC_OK;
  std::vector<double*> ext_in(P, nullptr), ext_out(P, nullptr);
  // phase 1: build the extended inputs (own chunk + neighbour halo), all devices
  for (int p = 0; p < P && rc == JWC_OK; p++) {
    const DeviceSlot& dev = ctx->slots[p];
    DeviceGuard guard(dev.ordinal);
    const int64_t len = n * (p + 1) / P - n * p / P, next = len + H;
    const int in_rows = inverse ? rows : 1, out_rows = inverse ? 1 : rows;
    cudaError_t e;
    if ((e = pool_alloc(dev, (void**)&ext_in[p], (size_t)(in_rows * next) * sizeof(double), dev.stream)) != cudaSuccess ||
        (e = pool_alloc(dev, (void**)&ext_out[p], (size_t)(out_rows * next) * sizeof(double), dev.stream)) != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("split staging allocation failed on slot %d: %s", p, cudaGetErrorString(e));
      rc = JWC_ERR_NOMEM;
      break;
    }
    if (!inverse) {
      // [halo from the left neighbour's tail | own chunk]
      const int q = (p + P - 1) % P;
      const int64_t qlen = n * (q + 1) / P - n * q / P;
      e = cudaMemcpyPeerAsync(ext_in[p], dev.ordinal, d_in[q] + (qlen - H), ctx->slots[q].ordinal, (size_t)H * sizeof(double),
                              dev.stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(ext_in[p] + H, d_in[p], (size_t)len * sizeof(double), cudaMemcpyDeviceToDevice, dev.stream);
    } else {
      // every coefficient row: [own chunk | halo from the right neighbour's head]
      const int q = (p + 1) % P;
      const int64_t qlen = n * (q + 1) / P - n * q / P;
      // one plain copy per row (at most 31): cudaMemcpy2D rejects pitches beyond cudaDeviceProp::memPitch (2 GiB),
      // which is exactly the long-series case this entry point exists for
      e = cudaSuccess;
      for (int r = 0; r < rows && e == cudaSuccess; r++)
        e = cudaMemcpyAsync(ext_in[p] + (int64_t)r * next, d_in[p] + (int64_t)r * len, (size_t)len * sizeof(double),
                            cudaMemcpyDeviceToDevice, dev.stream);
      for (int r = 0; r < rows && e == cudaSuccess; r++)
        e = cudaMemcpyPeerAsync(ext_in[p] + (int64_t)r * next + len, dev.ordinal, d_in[q] + (int64_t)r * qlen,
                                ctx->slots[q].ordinal, (size_t)H * sizeof(double), dev.stream);
    }
    if (e != cudaSuccess) { set_error("halo exchange failed on slot %d: %s", p, cudaGetErrorString(e)); rc = JWC_ERR_CUDA; }
  }
  // phase 2: the ordinary transform on the extended chunk, then keep the chunk part
  for (int p = 0; p < P && rc == JWC_OK; p++) {
    const DeviceSlot& dev = ctx->slots[p];
    DeviceGuard guard(dev.ordinal);
    const int64_t len = n * (p + 1) / P - n * p / P, next = len + H;
    rc = run_device(ctx, dev, dev.stream, inverse ? Op::ModwtInv : Op::ModwtFwd, ext_in[p], ext_out[p], 1, next, levels, fp,
                    L, flags);
    if (rc != JWC_OK) break;
    cudaError_t e;
    if (!inverse) {   // rows of the extended result, positions H .. H+len (row by row: no 2-D pitch limit)
      e = cudaSuccess;
      for (int r = 0; r < rows && e == cudaSuccess; r++)
        e = cudaMemcpyAsync(d_out[p] + (int64_t)r * len, ext_out[p] + (int64_t)r * next + H, (size_t)len * sizeof(double),
                            cudaMemcpyDeviceToDevice, dev.stream);
    } else            // positions 0 .. len
      e = cudaMemcpyAsync(d_out[p], ext_out[p], (size_t)len * sizeof(double), cudaMemcpyDeviceToDevice, dev.stream);
    if (e != cudaSuccess) { set_error("split result copy failed on slot %d: %s", p, cudaGetErrorString(e)); rc = JWC_ERR_CUDA; }
  }
  for (int p = 0; p < P; p++) {
    const DeviceSlot& dev = ctx->slots[p];
    DeviceGuard guard(dev.ordinal);
    if (ext_in[p]) cudaFreeAsync(ext_in[p], dev.stream);
    if (ext_out[p]) cudaFreeAsync(ext_out[p], dev.stream);
    cudaError_t e = cudaStreamSynchronize(dev.stream);
    if (e != cudaSuccess && rc == JWC_OK) { set_error("split: stream sync failed on slot %d: %s", p, cudaGetErrorString(e)); rc = JWC_ERR_CUDA; }
  }
  return rc;
}

}  // namespace
}  // namespace jwc

using namespace jwc;

extern "C" {

JWC_API const char* jwc_last_error(void) { return g_err; }
JWC_API const char* jwc_version(void) { return "jwavecuda 0.1 (sm_100a)"; }

JWC_API jwc_ctx* jwc_create(const int* devices, int ndev) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device available (libjwavecuda has no CPU fallback)");
    return nullptr;
  }
  std::vector<int> ords;
  if (devices == nullptr || ndev <= 0) {
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    ords.push_back(cur);
  } else {
    for (int i = 0; i < ndev; i++) {
      if (devices[i] < 0 || devices[i] >= count) {
        set_error("device ordinal %d out of range (0..%d)", devices[i], count - 1);
        return nullptr;
      }
      ords.push_back(devices[i]);
    }
  }
  jwc_ctx* ctx = new jwc_ctx();
  int prev = 0;
  cudaGetDevice(&prev);
  for (int o : ords) {
    DeviceSlot s;
    s.ordinal = o;
    cudaDeviceProp prop;
    if (cudaSetDevice(o) != cudaSuccess || cudaGetDeviceProperties(&prop, o) != cudaSuccess) {
      set_error("cannot open device %d: %s", o, cudaGetErrorString(cudaGetLastError()));
      cudaSetDevice(prev);
      jwc_destroy(ctx);
      return nullptr;
    }
    if (prop.major < 10) {
      set_error("device %d is sm_%d%d; libjwavecuda is built for sm_100a (B200) only", o, prop.major, prop.minor);
      cudaSetDevice(prev);
      jwc_destroy(ctx);
      return nullptr;
    }
    s.sm_count = prop.multiProcessorCount;
    s.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cannot create streams on device %d: %s", o, cudaGetErrorString(cudaGetLastError()));
      cudaSetDevice(prev);
      jwc_destroy(ctx);
      return nullptr;
    }
    // library scratch comes from a PRIVATE stream-ordered pool (the device's default pool, which other users of
    // cudaMallocAsync in the process share, is left alone); freed blocks stay in it until jwc_release_scratch
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = o;
    if (cudaMemPoolCreate(&s.pool, &props) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(s.pool, cudaMemPoolAttrReleaseThreshold, &thr);
    } else {
      (void)cudaGetLastError();
      s.pool = nullptr;   // Scratch falls back to the default pool
    }
    ctx->slots.push_back(s);
  }
  cudaSetDevice(prev);
  return ctx;
}

// Give the cached workspace back to the driver (all devices are synchronised first).  Safe at any quiet moment; the
// next call simply allocates again.
JWC_API int jwc_release_scratch(jwc_ctx* ctx) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  int prev = 0;
  cudaGetDevice(&prev);
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (auto& s : ctx->slots) {
    cudaSetDevice(s.ordinal);
    cudaDeviceSynchronize();
  }
  int rc = JWC_OK;
  for (auto& kv : ctx->arenas) {
    cudaSetDevice(kv.first.first);
    for (auto it = kv.second.begin(); it != kv.second.end();) {
      if (it->in_use) { ++it; rc = JWC_ERR_INVALID; continue; }   // a call is enqueueing right now: keep its blocks
      cudaFree(it->p);
      it = kv.second.erase(it);
    }
  }
  for (auto& s : ctx->slots)
    if (s.pool) cudaMemPoolTrimTo(s.pool, 0);
  cudaSetDevice(prev);
  if (rc != JWC_OK) set_error("jwc_release_scratch: calls in flight kept their workspace");
  return rc;
}

JWC_API void jwc_destroy(jwc_ctx* ctx) {
  if (!ctx) return;
  jwc_release_scratch(ctx);
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto& s : ctx->slots) {
    cudaSetDevice(s.ordinal);
    if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    for (auto& l : s.idle_lanes) jwc::lane_destroy(l);
    if (s.pool) cudaMemPoolDestroy(s.pool);
  }
  cudaSetDevice(prev);
  delete ctx;
}

JWC_API int jwc_num_devices(const jwc_ctx* ctx) { return ctx ? (int)ctx->slots.size() : 0; }
JWC_API int jwc_device_ordinal(const jwc_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) return -1;
  return ctx->slots[slot].ordinal;
}
JWC_API uint64_t jwc_launch_count(const jwc_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

static int* tuning_field(jwc_ctx* ctx, const char* key) {
  if (!ctx || !key) return nullptr;
  Tuning& t = ctx->tune;
  if (!strcmp(key, "modwt_tile")) return &t.modwt_tile;
  if (!strcmp(key, "modwt_threads")) return &t.modwt_threads;
  if (!strcmp(key, "modwt_group")) return &t.modwt_group;
  if (!strcmp(key, "modwt_smem")) return &t.modwt_smem;
  if (!strcmp(key, "dwt_tile")) return &t.dwt_tile;
  if (!strcmp(key, "dwt_threads")) return &t.dwt_threads;
  if (!strcmp(key, "dwt_group")) return &t.dwt_group;
  if (!strcmp(key, "dwt_smem")) return &t.dwt_smem;
  if (!strcmp(key, "dwt_qmf")) return &t.dwt_qmf;
  if (!strcmp(key, "h2d_chunk_mb")) return &t.h2d_chunk_mb;
  if (!strcmp(key, "h2d_buffers")) return &t.h2d_buffers;
  if (!strcmp(key, "dwt_tail")) return &t.dwt_tail;
  if (!strcmp(key, "dwt_whole")) return &t.dwt_whole;
  if (!strcmp(key, "dwt_tile_inv")) return &t.dwt_tile_inv;
  if (!strcmp(key, "wpt2d_fuse")) return &t.wpt2d_fuse;
  if (!strcmp(key, "modwt_small")) return &t.modwt_small;
  if (!strcmp(key, "small_per_cta")) return &t.small_per_cta;
  if (!strcmp(key, "force_generic")) return &t.force_generic;
  if (!strcmp(key, "l2_prefetch")) return &t.l2_prefetch;
  if (!strcmp(key, "pf_inv")) return &t.pf_inv;
  if (!strcmp(key, "modwt_logp")) return &t.modwt_logp;
  if (!strcmp(key, "top_barrier")) return &t.top_barrier;
  if (!strcmp(key, "dwt_upfront")) return &t.dwt_upfront;
  if (!strcmp(key, "dwt_k0")) return &t.dwt_k0;
  if (!strcmp(key, "dwt_fixed")) return &t.dwt_fixed;
  if (!strcmp(key, "modwt_force_wrap")) return &t.modwt_force_wrap;
  if (!strcmp(key, "modwt_threads_fwd")) return &t.modwt_threads_fwd;
  if (!strcmp(key, "modwt_plan_fwd")) return &t.modwt_plan_fwd;
  if (!strcmp(key, "modwt_plan_inv")) return &t.modwt_plan_inv;
  if (!strcmp(key, "modwt_tile_deep")) return &t.modwt_tile_deep;
  return nullptr;
}
JWC_API int jwc_set_tuning(jwc_ctx* ctx, const char* key, int value) {
  int* f = tuning_field(ctx, key);
  if (!f) { set_error("unknown tuning key '%s'", key ? key : "(null)"); return JWC_ERR_INVALID; }
  *f = value;
  return JWC_OK;
}
JWC_API int jwc_get_tuning(const jwc_ctx* ctx, const char* key, int* value) {
  int* f = tuning_field(const_cast<jwc_ctx*>(ctx), key);
  if (!f || !value) { set_error("unknown tuning key '%s'", key ? key : "(null)"); return JWC_ERR_INVALID; }
  *value = *f;
  return JWC_OK;
}

JWC_API void* jwc_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
JWC_API void jwc_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}
JWC_API void* jwc_alloc_device(jwc_ctx* ctx, int slot, size_t bytes) {
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) { set_error("bad context/slot"); return nullptr; }
  DeviceGuard guard(ctx->slots[slot].ordinal);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
JWC_API void jwc_free_device(jwc_ctx* ctx, int slot, void* p) {
  if (!ctx || !p || slot < 0 || slot >= (int)ctx->slots.size()) return;
  DeviceGuard guard(ctx->slots[slot].ordinal);
  cudaFree(p);
}
JWC_API int jwc_copy_to_device(jwc_ctx* ctx, int slot, void* dst, const void* src, size_t bytes) {
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) { set_error("bad context/slot"); return JWC_ERR_INVALID; }
  DeviceGuard guard(ctx->slots[slot].ordinal);
  cudaStream_t st = ctx->slots[slot].stream;
  JWC_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
  JWC_CUDA_CHECK(cudaStreamSynchronize(st));
  return JWC_OK;
}
JWC_API int jwc_copy_to_host(jwc_ctx* ctx, int slot, void* dst, const void* src, size_t bytes) {
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) { set_error("bad context/slot"); return JWC_ERR_INVALID; }
  DeviceGuard guard(ctx->slots[slot].ordinal);
  cudaStream_t st = ctx->slots[slot].stream;
  JWC_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
  JWC_CUDA_CHECK(cudaStreamSynchronize(st));
  return JWC_OK;
}
JWC_API int jwc_synchronize(jwc_ctx* ctx) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  for (auto& s : ctx->slots) {
    DeviceGuard guard(s.ordinal);
    JWC_CUDA_CHECK(cudaStreamSynchronize(s.stream));   // host-buffer calls drain their own lanes before returning
  }
  return JWC_OK;
}

#define JWC_DEFINE(name, OP)                                                                                        \
  JWC_API int jwc_##name(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n, int levels,         \
                         const double* f0, const double* f1, int L, unsigned flags) {                               \
    return run_host(ctx, OP, in, out, batch, n, levels, f0, f1, L, flags);                                          \
  }                                                                                                                 \
  JWC_API int jwc_##name##_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,             \
                               int64_t batch, int64_t n, int levels, const double* f0, const double* f1, int L,     \
                               unsigned flags) {                                                                    \
    return run_dev(ctx, slot, stream, OP, d_in, d_out, batch, n, levels, f0, f1, L, flags);                         \
  }

JWC_DEFINE(modwt_forward, Op::ModwtFwd)
JWC_DEFINE(modwt_inverse, Op::ModwtInv)
JWC_DEFINE(fwt_forward, Op::FwtFwd)
JWC_DEFINE(fwt_inverse, Op::FwtInv)
JWC_DEFINE(wpt_forward, Op::WptFwd)
JWC_DEFINE(wpt_inverse, Op::WptInv)

#define JWC_DEFINE_2D(name, OP)                                                                                     \
  JWC_API int jwc_##name(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t rows, int64_t cols,    \
                         int lvl_m, int lvl_n, const double* f0, const double* f1, int L, unsigned flags) {         \
    if (rows < 1) { set_error("matrix height must be >= 1 (got %lld)", (long long)rows); return JWC_ERR_INVALID; }  \
    Dim2 d2;                                                                                                        \
    d2.rows = rows;                                                                                                 \
    d2.lvl_m = lvl_m;                                                                                               \
    return run_host(ctx, OP, in, out, batch, cols, lvl_n, f0, f1, L, flags, d2);                                    \
  }                                                                                                                 \
  JWC_API int jwc_##name##_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,             \
                               int64_t batch, int64_t rows, int64_t cols, int lvl_m, int lvl_n, const double* f0,   \
                               const double* f1, int L, unsigned flags) {                                           \
    if (rows < 1) { set_error("matrix height must be >= 1 (got %lld)", (long long)rows); return JWC_ERR_INVALID; }  \
    Dim2 d2;                                                                                                        \
    d2.rows = rows;                                                                                                 \
    d2.lvl_m = lvl_m;                                                                                               \
    return run_dev(ctx, slot, stream, OP, d_in, d_out, batch, cols, lvl_n, f0, f1, L, flags, d2);                   \
  }

// 3-D: spaces [batch][p][q][r]; the reference's argument order forward(spc, lvlP, lvlQ, lvlR) with ITS use of them: the 2-D
// transform of every [q][r] matrix gets (lvlM, lvlN) = (lvlP, lvlQ), the lines along the first axis get lvlR
#define JWC_DEFINE_3D(name, OP)                                                                                     \
  JWC_API int jwc_##name(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t p, int64_t q,         \
                         int64_t r, int lvl_p, int lvl_q, int lvl_r, const double* f0, const double* f1, int L,     \
                         unsigned flags) {                                                                          \
    if (p < 1 || q < 1) { set_error("space dimensions must be >= 1 (got %lld x %lld)", (long long)p, (long long)q); return JWC_ERR_INVALID; } \
    Dim2 d2;                                                                                                        \
    d2.depth = p; d2.lvl_depth = lvl_r;                                                                             \
    d2.rows = q; d2.lvl_m = lvl_p;                                                                                  \
    return run_host(ctx, OP, in, out, batch, r, lvl_q, f0, f1, L, flags, d2);                                       \
  }                                                                                                                 \
  JWC_API int jwc_##name##_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,             \
                               int64_t batch, int64_t p, int64_t q, int64_t r, int lvl_p, int lvl_q, int lvl_r,     \
                               const double* f0, const double* f1, int L, unsigned flags) {                         \
    if (p < 1 || q < 1) { set_error("space dimensions must be >= 1 (got %lld x %lld)", (long long)p, (long long)q); return JWC_ERR_INVALID; } \
    Dim2 d2;                                                                                                        \
    d2.depth = p; d2.lvl_depth = lvl_r;                                                                             \
    d2.rows = q; d2.lvl_m = lvl_p;                                                                                  \
    return run_dev(ctx, slot, stream, OP, d_in, d_out, batch, r, lvl_q, f0, f1, L, flags, d2);                      \
  }

// CompressorMagnitude.compress: out = in where |in| >= mean|in| * threshold, else 0; *magnitude = mean|in|
JWC_API int jwc_compress_magnitude_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                       int64_t count, double threshold, double* d_magnitude) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), "device slot %d out of range", slot);
  JWC_REQUIRE(d_in != nullptr && d_out != nullptr && d_magnitude != nullptr, "NULL pointer");
  JWC_REQUIRE(count >= 0, "count must be >= 0");
  JWC_REQUIRE(threshold > 0.0, "Compressor - given threshold should be larger than zero!");
  const DeviceSlot& dev = ctx->slots[slot];
  DeviceGuard guard(dev.ordinal);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", dev.ordinal); return JWC_ERR_CUDA; }
  cudaStream_t st = stream ? (cudaStream_t)stream : dev.stream;
  return compress_magnitude(ctx, dev, st, d_in, d_out, count, threshold, d_magnitude);
}
JWC_API int jwc_compress_magnitude(jwc_ctx* ctx, const double* in, double* out, int64_t count, double threshold,
                                   double* magnitude) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(in != nullptr && out != nullptr, "NULL pointer");
  JWC_REQUIRE(count >= 0, "count must be >= 0");
  JWC_REQUIRE(threshold > 0.0, "Compressor - given threshold should be larger than zero!");
  if (magnitude) *magnitude = 0.0;
  if (count == 0) return JWC_OK;
  const DeviceSlot& dev = ctx->slots[0];   // a global mean: one device
  DeviceGuard guard(dev.ordinal);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", dev.ordinal); return JWC_ERR_CUDA; }
  Lane lane;
  if (!lane_acquire(ctx, 0, &lane)) { set_error("cannot create streams on device %d", dev.ordinal); return JWC_ERR_CUDA; }
  cudaStream_t st = lane.compute;
  double *d_x = nullptr, *d_mag = nullptr;
  int rc = JWC_OK;
  cudaError_t e;
  if ((e = jwc::pool_alloc(dev, (void**)&d_x, (size_t)count * sizeof(double), st)) != cudaSuccess ||
      (e = jwc::pool_alloc(dev, (void**)&d_mag, sizeof(double), st)) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("device staging allocation failed: %s", cudaGetErrorString(e));
    rc = JWC_ERR_NOMEM;
  }
  double mag = 0.0;
  if (rc == JWC_OK) {
    if ((e = cudaMemcpyAsync(d_x, in, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, st)) != cudaSuccess) {
      set_error("cudaMemcpyAsync(H2D) failed: %s", cudaGetErrorString(e));
      rc = JWC_ERR_CUDA;
    }
  }
  if (rc == JWC_OK) rc = compress_magnitude(ctx, dev, st, d_x, d_x, count, threshold, d_mag);   // in place
  if (rc == JWC_OK) {
    if ((e = cudaMemcpyAsync(out, d_x, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(&mag, d_mag, sizeof(double), cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
      set_error("cudaMemcpyAsync(D2H) failed: %s", cudaGetErrorString(e));
      rc = JWC_ERR_CUDA;
    }
  }
  e = cudaStreamSynchronize(st);
  if (rc == JWC_OK && e != cudaSuccess) { set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(e)); rc = JWC_ERR_CUDA; }
  if (d_x) cudaFreeAsync(d_x, st);
  if (d_mag) cudaFreeAsync(d_mag, st);
  lane_release(ctx, 0, lane);
  if (rc == JWC_OK && magnitude) *magnitude = mag;
  return rc;
}

// MODWTSlidingWindowTest.java:20-70: overlapping windows of one long series, each through forwardMODWT
JWC_API int jwc_modwt_forward_windows(jwc_ctx* ctx, const double* series, double* coeffs, int64_t series_len,
                                      int64_t window, int64_t hop, int levels, const double* g, const double* h, int L,
                                      unsigned flags) {
  if (window < 1 || hop < 1 || series_len < window) {
    set_error("need 1 <= window <= series length and hop >= 1 (window %lld, hop %lld, series %lld)", (long long)window,
              (long long)hop, (long long)series_len);
    return JWC_ERR_INVALID;
  }
  Dim2 d2;
  d2.hop = hop;
  return run_host(ctx, Op::ModwtFwd, series, coeffs, (series_len - window) / hop + 1, window, levels, g, h, L, flags, d2);
}
JWC_API int jwc_modwt_forward_windows_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_series, double* d_coeffs,
                                          int64_t series_len, int64_t window, int64_t hop, int levels, const double* g,
                                          const double* h, int L, unsigned flags) {
  if (window < 1 || hop < 1 || series_len < window) {
    set_error("need 1 <= window <= series length and hop >= 1 (window %lld, hop %lld, series %lld)", (long long)window,
              (long long)hop, (long long)series_len);
    return JWC_ERR_INVALID;
  }
  Dim2 d2;
  d2.hop = hop;
  return run_dev(ctx, slot, stream, Op::ModwtFwd, d_series, d_coeffs, (series_len - window) / hop + 1, window, levels, g,
                 h, L, flags, d2);
}

// MODWTSlidingWindowTest.java:20-70 followed by CompressorMagnitude.compress (CompressorMagnitude.java:78-90 +
// Compressor.java:97-112) over ALL coefficients of all windows: the sum of |c| is taken in the transform's store epilogue
// (one partial per CTA), so the chain moves (J+1) n written + one read + one write per window, no separate reduction
// pass.  Shapes the whole-window kernel does not take run the ordinary transform and the separate reduction.
JWC_API int jwc_modwt_forward_windows_compress_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_series,
                                                   double* d_coeffs, int64_t series_len, int64_t window, int64_t hop,
                                                   int levels, const double* g, const double* h, int L, unsigned flags,
                                                   double threshold, double* d_magnitude) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), "device slot %d out of range", slot);
  if (window < 1 || hop < 1 || series_len < window) {
    set_error("need 1 <= window <= series length and hop >= 1 (window %lld, hop %lld, series %lld)", (long long)window,
              (long long)hop, (long long)series_len);
    return JWC_ERR_INVALID;
  }
  JWC_REQUIRE(d_magnitude != nullptr, "NULL pointer");
  JWC_REQUIRE(threshold > 0.0, "Compressor - given threshold should be larger than zero!");
  const int64_t nwin = (series_len - window) / hop + 1;
  Dim2 d2;
  d2.hop = hop;
  int rc = validate(Op::ModwtFwd, d_series, d_coeffs, nwin, window, levels, g, h, L, d2);
  if (rc != JWC_OK) return rc;
  FilterPair fp;
  load_filters(fp, g, h, L);
  const DeviceSlot& dev = ctx->slots[slot];
  DeviceGuard guard(dev.ordinal);
  if (!guard.ok) { set_error("cudaSetDevice(%d) failed", dev.ordinal); return JWC_ERR_CUDA; }
  cudaStream_t st = stream ? (cudaStream_t)stream : dev.stream;
  const int64_t count = nwin * (int64_t)(levels + 1) * window;
  Scratch ws(ctx, dev, st);
  AbsSum abs;
  abs.ws = &ws;
  rc = JWC_ERR_UNSUPPORTED;
  if (!(flags & (JWC_FLAG_EXACT | JWC_FLAG_FORCE_GENERIC)) && ctx->tune.force_generic == 0)
    rc = small_modwt_forward(ctx, dev, st, d_series, d_coeffs, nwin, window, levels, fp, L, hop, &abs);
  if (rc == JWC_ERR_UNSUPPORTED) {
    rc = run_device(ctx, dev, st, Op::ModwtFwd, d_series, d_coeffs, nwin, window, levels, fp, L, flags, d2);
    if (rc != JWC_OK) return rc;
    return compress_magnitude(ctx, dev, st, d_coeffs, d_coeffs, count, threshold, d_magnitude);
  }
  if (rc != JWC_OK) return rc;
  return compress_select_from_parts(ctx, dev, st, d_coeffs, d_coeffs, count, threshold, d_magnitude, abs.parts, abs.nparts);
}

#define JWC_DEFINE_AED(name, OP)                                                                                    \
  JWC_API int jwc_##name(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n, const double* f0,   \
                         const double* f1, int L, unsigned flags) {                                                 \
    Dim2 d2;                                                                                                        \
    d2.aed = true;                                                                                                  \
    return run_host(ctx, OP, in, out, batch, n, 0, f0, f1, L, flags, d2);                                           \
  }                                                                                                                 \
  JWC_API int jwc_##name##_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,             \
                               int64_t batch, int64_t n, const double* f0, const double* f1, int L,                 \
                               unsigned flags) {                                                                    \
    Dim2 d2;                                                                                                        \
    d2.aed = true;                                                                                                  \
    return run_dev(ctx, slot, stream, OP, d_in, d_out, batch, n, 0, f0, f1, L, flags, d2);                          \
  }

JWC_DEFINE_AED(fwt_aed_forward, Op::FwtFwd)
JWC_DEFINE_AED(fwt_aed_inverse, Op::FwtInv)
JWC_DEFINE_AED(wpt_aed_forward, Op::WptFwd)
JWC_DEFINE_AED(wpt_aed_inverse, Op::WptInv)

JWC_DEFINE_2D(fwt2d_forward, Op::FwtFwd)
JWC_DEFINE_2D(fwt2d_inverse, Op::FwtInv)
JWC_DEFINE_2D(wpt2d_forward, Op::WptFwd)
JWC_DEFINE_2D(wpt2d_inverse, Op::WptInv)

JWC_DEFINE_3D(fwt3d_forward, Op::FwtFwd)
JWC_DEFINE_3D(fwt3d_inverse, Op::FwtInv)
JWC_DEFINE_3D(wpt3d_forward, Op::WptFwd)
JWC_DEFINE_3D(wpt3d_inverse, Op::WptInv)

static int dwt_split_entry(jwc_ctx* ctx, bool inverse, bool tree, const double* const* d_in, double* const* d_out,
                           int64_t n, int levels, const double* lo, const double* hi, int L, unsigned flags) {
  if (!ctx) { set_error("context is NULL"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(d_in != nullptr && d_out != nullptr && lo != nullptr && hi != nullptr, "NULL pointer");
  JWC_REQUIRE(L >= 1 && L <= JWC_MAX_TAPS, "filter length %d outside 1..%d", L, JWC_MAX_TAPS);
  JWC_REQUIRE(n >= 1 && (n & (n - 1)) == 0, "given array length is not 2^p (got %lld)", (long long)n);
  int p2 = 0;
  while (((int64_t)1 << (p2 + 1)) <= n) p2++;
  JWC_REQUIRE(levels >= 0 && levels <= p2, "given level %d is out of range for given array of length %lld", levels, (long long)n);
  JWC_REQUIRE((flags & (JWC_FLAG_EXACT | JWC_FLAG_FORCE_GENERIC)) == 0, "split transforms run on the fused kernels only");
  for (size_t p = 0; p < ctx->slots.size(); p++)
    JWC_REQUIRE(d_in[p] != nullptr && d_out[p] != nullptr, "chunk pointer %d is NULL", (int)p);
  FilterPair fp;
  load_filters(fp, lo, hi, L);
  const int rc = split_dwt(ctx, inverse, tree, d_in, d_out, n, levels, fp, L);
  if (rc == JWC_ERR_UNSUPPORTED && jwc_last_error()[0] == 0) set_error("shape outside the split FWT/WPT path");
  return rc;
}
JWC_API int jwc_fwt_forward_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags) {
  return dwt_split_entry(ctx, false, false, d_in_chunks, d_out_chunks, n, levels, lo, hi, L, flags);
}
JWC_API int jwc_fwt_inverse_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags) {
  return dwt_split_entry(ctx, true, false, d_in_chunks, d_out_chunks, n, levels, lo, hi, L, flags);
}
JWC_API int jwc_wpt_forward_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags) {
  return dwt_split_entry(ctx, false, true, d_in_chunks, d_out_chunks, n, levels, lo, hi, L, flags);
}
JWC_API int jwc_wpt_inverse_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags) {
  return dwt_split_entry(ctx, true, true, d_in_chunks, d_out_chunks, n, levels, lo, hi, L, flags);
}
JWC_API int jwc_dwt_split_levels(const jwc_ctx* ctx, int64_t n, int levels) {
  if (!ctx || n < 1 || (n & (n - 1)) || levels < 0) return -1;
  const int P = (int)ctx->slots.size();
  if (P < 1 || (P & (P - 1)) || n % P) return -1;
  int steps = 0;
  for (int64_t h = n; h >= 2 && steps < levels; h >>= 1) steps++;
  return dwt_split_levels(n, P, steps, false);
}

JWC_API int jwc_modwt_forward_split_dev(jwc_ctx* ctx, const double* const* d_x_chunks, double* const* d_coeff_chunks,
                                        int64_t n, int levels, const double* g, const double* h, int L, unsigned flags) {
  return modwt_split(ctx, false, d_x_chunks, d_coeff_chunks, n, levels, g, h, L, flags);
}
JWC_API int jwc_modwt_inverse_split_dev(jwc_ctx* ctx, const double* const* d_coeff_chunks, double* const* d_x_chunks,
                                        int64_t n, int levels, const double* g, const double* h, int L, unsigned flags) {
  return modwt_split(ctx, true, d_coeff_chunks, d_x_chunks, n, levels, g, h, L, flags);
}

}  // extern "C"
