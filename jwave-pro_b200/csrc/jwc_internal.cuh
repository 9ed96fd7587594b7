// jwc_internal.cuh -- shared declarations of libjwavecuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <list>
#include <map>
#include <vector>

#include "../../include/jwavecuda.h"

namespace jwc {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define JWC_CUDA_CHECK(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      jwc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return JWC_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)
#define JWC_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      jwc::set_error(__VA_ARGS__);   \
      return JWC_ERR_INVALID;        \
    }                                \
  } while (0)

// every even filter length the fused kernels are instantiated for.  -DJWC_ONLY_L=16 builds one length only (SASS /
// register probes while tuning: tools/sass_probe.sh); the shipped library always has them all.
#ifdef JWC_ONLY_L
#define JWC_ALL_L(X) X(JWC_ONLY_L)
#define JWC_QMF_L(X) X(JWC_ONLY_L)
#else
#define JWC_QMF_L(X) X(12) X(14) X(16) X(18) X(20)
#define JWC_ALL_L(X)                                                                                                  \
  X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(18) X(20) X(22) X(24) X(26) X(28) X(30) X(32) X(34) X(36) X(38) X(40)
#endif

// ---- filters travel as kernel parameters (constant bank): statically indexed taps become
//      immediate constant operands of DFMA, dynamically indexed ones an LDC.
struct FilterPair {
  double f0[JWC_MAX_TAPS];  // MODWT: g~   FWT/WPT: scaling (low-pass) filter of the direction
  double f1[JWC_MAX_TAPS];  // MODWT: h~   FWT/WPT: wavelet (high-pass) filter of the direction
};

struct Tuning {
  int modwt_tile = 0;       // 0 = auto
  int modwt_threads = 0;
  int modwt_group = 0;      // max levels fused per pass, 0 = auto
  int modwt_smem = 0;       // shared-memory budget per CTA in bytes, 0 = auto
  int dwt_tile = 0;
  int dwt_threads = 0;
  int dwt_group = 0;
  int dwt_smem = 0;
  int dwt_qmf = 0;          // -1 = never use the register-resident-taps (QMF) kernel variants
  int h2d_chunk_mb = 0;     // host pipeline chunk, 0 = auto
  int small_per_cta = 0;    // signals per CTA of the whole-signal MODWT kernels: 0 = auto (4 / 2 / 1 by length), else 1, 2 or 4
  int modwt_small = 0;      // whole-signal forward MODWT kernel for n <= 2048: 0 = auto, -1 = off, 1 = whenever it fits
  int wpt2d_fuse = 0;       // two column levels per launch in the 2-D packet transform (forward): 0 = auto, -1 = off
  int dwt_tile_inv = 0;     // tiled in-place kernel for the pyramid-inverse passes of long signals: 1 = on (default off: slower)
  int dwt_whole = 0;        // whole-signal in-place FWT kernel for 512 < n <= 4096: 0 = auto, -1 = off
  int dwt_tail = 0;         // warp-per-signal pyramid tail for short signals: 0 = auto, -1 = off
  int h2d_buffers = 0;      // host pipeline staging depth (1..4), 0 = auto
  int force_generic = 0;
  int l2_prefetch = 0;      // 0 = auto (one wave of CTAs ahead), -1 = off, > 0 = distance in CTAs
  int pf_inv = 0;           // L2 prefetch of the FWT / WPT inverse passes on its own: 0 = auto, -1 = off, > 0 = distance in CTAs
  int modwt_plan_fwd = 0, modwt_plan_inv = 0;   // > 0: levels per fused MODWT pass as decimal digits (2222, 431): experiments
  int modwt_threads_fwd = 0; // threads per CTA of the forward MODWT passes only (experiments)
  int modwt_force_wrap = 0; // 1 = phase-split MODWT passes always run the cycle-walk (WRAP) kernel instantiation (experiments / tests)
  int dwt_fixed = 0;        // > 0: per-pass fixed cost of the FWT planner's model, 0.01 ps per sample (experiments)
  int dwt_k0 = 0;           // > 0: levels fused by the first FWT / WPT pass (experiments; 0 = planner's choice)
  int dwt_upfront = 1;      // 1 = pyramid inverse requests every detail tile of a pass in the prologue (jwc_dwt_fast.cu); 0 = one level ahead (round 1)
  int top_barrier = 0;      // 1 = inverse tile kernels wait for their TMA tiles with one thread + a block barrier (round-1 form);
                            // >= 16 = per-thread wait with that suspend-time hint in ns (measured: no effect, r2_sweeps.txt call r7e)
  int modwt_logp = 0;       // phases per CTA (log2) of the phase-split MODWT passes: 0 = auto, 1 or 2 = forced
  int modwt_tile_deep = 0;  // tile override (decimated samples) of the phase-split MODWT passes only, 0 = auto
};

// One host-buffer call in flight owns one lane (three streams: kernels, H2D, D2H), so concurrent calls on a context --
// e.g. a forward and an inverse from two host threads, which together keep both PCIe directions busy -- never queue
// behind each other's copies.  Lanes are created on demand and recycled.
struct Lane {
  cudaStream_t compute = nullptr, copy_in = nullptr, copy_out = nullptr;
};

struct DeviceSlot {
  int ordinal = 0;
  cudaStream_t stream = nullptr;      // compute stream of the slot (device-pointer calls that pass NULL)
  int sm_count = 0;
  int max_smem_optin = 0;
  cudaMemPool_t pool = nullptr;       // private stream-ordered pool of the context on this device (library scratch)
  std::vector<Lane> idle_lanes;       // guarded by jwc_ctx::mu
};

}  // namespace jwc

namespace jwc {
// A piece of device workspace kept by the context for one (device, stream) pair.
struct ScratchBlock {
  void* p = nullptr;
  size_t bytes = 0;
  bool in_use = false;   // held by a call that is still enqueueing work
};
}  // namespace jwc

struct jwc_ctx {
  std::vector<jwc::DeviceSlot> slots;
  std::atomic<uint64_t> launches{0};
  std::atomic<int> host_calls{0};   // host-buffer calls in flight (copy pacing, see run_host_slot)
  jwc::Tuning tune;
  std::mutex mu;
  // workspace arenas, guarded by mu (std::list: blocks keep their address while others are added)
  std::map<std::pair<int, cudaStream_t>, std::list<jwc::ScratchBlock>> arenas;
};

namespace jwc {

// Bound on the number of (device, stream) arenas a context keeps: beyond it, arenas without a block in use are given
// back to the driver (device-synchronising, so rare by construction).  Caller holds ctx->mu.
constexpr size_t kMaxArenas = 48;
inline void evict_idle_arenas(jwc_ctx* ctx) {
  int prev = 0;
  cudaGetDevice(&prev);
  int synced = -1;
  for (auto it = ctx->arenas.begin(); it != ctx->arenas.end();) {
    bool busy = false;
    for (auto& b : it->second) busy = busy || b.in_use;
    if (busy) { ++it; continue; }
    if (synced != it->first.first) {   // map is ordered by device: one synchronisation per device
      cudaSetDevice(it->first.first);
      cudaDeviceSynchronize();
      synced = it->first.first;
    }
    for (auto& b : it->second) cudaFree(b.p);
    it = ctx->arenas.erase(it);
  }
  cudaSetDevice(prev);
}

// Workspace of one call.  Blocks come from the context's arena of the call's (device, stream) and go back to it when
// the call has finished ENQUEUEING: everything that uses a block is ordered on that one stream, so the next call on
// the stream may reuse it at once, and a call that is enqueueing concurrently from another host thread never gets a
// block that is still held (re-entrancy requirement of SURVEY.md section 8b "threading").  In steady state a call
// makes no allocator call at all -- per-call cudaMallocAsync / cudaFreeAsync pairs measured 2-5x slower transforms
// whenever something polled the driver (nvidia-smi -lms 20 next to a 2-D FWT: 3.6 ms -> 7.7-17 ms).  Blocks live until
// jwc_release_scratch / jwc_destroy.
struct Scratch {
  jwc_ctx* ctx;
  int ordinal;
  cudaStream_t stream;
  cudaMemPool_t pool;
  std::vector<ScratchBlock*> held;
  Scratch(jwc_ctx* c, const DeviceSlot& dev, cudaStream_t s) : ctx(c), ordinal(dev.ordinal), stream(s), pool(dev.pool) {}
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
  double* get(size_t n_doubles) {
    if (n_doubles == 0) n_doubles = 1;
    const size_t bytes = (n_doubles * sizeof(double) + 255) & ~(size_t)255;
    {
      std::lock_guard<std::mutex> lk(ctx->mu);
      ScratchBlock* best = nullptr;
      for (auto& b : ctx->arenas[std::make_pair(ordinal, stream)])
        if (!b.in_use && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
      if (best) {
        best->in_use = true;
        held.push_back(best);
        return static_cast<double*>(best->p);
      }
    }
    void* p = nullptr;
    const cudaError_t e = pool ? cudaMallocFromPoolAsync(&p, bytes, pool, stream) : cudaMallocAsync(&p, bytes, stream);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    // callers that come with ever new (short-lived) streams would otherwise pile up cached blocks without bound
    if (ctx->arenas.size() >= kMaxArenas && !ctx->arenas.count(std::make_pair(ordinal, stream))) evict_idle_arenas(ctx);
    auto& arena = ctx->arenas[std::make_pair(ordinal, stream)];
    arena.push_back(ScratchBlock{p, bytes, true});
    held.push_back(&arena.back());
    return static_cast<double*>(p);
  }
  void release() {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (ScratchBlock* b : held) b->in_use = false;
    held.clear();
  }
  ~Scratch() { release(); }
};

// ---- generic (any shape, one level per launch) kernels: jwc_generic.cu ---------------------------------
int generic_modwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_x, double* d_coeffs,
                          int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool exact);
// continue a forward MODWT from V_{first_level-1} (d_v, signal stride v_sig) for levels first_level .. levels
int generic_modwt_forward_from(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_v, int64_t v_sig,
                               int first_level, double* d_coeffs, int64_t batch, int64_t n, int levels,
                               const FilterPair& f, int L, bool exact = false);
int generic_modwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_coeffs, double* d_x,
                          int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool exact);
// tree = false: FWT (only the length-h prefix is transformed each level); tree = true: WPT (every block).
// ld = distance in doubles between consecutive signals in d_in and in d_out (0 = dense, ld = n): lets a transform run
// on a column block of a wider array (Ancient-Egyptian blocks of arbitrary-length signals).
int generic_dwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, bool exact,
                        int64_t ld = 0);
int generic_dwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, bool exact,
                        int64_t ld = 0);

// ---- fused tile kernels: jwc_modwt_fast.cu / jwc_dwt_fast.cu --------------------------------------------
// each returns JWC_ERR_UNSUPPORTED (without setting an error) when the shape is outside its fast path.
// x_sig: distance between consecutive input signals (0 = n); x_sig < n makes the batch a set of overlapping windows of
// one series (sliding-window analysis without materialising the windows)
int fast_modwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_x, double* d_coeffs,
                       int64_t batch, int64_t n, int levels, const FilterPair& f, int L, int64_t x_sig = 0);
int fast_modwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_coeffs, double* d_x,
                       int64_t batch, int64_t n, int levels, const FilterPair& f, int L);
int small_modwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_coeffs, double* d_x,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L);
// whole-signal in-place FWT for 512 < n <= 4096 (jwc_dwt_whole.cu)
int whole_dwt_levels(const jwc_ctx* ctx, int64_t n, int steps, int L, bool tree = false);
int whole_dwt(jwc_ctx* ctx, cudaStream_t st, const double* d_in, double* d_out, int64_t batch, int64_t n, int steps,
              const FilterPair& f, int L, int64_t ld, bool inverse, const double* d_prefix = nullptr,
              int64_t prefix_sig = 0, bool tree = false);
int tile_dwt_inverse_pass(jwc_ctx* ctx, cudaStream_t st, const double* ain, int64_t ain_sig, const double* din,
                          int64_t din_sig, double* out, int64_t out_sig, int64_t N, int l0, int k, int64_t batch,
                          const FilterPair& f, int L);
// whole-signal-in-shared-memory forward MODWT for short signals / analysis windows (jwc_modwt_small.cu)
// abs != nullptr: the kernel also leaves sum |coefficient| of each CTA in abs->parts (allocated from abs->ws)
struct AbsSum {
  Scratch* ws = nullptr;
  double* parts = nullptr;
  int64_t nparts = 0;
  bool fused = false;
};
int small_modwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_x, double* d_coeffs,
                        int64_t batch, int64_t n, int levels, const FilterPair& f, int L, int64_t x_sig = 0,
                        AbsSum* abs = nullptr);
// warp-per-signal deep end of the FWT pyramid for short signals (jwc_dwt_tail.cu)
constexpr int kDwtTailLen = 512;        // block length at which the tail takes over
constexpr int64_t kDwtTailMaxN = 16384; // longest signal that uses it
int dwt_tail_start(int64_t n, int steps);
int dwt_tail_forward(jwc_ctx* ctx, cudaStream_t st, const double* src, int64_t src_sig, double* d_out, int64_t n,
                     int h0, int nlev, int64_t batch, const FilterPair& f, int L);
int dwt_tail_inverse(jwc_ctx* ctx, cudaStream_t st, const double* d_in, int64_t n, double* dst, int64_t dst_sig, int h0,
                     int nlev, int64_t batch, const FilterPair& f, int L);

// magnitude thresholding of a coefficient buffer (jwc_compress.cu); d_mag = one device double for the magnitude
int compress_magnitude(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                       int64_t count, double threshold, double* d_mag);
// same, with the sum of |c| already available as `nparts` partial sums in d_parts (fused into the transform's stores):
// the magnitude is their fixed-order total / count, then ONE select pass
int compress_select_from_parts(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                               int64_t count, double threshold, double* d_mag, const double* d_parts, int64_t nparts);

// column passes of the 2-D FWT / WPT (jwc_dwt2d.cu); d_src / d_in must not overlap the destination
int dwt2d_column_steps(int64_t rows, int levels);
int dwt2d_columns_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_src, double* d_out,
                          int64_t batch, int64_t rows, int64_t cols, int levels, const FilterPair& f, int L, bool tree,
                          bool exact);
int dwt2d_columns_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_dst,
                          int64_t batch, int64_t rows, int64_t cols, int levels, const FilterPair& f, int L, bool tree,
                          bool exact);

int fast_dwt_forward(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                     int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, int64_t ld = 0);
int fast_dwt_inverse(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                     int64_t batch, int64_t n, int levels, const FilterPair& f, int L, bool tree, int64_t ld = 0);

// one long series split over the context's devices: FWT / WPT (jwc_dwt_fast.cu); chunk p = samples [p n/P, (p+1) n/P)
int dwt_split_levels(int64_t n, int P, int steps, bool tree);
int split_dwt(jwc_ctx* ctx, bool inverse, bool tree, const double* const* d_in, double* const* d_out, int64_t n,
              int levels, const FilterPair& f, int L);

// The dynamic-shared-memory cap of a kernel is per-function state shared by every host thread: setting it to "what this
// launch needs" races with a concurrent launch of the same instantiation that needs more (found by the concurrent
// device-call test: cudaErrorInvalidValue at launch).  So every kernel gets the SAME cap once per device: the opt-in
// maximum minus its static shared memory; the carve-out actually used still follows each launch's own request.
inline cudaError_t allow_max_dynamic_smem_impl(const void* kern) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, bool> done;   // (kernel, device ordinal)
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(mu);
    if (done.count(std::make_pair(kern, dev))) return cudaSuccess;
  }
  int optin = 0;
  cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> lk(mu);
    done[std::make_pair(kern, dev)] = true;
  }
  return e;
}
template <typename Kern>
inline cudaError_t allow_max_dynamic_smem(Kern kern) {
  return allow_max_dynamic_smem_impl(reinterpret_cast<const void*>(kern));
}

inline void count_launch(jwc_ctx* ctx, uint64_t k = 1) { ctx->launches.fetch_add(k, std::memory_order_relaxed); }

template <bool EXACT>
__device__ __forceinline__ double mac(double acc, double a, double b) {
  if (EXACT) return __dadd_rn(acc, __dmul_rn(a, b));  // two roundings, like the JVM
  return fma(a, b, acc);
}

}  // namespace jwc
