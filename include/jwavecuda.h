/* jwavecuda.h -- C ABI of libjwavecuda.so: the B200 (sm_100a) wavelet filter-bank engine that sits
 * behind JWave-Pro's BasicTransform API for three transforms:
 *
 *   MODWT  (a-trous circular convolution)   replaces  transforms/MODWTTransform.java:256-306 forwardMODWT,
 *                                                      :337-375 inverseMODWT (direct method :677-716)
 *   FWT    (decimated periodic pyramid)     replaces  transforms/FastWaveletTransform.java:71-101,119-153
 *                                                      + transforms/wavelets/Wavelet.java:236-303
 *   WPT    (full packet tree)               replaces  transforms/WaveletPacketTransform.java:73-124,141-191
 *
 * (paths relative to /root/reference/src/main/java/jwave/).  The reference has no FFI of its own; these
 * entry points are what its Java subclasses CudaMODWTTransform / CudaFastWaveletTransform /
 * CudaWaveletPacketTransform bind through Panama FFM (java/…/JwcNative.java), and what the
 * Python ctypes mirror in jwave-pro_b200/ binds for the tests.
 *
 * Conventions
 *  - plain C, no CUDA/torch types; every size is int64_t or int; all arrays are IEEE fp64, row-major, dense.
 *  - return value: 0 = JWC_OK, negative = error (jwc_last_error() gives the text, thread-local).
 *    Nothing here aborts, exits or throws across the boundary.
 *  - There is NO CPU fallback: every transform runs as CUDA kernels; without a usable device
 *    jwc_create() returns NULL.
 *  - Batches: `batch` independent signals of length n each; signal b starts at in + b*n.
 *  - Input and output buffers of one call must not overlap (the reference never mutates its input either).
 *  - A context may be used from several host threads at once (calls are re-entrant; scratch memory is
 *    stream-ordered per call).  Filters are passed per call; the library keeps no filter state.
 */
#ifndef JWAVECUDA_H
#define JWAVECUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define JWC_API __attribute__((visibility("default")))
#else
#define JWC_API
#endif

typedef struct jwc_ctx jwc_ctx;

enum {
  JWC_OK = 0,
  JWC_ERR_INVALID = -1,     /* bad argument (NULL pointer, n < 1, level out of range, L > JWC_MAX_TAPS ...) */
  JWC_ERR_CUDA = -2,        /* a CUDA runtime call failed; text in jwc_last_error() */
  JWC_ERR_NOMEM = -3,       /* device or pinned allocation failed */
  JWC_ERR_UNSUPPORTED = -4  /* shape outside what the build supports */
};

/* flags (bit-or) */
#define JWC_FLAG_EXACT 1u         /* unfused multiply/add in the reference's summation order: results are
                                     bit-identical to the JVM arithmetic (slower kernels).  Default uses FMA;
                                     then |err| <= 1e-12*max|x| versus the reference. */
#define JWC_FLAG_FORCE_GENERIC 2u /* one kernel per level, no tile fusion (debug / cross-check) */

#define JWC_MAX_TAPS 64           /* longest filter accepted (reference max on the path: Daubechies20 = 40) */
#define JWC_MODWT_MAX_LEVEL 13    /* transforms/MODWTTransform.java:111 MAX_DECOMPOSITION_LEVEL */

/* ---- context ---------------------------------------------------------------------------------------- */

/* devices: CUDA ordinals the context may use (host-buffer calls shard the batch over all of them, by
 * signal, no collective); devices == NULL or ndev <= 0 means "the current device only". */
JWC_API jwc_ctx* jwc_create(const int* devices, int ndev);
JWC_API void jwc_destroy(jwc_ctx* ctx);
JWC_API int jwc_num_devices(const jwc_ctx* ctx);
JWC_API int jwc_device_ordinal(const jwc_ctx* ctx, int slot);
JWC_API const char* jwc_last_error(void);
JWC_API const char* jwc_version(void);
/* number of CUDA kernels this context has launched so far (all slots) */
JWC_API uint64_t jwc_launch_count(const jwc_ctx* ctx);
/* tuning knobs for sweeps ("modwt_tile", "modwt_threads", ...); unknown key -> JWC_ERR_INVALID */
JWC_API int jwc_set_tuning(jwc_ctx* ctx, const char* key, int value);
JWC_API int jwc_get_tuning(const jwc_ctx* ctx, const char* key, int* value);

/* ---- memory helpers (pinned host staging = what the Java side wraps in MemorySegments) --------------- */
JWC_API void* jwc_alloc_pinned(size_t bytes);
JWC_API void jwc_free_pinned(void* p);
JWC_API void* jwc_alloc_device(jwc_ctx* ctx, int slot, size_t bytes);
JWC_API void jwc_free_device(jwc_ctx* ctx, int slot, void* p);
JWC_API int jwc_copy_to_device(jwc_ctx* ctx, int slot, void* dst_dev, const void* src_host, size_t bytes);
JWC_API int jwc_copy_to_host(jwc_ctx* ctx, int slot, void* dst_host, const void* src_dev, size_t bytes);
JWC_API int jwc_synchronize(jwc_ctx* ctx);
/* Device workspace (intermediate levels, host-pipeline staging) is cached per (device, stream) inside the context so
 * that a transform call makes no allocator call in steady state; this returns it to the driver (synchronises the
 * context's devices).  jwc_destroy does the same. */
JWC_API int jwc_release_scratch(jwc_ctx* ctx);

/* ---- MODWT ------------------------------------------------------------------------------------------
 * g, h: the level-1 MODWT filters g~ = (scalingDeCom/||.||)/sqrt2, h~ = (waveletDeCom/||.||)/sqrt2
 *       (MODWTTransform.java:462-475), L taps each; level-j upsampling (:618-630) is implicit.
 * forward:  x [batch][n]  ->  coeffs [batch][levels+1][n], rows W_1..W_J, V_J   (:298-303, flat form :406-416)
 *           W_j[t] = sum_m h[m] V_{j-1}[(t - m 2^(j-1)) mod n],  V_j likewise with g           (:677-690)
 * inverse:  coeffs -> x,  V_{j-1}[t] = sum_m g[m] V_j[(t + m 2^(j-1)) mod n] + sum_m h[m] W_j[...]  (:703-716, :363-369)
 * Any n >= 1 (not only 2^p); 1 <= levels <= min(13, floor(log2 n)) as the reference enforces (:257-282) --
 * the library itself only requires levels >= 1 and (L-1)*2^(levels-1) representable; callers validate.
 * Host-pointer variants shard the batch over the context's devices and pipeline H2D / kernels / D2H.   */
JWC_API int jwc_modwt_forward(jwc_ctx* ctx, const double* x, double* coeffs, int64_t batch, int64_t n, int levels,
                              const double* g, const double* h, int L, unsigned flags);
JWC_API int jwc_modwt_inverse(jwc_ctx* ctx, const double* coeffs, double* x, int64_t batch, int64_t n, int levels,
                              const double* g, const double* h, int L, unsigned flags);
/* device-pointer variants: buffers live on device `slot` of the context; `stream` is a cudaStream_t
 * (NULL = the context's own stream for that slot); asynchronous with respect to the host. */
JWC_API int jwc_modwt_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_x, double* d_coeffs,
                                  int64_t batch, int64_t n, int levels, const double* g, const double* h, int L,
                                  unsigned flags);
JWC_API int jwc_modwt_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_coeffs, double* d_x,
                                  int64_t batch, int64_t n, int levels, const double* g, const double* h, int L,
                                  unsigned flags);

/* One long series split over the context's devices (a series too long, or too slow, for one GPU): device slot p holds
 * the contiguous chunk [n*p/P, n*(p+1)/P) of x in d_x_chunks[p] and of every coefficient row in d_coeff_chunks[p]
 * ([levels+1][chunk_len], row-major).  One halo exchange of (L-1)(2^levels - 1) samples per transform between ring
 * neighbours (cudaMemcpyPeerAsync over NVLink), no collective.  Every chunk must be at least one halo long.  The chunk
 * buffers must be ready (producers synchronised) on entry; the call returns after all devices have finished. */
JWC_API int jwc_modwt_forward_split_dev(jwc_ctx* ctx, const double* const* d_x_chunks, double* const* d_coeff_chunks,
                                        int64_t n, int levels, const double* g, const double* h, int L, unsigned flags);
JWC_API int jwc_modwt_inverse_split_dev(jwc_ctx* ctx, const double* const* d_coeff_chunks, double* const* d_x_chunks,
                                        int64_t n, int levels, const double* g, const double* h, int L, unsigned flags);

/* FWT / WPT of one long series split over the context's devices (the decimated counterpart of the two calls above;
 * reference loops: transforms/FastWaveletTransform.java:85-99,133-151, transforms/WaveletPacketTransform.java:98-120,167-187).
 * n = 2^m, P = number of device slots = 2^q, chunk p = samples [p n/P, (p+1) n/P) on slot p.  The data never moves: after
 * l levels slot p holds the (n/P)/2^l coefficients of every node whose support starts in its chunk; a fused pass pulls
 * (L-2)(2^k-1) samples (forward, right neighbour's head) or <= L-2 coefficients per child array (inverse, left
 * neighbour's tail) from the ring neighbour with cudaMemcpyPeerAsync.  No collective.
 * Output ("local layout", input of the inverse): chunk p is the transform's own layout of a signal of length n/P filled
 * with slot p's coefficients --
 *   WPT:  [leaf 0 part | leaf 1 part | ...]; part c = leaf c of the reference at positions [p len_c/P, (p+1) len_c/P)
 *   FWT:  [T_p | D_ls part | ... | D_1 part], D_l part = D_l[p (n/P)/2^l, (p+1) (n/P)/2^l); ls = jwc_dwt_split_levels()
 *         levels run split (input part >= 2048 samples); the remaining pyramid on A_ls (n/2^ls samples, the small
 *         remainder) is gathered on slot 0, transformed there, and its array [A_J | D_J .. D_{ls+1}] cut into P equal
 *         contiguous pieces T_p.
 * Concatenating the parts of a band over p gives that band of the unsplit result bit for bit.  WPT needs
 * levels <= jwc_dwt_split_levels (packet parts of at least 1024 samples per device).  Chunk buffers must be ready on entry;
 * the calls return after all devices have finished. */
JWC_API int jwc_fwt_forward_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt_inverse_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_forward_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_inverse_split_dev(jwc_ctx* ctx, const double* const* d_in_chunks, double* const* d_out_chunks, int64_t n,
                                      int levels, const double* lo, const double* hi, int L, unsigned flags);
/* number of levels of an n-sample, `levels`-level transform that run split on this context (-1: shape not splittable) */
JWC_API int jwc_dwt_split_levels(const jwc_ctx* ctx, int64_t n, int levels);

/* Sliding-window analysis (the reference's motivating workload, test/.../MODWTSlidingWindowTest.java:20-70: 512-sample
 * windows, 8 levels, step 64): window w = series[w*hop .. w*hop + window), w = 0 .. (series_len - window)/hop, each
 * through forwardMODWT.  The windows are never materialised: the kernels read window w at series + w*hop.
 * coeffs: [nwin][levels+1][window].  An odd hop forces the element-wise loaders (no 16-byte alignment). */
JWC_API int jwc_modwt_forward_windows(jwc_ctx* ctx, const double* series, double* coeffs, int64_t series_len,
                                      int64_t window, int64_t hop, int levels, const double* g, const double* h, int L,
                                      unsigned flags);
JWC_API int jwc_modwt_forward_windows_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_series, double* d_coeffs,
                                          int64_t series_len, int64_t window, int64_t hop, int levels, const double* g,
                                          const double* h, int L, unsigned flags);

/* ---- FWT --------------------------------------------------------------------------------------------
 * lo, hi: scalingDeCom / waveletDeCom for forward, scalingReCon / waveletReCon for inverse (Wavelet.java:178-219).
 * n must be 2^p, 0 <= levels <= p (FastWaveletTransform.java:74-83); levels = 0 copies.
 * forward: in [batch][n] -> out [batch][n] in the reference's in-place layout [A_J | D_J | ... | D_1]
 *          one step on a length-h prefix: lo[i] = sum_j x[(2i+j) mod h] lo[j], hi likewise (Wavelet.java:236-260)
 * inverse: synthesis out[(2i+j) mod h] += c[i] lo[j] + c[i+h/2] hi[j] (Wavelet.java:277-303), h = 2n/2^levels .. n */
JWC_API int jwc_fwt_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n, int levels,
                            const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n, int levels,
                            const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                int64_t batch, int64_t n, int levels, const double* lo, const double* hi, int L,
                                unsigned flags);
JWC_API int jwc_fwt_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                int64_t batch, int64_t n, int levels, const double* lo, const double* hi, int L,
                                unsigned flags);

/* ---- WPT --------------------------------------------------------------------------------------------
 * Same step, applied to every aligned block of length h = n/2^l at level l (WaveletPacketTransform.java:98-120,
 * :167-187); leaves in natural (Paley) order.  n = 2^p, 0 <= levels <= p. */
JWC_API int jwc_wpt_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n, int levels,
                            const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n, int levels,
                            const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                int64_t batch, int64_t n, int levels, const double* lo, const double* hi, int L,
                                unsigned flags);
JWC_API int jwc_wpt_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                int64_t batch, int64_t n, int levels, const double* lo, const double* hi, int L,
                                unsigned flags);

/* ---- 2-D FWT / WPT ----------------------------------------------------------------------------------
 * The reference's matrix overloads: transforms/BasicTransform.java:361-399 forward(double[][], lvlM, lvlN) sends every
 * row through forward(row, lvlN) and then every column of the result through forward(col, lvlM); :436-474
 * reverse(double[][], lvlM, lvlN) undoes the columns (lvlM) first, then the rows (lvlN).
 * transforms/ParallelTransform.java:70-91,222-271 is the same arithmetic with rows spread over host threads.
 * in, out: [batch][rows][cols] row-major, rows and cols both 2^p, 0 <= lvl_m <= log2(rows), 0 <= lvl_n <= log2(cols);
 * lo/hi as for the 1-D calls (DeCom for forward, ReCon for inverse).  The column pass runs in place on the row-major
 * matrix (no transposes); the host variants shard the batch of matrices over the context's devices. */
JWC_API int jwc_fwt2d_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t rows, int64_t cols,
                              int lvl_m, int lvl_n, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt2d_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t rows, int64_t cols,
                              int lvl_m, int lvl_n, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt2d_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t rows, int64_t cols,
                              int lvl_m, int lvl_n, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt2d_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t rows, int64_t cols,
                              int lvl_m, int lvl_n, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt2d_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                  int64_t batch, int64_t rows, int64_t cols, int lvl_m, int lvl_n, const double* lo,
                                  const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt2d_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                  int64_t batch, int64_t rows, int64_t cols, int lvl_m, int lvl_n, const double* lo,
                                  const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt2d_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                  int64_t batch, int64_t rows, int64_t cols, int lvl_m, int lvl_n, const double* lo,
                                  const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt2d_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                  int64_t batch, int64_t rows, int64_t cols, int lvl_m, int lvl_n, const double* lo,
                                  const double* hi, int L, unsigned flags);

/* ---- 3-D FWT / WPT ----------------------------------------------------------------------------------
 * The reference's space overloads: transforms/BasicTransform.java:487-565 forward(double[][][] spcTime, lvlP, lvlQ, lvlR)
 * sends every matrix spcTime[i] ([q][r]) through the 2-D forward(mat, lvlP, lvlQ) -- i.e. its rows (length r) with lvlQ
 * levels and its columns (length q) with lvlP levels -- and then every line along the first axis (length p) through the
 * 1-D forward(line, lvlR); :579-640 reverse(double[][][], lvlP, lvlQ, lvlR) keeps that order (2-D reverse of every
 * matrix first, then the first axis).  The level arguments are passed through exactly as the reference uses them, so
 * lvl_p must be <= log2(q), lvl_q <= log2(r) and lvl_r <= log2(p) (for a cube: any level up to log2 of the edge).
 * in, out: [batch][p][q][r] row-major, p, q, r all 2^p.  The first-axis pass is the 2-D column pass on the view
 * [batch][p][q * r], in place on the row-major space (no transposes). */
JWC_API int jwc_fwt3d_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r,
                              int lvl_p, int lvl_q, int lvl_r, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt3d_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r,
                              int lvl_p, int lvl_q, int lvl_r, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt3d_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r,
                              int lvl_p, int lvl_q, int lvl_r, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt3d_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t p, int64_t q, int64_t r,
                              int lvl_p, int lvl_q, int lvl_r, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt3d_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out, int64_t batch,
                                  int64_t p, int64_t q, int64_t r, int lvl_p, int lvl_q, int lvl_r, const double* lo,
                                  const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt3d_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out, int64_t batch,
                                  int64_t p, int64_t q, int64_t r, int lvl_p, int lvl_q, int lvl_r, const double* lo,
                                  const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt3d_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out, int64_t batch,
                                  int64_t p, int64_t q, int64_t r, int lvl_p, int lvl_q, int lvl_r, const double* lo,
                                  const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt3d_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out, int64_t batch,
                                  int64_t p, int64_t q, int64_t r, int lvl_p, int lvl_q, int lvl_r, const double* lo,
                                  const double* hi, int L, unsigned flags);

/* ---- magnitude thresholding of a coefficient buffer ---------------------------------------------------------
 * compressions/CompressorMagnitude.java:78-140 + compressions/Compressor.java:97-170: magnitude = mean |c| over all
 * `count` values (array, matrix or space alike); out[i] = in[i] if |in[i]| >= magnitude * threshold, else 0.
 * threshold > 0.  The device variant leaves the magnitude in *d_magnitude (device memory) and may run in place
 * (d_out == d_in); chained after a *_dev transform on the same stream it costs one read for the sum and one read +
 * write for the select, with no host round trip. */
JWC_API int jwc_compress_magnitude(jwc_ctx* ctx, const double* in, double* out, int64_t count, double threshold,
                                   double* magnitude);
JWC_API int jwc_compress_magnitude_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                       int64_t count, double threshold, double* d_magnitude);

/* The sliding-window transform and the thresholding that follows it in the reference's compression path, chained on the
 * device: coeffs [nwin][levels+1][window] as jwc_modwt_forward_windows_dev writes them, then CompressorMagnitude over ALL
 * of them, in place.  The sum of |c| is taken in the transform's store epilogue (per-CTA partial sums, fixed reduction
 * order), so no separate reduction pass reads the coefficients again; *d_magnitude (device memory) receives mean |c|. */
JWC_API int jwc_modwt_forward_windows_compress_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_series,
                                                   double* d_coeffs, int64_t series_len, int64_t window, int64_t hop,
                                                   int levels, const double* g, const double* h, int L, unsigned flags,
                                                   double threshold, double* d_magnitude);

/* ---- arbitrary-length FWT / WPT: Ancient-Egyptian decomposition -----------------------------------------
 * transforms/AncientEgyptianDecomposition.java:97-181 with tools/MathToolKit.java:57-84 decompose(): a signal of any
 * length n >= 1 is cut into blocks of descending powers of two (42 = 32 | 8 | 2); every block is transformed on its own
 * by the wrapped transform at FULL depth (forward(double[]) of a 2^p block = p levels; a block of length 1 is copied)
 * and written back at its position.  in, out: [batch][n]; every block is one batched launch sequence over all signals
 * (row stride n), nothing is gathered or copied.  lo/hi as for the 1-D calls. */
JWC_API int jwc_fwt_aed_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n,
                                const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt_aed_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n,
                                const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_aed_forward(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n,
                                const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_aed_inverse(jwc_ctx* ctx, const double* in, double* out, int64_t batch, int64_t n,
                                const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt_aed_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                    int64_t batch, int64_t n, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_fwt_aed_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                    int64_t batch, int64_t n, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_aed_forward_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                    int64_t batch, int64_t n, const double* lo, const double* hi, int L, unsigned flags);
JWC_API int jwc_wpt_aed_inverse_dev(jwc_ctx* ctx, int slot, void* stream, const double* d_in, double* d_out,
                                    int64_t batch, int64_t n, const double* lo, const double* hi, int L, unsigned flags);

/* ---- diagnostics: roofline denominators measured with the library's own kernels on device `slot` ---------------
 * jwc_diag_dfma_tflops: sustained fp64 FMA rate (TFLOP/s, 16 independent chains per thread, uniform operands -- the
 * operand form of the filter inner loops); jwc_diag_copy_gbs: plain device-to-device copy of `bytes` bytes (read +
 * write GB/s).  bench.py reports the fp64-bound configurations (Daubechies20 MODWT, Symlet8 WPT) against the first
 * (SURVEY.md section 8d).  Both synchronise the slot's stream. */
JWC_API int jwc_diag_dfma_tflops(jwc_ctx* ctx, int slot, double* tflops);
JWC_API int jwc_diag_copy_gbs(jwc_ctx* ctx, int slot, size_t bytes, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* JWAVECUDA_H */
