// jwc_compress.cu -- magnitude thresholding of a coefficient buffer, the step that follows the transform in the
// reference's compression path.
//
// Reference: compressions/CompressorMagnitude.java:78-140 (magnitude = mean of |c| over the whole array / matrix /
// space) and compressions/Compressor.java:97-170 (keep c where |c| >= magnitude * threshold, else 0).
// Two launches: a fixed-shape tree reduction of sum |c| (per-CTA partials in a fixed order, so the result does not
// depend on scheduling; it is not the reference's left-to-right sum, the difference is O(1e-16) relative), then the
// element-wise select, which reads the magnitude from device memory -- no host round trip between the two.
#include "jwc_internal.cuh"

namespace jwc {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxParts = 1024;

__global__ void __launch_bounds__(kThreads) abs_sum_partials_kernel(const double* __restrict__ x, int64_t count,
                                                                    double* __restrict__ parts) {
  __shared__ double red[kThreads];
  // contiguous slab per CTA, strided inside it: every element has a fixed (CTA, thread, step) position
  const int64_t per = (count + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = (lo + per < count) ? lo + per : count;
  double s = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) s += fabs(x[i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = kThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) parts[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(kThreads) magnitude_kernel(const double* __restrict__ parts, int64_t nparts, int64_t count,
                                                             double* __restrict__ magnitude) {
  __shared__ double red[kThreads];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < nparts; i += kThreads) s += parts[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = kThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *magnitude = red[0] / (double)count;   // CompressorMagnitude.java:87
}

__global__ void __launch_bounds__(kThreads) threshold_kernel(const double* __restrict__ x, double* __restrict__ y,
                                                             int64_t count, const double* __restrict__ magnitude,
                                                             double threshold) {
  const double cut = *magnitude * threshold;                   // Compressor.java:104
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < count; i += (int64_t)gridDim.x * kThreads) {
    const double v = x[i];
    y[i] = (fabs(v) >= cut) ? v : 0.0;
  }
}

}  // namespace

// d_mag: one double of device memory that receives the magnitude
int compress_magnitude(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                       int64_t count, double threshold, double* d_mag) {
  if (count <= 0) return JWC_OK;
  Scratch ws(ctx, dev, st);
  int nparts = (int)((count + 65535) / 65536);
  if (nparts > kMaxParts) nparts = kMaxParts;
  if (nparts < 1) nparts = 1;
  double* parts = ws.get((size_t)nparts);
  if (!parts) { set_error("scratch allocation failed"); return JWC_ERR_NOMEM; }
  abs_sum_partials_kernel<<<nparts, kThreads, 0, st>>>(d_in, count, parts);
  magnitude_kernel<<<1, kThreads, 0, st>>>(parts, nparts, count, d_mag);
  int64_t blocks = (count + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)dev.sm_count * 16;
  if (blocks > cap) blocks = cap;
  threshold_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(d_in, d_out, count, d_mag, threshold);
  count_launch(ctx, 3);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

int compress_select_from_parts(jwc_ctx* ctx, const DeviceSlot& dev, cudaStream_t st, const double* d_in, double* d_out,
                               int64_t count, double threshold, double* d_mag, const double* d_parts, int64_t nparts) {
  if (count <= 0) return JWC_OK;
  magnitude_kernel<<<1, kThreads, 0, st>>>(d_parts, nparts, count, d_mag);
  int64_t blocks = (count + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)dev.sm_count * 16;
  if (blocks > cap) blocks = cap;
  threshold_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(d_in, d_out, count, d_mag, threshold);
  count_launch(ctx, 2);
  JWC_CUDA_CHECK(cudaGetLastError());
  return JWC_OK;
}

}  // namespace jwc
