// jwc_diag.cu -- in-run roofline denominators for bench.py: the fp64 FMA rate and the plain-copy rate of the device
// the transforms run on, measured with the library's own kernels on the caller's stream (CUDA events around the
// launches).  SURVEY.md section 8(d): the db20 / sym8 configurations are bounded by the DFMA pipe, not by HBM, so their
// fraction is reported against this number next to the HBM one.
#include "jwc_internal.cuh"

namespace jwc {
namespace {

// 16 independent accumulator chains per thread, both multiplier and addend uniform (the operand form of the filter
// inner loops: DFMA R, R, UR, R).  2 * 16 * iters flop per thread.
__global__ void __launch_bounds__(256) diag_dfma_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = (double)(threadIdx.x + i);
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chain alive, never true in practice
}

__global__ void __launch_bounds__(256) diag_copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) out[i] = in[i];
}

}  // namespace
}  // namespace jwc

using namespace jwc;

extern "C" {

JWC_API int jwc_diag_dfma_tflops(jwc_ctx* ctx, int slot, double* tflops) {
  if (!ctx || !tflops) { set_error("NULL argument"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), "device slot %d out of range", slot);
  const DeviceSlot& dev = ctx->slots[slot];
  int prev = 0;
  cudaGetDevice(&prev);
  JWC_CUDA_CHECK(cudaSetDevice(dev.ordinal));
  int rc = JWC_OK;
  double* d_out = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const int threads = 256, blocks = dev.sm_count * 8, iters = 4096;
  cudaError_t e;
  if ((e = cudaMalloc((void**)&d_out, (size_t)threads * blocks * sizeof(double))) != cudaSuccess ||
      (e = cudaEventCreate(&e0)) != cudaSuccess || (e = cudaEventCreate(&e1)) != cudaSuccess) {
    set_error("diag setup failed: %s", cudaGetErrorString(e));
    rc = JWC_ERR_CUDA;
  }
  double best = 0.0;
  for (int rep = 0; rep < 4 && rc == JWC_OK; rep++) {   // first repetition warms up
    cudaEventRecord(e0, dev.stream);
    diag_dfma_kernel<<<blocks, threads, 0, dev.stream>>>(d_out, iters, 1.0000001, 1e-9);
    count_launch(ctx);
    cudaEventRecord(e1, dev.stream);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { set_error("diag kernel failed: %s", cudaGetErrorString(e)); rc = JWC_ERR_CUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 16.0 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (d_out) cudaFree(d_out);
  cudaSetDevice(prev);
  *tflops = best;
  return rc;
}

JWC_API int jwc_diag_copy_gbs(jwc_ctx* ctx, int slot, size_t bytes, double* gbs) {
  if (!ctx || !gbs) { set_error("NULL argument"); return JWC_ERR_INVALID; }
  JWC_REQUIRE(slot >= 0 && slot < (int)ctx->slots.size(), "device slot %d out of range", slot);
  JWC_REQUIRE(bytes >= 1024, "copy size too small");
  const DeviceSlot& dev = ctx->slots[slot];
  int prev = 0;
  cudaGetDevice(&prev);
  JWC_CUDA_CHECK(cudaSetDevice(dev.ordinal));
  int rc = JWC_OK;
  const size_t n2 = bytes / sizeof(double2);
  double2 *a = nullptr, *b = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc((void**)&a, n2 * sizeof(double2))) != cudaSuccess ||
      (e = cudaMalloc((void**)&b, n2 * sizeof(double2))) != cudaSuccess ||
      (e = cudaEventCreate(&e0)) != cudaSuccess || (e = cudaEventCreate(&e1)) != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("diag setup failed: %s", cudaGetErrorString(e));
    rc = JWC_ERR_NOMEM;
  }
  double best = 0.0;
  if (rc == JWC_OK) cudaMemsetAsync(a, 0, n2 * sizeof(double2), dev.stream);
  for (int rep = 0; rep < 6 && rc == JWC_OK; rep++) {
    cudaEventRecord(e0, dev.stream);
    diag_copy_kernel<<<dev.sm_count * 32, 256, 0, dev.stream>>>(a, b, n2);
    count_launch(ctx);
    cudaEventRecord(e1, dev.stream);
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) { set_error("diag kernel failed: %s", cudaGetErrorString(e)); rc = JWC_ERR_CUDA; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = 2.0 * (double)(n2 * sizeof(double2)) / (ms * 1e-3) / 1e9;   // read + write bytes
    if (rep > 0 && g > best) best = g;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (a) cudaFree(a);
  if (b) cudaFree(b);
  cudaSetDevice(prev);
  *gbs = best;
  return rc;
}

}  // extern "C"
