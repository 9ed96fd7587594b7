// tools/microbench.cu -- B200 denominators this project needs beyond MEASURED_PEAKS.json:
//   fp64 FMA peak (second roofline for the db20 / WPT configs), shared-memory LDS.64 bandwidth, plain copy and
//   1-D bulk (TMA, cp.async.bulk) copy bandwidth.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// DFMA with one operand from shared memory per 2 FMAs (the a-trous inner loop without register reuse)
__global__ void dfma_lds_kernel(double* out, int iters, double a, double b, int reuse) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0, acc6 = 0, acc7 = 0;
  int idx = threadIdx.x;
  if (reuse == 1) {
    for (int i = 0; i < iters; i++) {
      double v = sm[(idx + i) & 4095];
      acc0 = fma(v, a, acc0); acc1 = fma(v, b, acc1);
    }
  } else if (reuse == 4) {
    for (int i = 0; i < iters; i++) {
      double v = sm[(idx + i) & 4095];
      acc0 = fma(v, a, acc0); acc1 = fma(v, b, acc1); acc2 = fma(v, a, acc2); acc3 = fma(v, b, acc3);
      acc4 = fma(v, a, acc4); acc5 = fma(v, b, acc5); acc6 = fma(v, a, acc6); acc7 = fma(v, b, acc7);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3 + acc4 + acc5 + acc6 + acc7;
}

__global__ void lds_kernel(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc = 0;
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; u++) acc += sm[(idx + (i + u) * 32) & 4095];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n2) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) out[i] = in[i];
}

// one read, `nw` writes (the MODWT forward traffic shape: 1 : J+1)
__global__ void fanout_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n2, int nw) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 v = in[i];
    for (int k = 0; k < nw; k++) { out[(size_t)k * n2 + i] = v; v.x += 1.0; }
  }
}

// bulk (TMA 1-D) copy: global -> smem -> global, TILE bytes per CTA iteration
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int TILE_BYTES>
__global__ void bulk_copy_kernel(const char* __restrict__ in, char* __restrict__ out, size_t bytes) {
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = smem_u32(&bar), buf_a = smem_u32(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  uint32_t phase = 0;
  size_t ntiles = bytes / TILE_BYTES;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(TILE_BYTES));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(buf_a), "l"(in + t * TILE_BYTES), "r"(TILE_BYTES), "r"(bar_a) : "memory");
      // wait
      asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}"
                   ::"r"(bar_a), "r"(phase) : "memory");
      asm volatile("fence.proxy.async.shared::cta;");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + t * TILE_BYTES), "r"(buf_a), "r"(TILE_BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    phase ^= 1;
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d smem/block optin=%zu clock=%d kHz\n", prop.name, prop.major, prop.minor,
         prop.multiProcessorCount, prop.sharedMemPerBlockOptin, prop.clockRate);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  double* dout;
  CK(cudaMalloc(&dout, sizeof(double) * 148 * 16 * 1024));
  const int sms = prop.multiProcessorCount;
  // --- DFMA peak
  for (int blocks_per_sm : {1, 2, 4}) {
    int iters = 20000;
    dfma_kernel<<<sms * blocks_per_sm, 512>>>(dout, 100, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dfma_kernel<<<sms * blocks_per_sm, 512>>>(dout, iters, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * 8 * iters * 512.0 * sms * blocks_per_sm;
    printf("dfma: %d CTA/SM x512 thr: %.2f TFLOP/s fp64 (%.3f ms)\n", blocks_per_sm, fl / ms / 1e9, ms);
  }
  CK(cudaFuncSetAttribute(dfma_lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  for (int reuse : {1, 4}) {
    int iters = 20000;
    CK(cudaEventRecord(e0));
    dfma_lds_kernel<<<sms * 2, 512, 32768>>>(dout, iters, 1.0000001, 1e-9, reuse);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = 2.0 * (reuse == 1 ? 2 : 8) * iters * 512.0 * sms * 2;
    printf("dfma+lds: %d FMA per LDS.64: %.2f TFLOP/s fp64, %.1f GB/s smem/SM-agg (%.3f ms)\n", reuse == 1 ? 2 : 8,
           fl / ms / 1e9, 8.0 * iters * 512.0 * sms * 2 / ms / 1e6, ms);
  }
  CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  {
    int iters = 40000;
    CK(cudaEventRecord(e0));
    lds_kernel<<<sms * 2, 512, 32768>>>(dout, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    double bytes = 8.0 * iters * 512.0 * sms * 2;
    printf("lds.64: %.1f GB/s aggregate = %.1f B/clk/SM at %d MHz nominal\n", bytes / ms / 1e6,
           bytes / ms / 1e3 / sms / (prop.clockRate / 1e3) , prop.clockRate / 1000);
  }
  // --- copy bandwidths
  size_t bytes = (size_t)4 << 30;
  char *a, *b;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes * 2));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes * 2));
  for (int bps : {4, 8, 16, 32}) {
    copy_kernel<<<sms * bps, 256>>>((const double2*)a, (double2*)b, bytes / 16);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 5; r++) copy_kernel<<<sms * bps, 256>>>((const double2*)a, (double2*)b, bytes / 16);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("copy ld/st.128 grid=%dxSM: %.1f GB/s (read+write)\n", bps, 2.0 * bytes * 5 / ms / 1e6);
  }
  {
    size_t nin = (size_t)1 << 30;  // 1 GiB in, 7 GiB out
    char* c;
    CK(cudaMalloc(&c, nin * 7));
    fanout_kernel<<<sms * 16, 256>>>((const double2*)a, (double2*)c, nin / 16, 7);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; r++) fanout_kernel<<<sms * 16, 256>>>((const double2*)a, (double2*)c, nin / 16, 7);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("fanout 1 read : 7 writes: %.1f GB/s (read+write)\n", 8.0 * nin * 3 / ms / 1e6);
    CK(cudaFree(c));
  }
  {
    constexpr int TB = 32768;
    CK(cudaFuncSetAttribute(bulk_copy_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TB));
    for (int bps : {2, 4, 6}) {
      bulk_copy_kernel<TB><<<sms * bps, 32, TB>>>(a, b, bytes);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int r = 0; r < 5; r++) bulk_copy_kernel<TB><<<sms * bps, 32, TB>>>(a, b, bytes);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("bulk copy (cp.async.bulk 32 KiB tiles, serial per CTA) %d CTA/SM: %.1f GB/s (read+write)\n", bps,
             2.0 * bytes * 5 / ms / 1e6);
    }
  }
  printf("done\n");
  return 0;
}
