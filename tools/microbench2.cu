// tools/microbench2.cu -- DFMA operand-form throughput on B200: (reg, uniform/const, reg) vs (reg, reg, reg),
// and DFMA mixed with one indexed constant load (LDC) per N DFMA -- the two tap delivery schemes of the long-filter kernels.
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Taps { double t[64]; };

__global__ void dfma_ur(double* out, int iters, double a, double b) {
  double x[8];
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  for (int i = 0; i < iters; i++)
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = fma(x[k], a, b);
  double s = 0; for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_rrr(double* out, int iters, const double* __restrict__ p) {
  double x[8];
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  const double a = p[threadIdx.x], b = p[threadIdx.x + 32];   // per-thread values: ordinary registers
  const double c = p[threadIdx.x + 64], d = p[threadIdx.x + 96];
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) { x[k] = fma(x[k], a, b); x[k + 1] = fma(x[k + 1], c, d); }
  }
  double s = 0; for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int PER>
__global__ void dfma_ldc(double* out, int iters, const __grid_constant__ Taps t) {
  double x[8];
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  const double* ct = t.t + z;
  for (int i = 0; i < iters; i++) {
    const double a = ct[i & 63];           // one LDC.64 (register-indexed) per 8 / PER DFMA groups
#pragma unroll
    for (int r = 0; r < PER; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = fma(x[k], a, 1e-9);
  }
  double s = 0; for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double *dout, *dp; CK(cudaMalloc(&dout, 8 * 148 * 4 * 512)); CK(cudaMalloc(&dp, 8 * 256));
  double hp[256]; for (int i = 0; i < 256; i++) hp[i] = 1.0 + 1e-7 * i; CK(cudaMemcpy(dp, hp, sizeof(hp), cudaMemcpyHostToDevice));
  Taps t; for (int i = 0; i < 64; i++) t.t[i] = 1.0 + 1e-8 * i;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); float ms; const int iters = 20000;
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(e0)); dfma_ur<<<sms * 4, 256>>>(dout, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) printf("DFMA reg,uniform,uniform : %.2f TFLOP/s\n", 2.0 * 8 * iters * 256.0 * sms * 4 / ms / 1e9);
    CK(cudaEventRecord(e0)); dfma_rrr<<<sms * 4, 256>>>(dout, iters, dp); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) printf("DFMA reg,reg,reg         : %.2f TFLOP/s\n", 2.0 * 8 * iters * 256.0 * sms * 4 / ms / 1e9);
    CK(cudaEventRecord(e0)); dfma_ldc<1><<<sms * 4, 256>>>(dout, iters, t); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) printf("DFMA + 1 LDC.64 per 8    : %.2f TFLOP/s\n", 2.0 * 8 * iters * 256.0 * sms * 4 / ms / 1e9);
    CK(cudaEventRecord(e0)); dfma_ldc<2><<<sms * 4, 256>>>(dout, iters, t); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep) printf("DFMA + 1 LDC.64 per 16   : %.2f TFLOP/s\n", 2.0 * 16 * iters * 256.0 * sms * 4 / ms / 1e9);
  }
  return 0;
}
