// tools/microbench3.cu -- the inner-loop shape of the filter kernels: one sample x feeds 7 accumulators through 7 taps.
//   variant U: taps are uniform (kernel parameters -> UR operands)      DFMA acc, x.reuse, UR, acc
//   variant R: taps live in ordinary registers                          DFMA acc, x.reuse, Rtap, acc
// Measures what the fp64 pipe sustains for each operand form (B200).
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
struct Taps { double t[16]; };

__global__ void shape_u(double* out, int iters, const __grid_constant__ Taps t) {
  double acc[7], x = threadIdx.x * 1e-3;
  for (int q = 0; q < 7; q++) acc[q] = q;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int q = 0; q < 7; q++) acc[q] = fma(x, t.t[(q + r) & 15], acc[q]);
      x += 1e-9;
    }
  }
  double s = 0; for (int q = 0; q < 7; q++) s += acc[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void shape_r(double* out, int iters, const double* __restrict__ p) {
  double acc[7], tap[14], x = threadIdx.x * 1e-3;
  for (int q = 0; q < 7; q++) acc[q] = q;
  for (int q = 0; q < 14; q++) tap[q] = p[threadIdx.x + 32 * q];   // per-thread values: ordinary registers
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int q = 0; q < 7; q++) acc[q] = fma(x, tap[q + (r & 7) % 8 < 14 ? (q + r) % 14 : 0], acc[q]);
      x += 1e-9;
    }
  }
  double s = 0; for (int q = 0; q < 7; q++) s += acc[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double *dout, *dp; CK(cudaMalloc(&dout, 8 * 148 * 4 * 512)); CK(cudaMalloc(&dp, 8 * 1024));
  double hp[1024]; for (int i = 0; i < 1024; i++) hp[i] = 1e-7 * i; CK(cudaMemcpy(dp, hp, sizeof(hp), cudaMemcpyHostToDevice));
  Taps t; for (int i = 0; i < 16; i++) t.t[i] = 1e-8 * i;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); float ms; const int iters = 4000;
  for (int rep = 0; rep < 2; rep++) {
    for (int threads : {128, 256}) {
      CK(cudaEventRecord(e0)); shape_u<<<sms * 4, threads>>>(dout, iters, t); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep) printf("taps uniform (UR)  %3d thr x4 CTA/SM: %.2f TFLOP/s\n", threads, 2.0 * 56 * iters * threads * sms * 4 / ms / 1e9);
      CK(cudaEventRecord(e0)); shape_r<<<sms * 4, threads>>>(dout, iters, dp); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep) printf("taps in registers  %3d thr x4 CTA/SM: %.2f TFLOP/s\n", threads, 2.0 * 56 * iters * threads * sms * 4 / ms / 1e9);
    }
  }
  return 0;
}
