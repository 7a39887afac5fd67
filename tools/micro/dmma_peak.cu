// Microbenchmark: FP64 DFMA vs DMMA (mma.sync m8n8k4 f64) issue throughput on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_k(double* out, int iters) {
  double a0 = threadIdx.x, a1 = 1.1, a2 = 2.2, a3 = 3.3, a4 = 4.4, a5 = 5.5, a6 = 6.6, a7 = 7.7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma_k(double* out, int iters) {
  double c[8][2];
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma(c[j][0], c[j][1], a, b);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 20000; float ms;
  for (int warps = 4; warps <= 32; warps *= 2) {
    dfma_k<<<148, warps * 32>>>(out, 100);
    cudaEventRecord(e0); dfma_k<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 148.0 * warps * 32 * iters * 8 * 2;
    printf("DFMA  warps/SM %2d: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
    dmma_k<<<148, warps * 32>>>(out, 100);
    cudaEventRecord(e0); dmma_k<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 148.0 * warps * iters * 8 * 512.0;
    printf("DMMA  warps/SM %2d: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
