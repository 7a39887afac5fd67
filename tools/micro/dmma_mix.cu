// Microbenchmark: DMMA throughput with the Gram kernel's instruction mix (16-byte fragment loads from shared
// memory + DMUL scaling per 9 DMMAs), as a function of warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#define LD 72
__global__ void mix_k(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 2 * 64 * LD; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
  double acc[2][9][2];
  for (int r = 0; r < 2; ++r) for (int j = 0; j < 9; ++j) acc[r][j][0] = acc[r][j][1] = 0.;
  const double* Yb = sm + 2 * fk + fr * LD;
  for (int it = 0; it < iters; ++it) {
#pragma unroll 2
    for (int blk = 0; blk < 8; ++blk) {
      const double2 w = *reinterpret_cast<const double2*>(sm + 8 * blk + 2 * fk);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const double* Yr = Yb + r * 64 * LD + 8 * blk;
        double2 a1 = *reinterpret_cast<const double2*>(Yr), a2 = *reinterpret_cast<const double2*>(Yr + 56 * LD);
        a1.x *= w.x; a1.y *= w.y; a2.x *= w.x; a2.y *= w.y;
        double2 b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(Yr + (16 * j) * LD);
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[r][j][0], acc[r][j][1], a1.x, b[j].x);
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[r][4 + j][0], acc[r][4 + j][1], a2.x, b[j].x);
        dmma(acc[r][8][0], acc[r][8][1], a2.x, b[3].y);
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[r][j][0], acc[r][j][1], a1.y, b[j].y);
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[r][4 + j][0], acc[r][4 + j][1], a2.y, b[j].y);
        dmma(acc[r][8][0], acc[r][8][1], a2.y, b[3].x);
      }
    }
  }
  double s = 0;
  for (int r = 0; r < 2; ++r) for (int j = 0; j < 9; ++j) s += acc[r][j][0] + acc[r][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int smem = 2 * 64 * LD * 8;
  cudaFuncSetAttribute(mix_k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int iters = 2000; float ms;
  for (int warps = 4; warps <= 32; warps *= 2) {
    mix_k<<<148, warps * 32, smem>>>(out, 10);
    cudaEventRecord(e0); mix_k<<<148, warps * 32, smem>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fl = 148.0 * warps * iters * 8.0 * 2 * 18 * 512.0;
    printf("mix   warps/SM %2d: %.2f TFLOP/s  (%s)\n", warps, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
