// EXPERIMENT (not part of libhalob200.so): top-hat sigma(R) with sin/cos of k*R advanced by angle addition.
//
// The shipped kernel (csrc/hm_sigma.cu) calls sincos and divides once per (k_j, R_i) term and is FP64-ALU bound:
// 93 % of the end-to-end pipeline of BASELINE config 5.  Here a thread owns a CONTIGUOUS block of k points for 8
// radii, seeds (sin, cos) with sincos at the start of every block of 64 points and advances them with
//     sin(x+d) = sin x cos d + cos x sin d,   cos(x+d) = cos x cos d - sin x sin d,   d = (k_{j+1} - k_j) R,
// sin d and cos d from their Taylor series (d = 2.3e-4 x on the 1e5-point logarithmic grid), falling back to sincos
// where x > 300 (d not small).  1/x^3 is formed from a per-k and a per-R factor, so the loop has no division.
// The numerics were checked on the host first (tools/proto_sigma_rotation.py: 8e-15 against the direct sum).
// This program times both kernels on synthetic inputs and prints the largest relative difference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/experimental/sigma_rot tools/experimental/sigma_rot.cu
#include <cmath>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define NR 8          // radii per thread
#define BLK 64        // k points between sincos seeds
#define XROT 300.     // above this the increment is not small: direct sincos

__device__ __forceinline__ double tophat_series(double x) {   // |x| < 0.5, as csrc/hm_sigma.cu
  const double x2 = x * x;
  double t = -3. / (25. * 2.585201673888498e22);
  t = fma(t, x2, 3. / (23. * 5.1090942171709440000e19));
  t = fma(t, x2, -3. / (21. * 1.21645100408832000e17));
  t = fma(t, x2, 3. / (19. * 3.55687428096000e14));
  t = fma(t, x2, -3. / (17. * 1.307674368000e12));
  t = fma(t, x2, 3. / (15. * 6.227020800e9));
  t = fma(t, x2, -3. / (13. * 3.9916800e7));
  t = fma(t, x2, 3. / (11. * 362880.));
  t = fma(t, x2, -3. / (9. * 5040.));
  t = fma(t, x2, 3. / (7. * 120.));
  t = fma(t, x2, -3. / (5. * 6.));
  return fma(t, x2, 1.);
}

__device__ __forceinline__ double tophat_direct(double x) {
  const double ax = fabs(x);
  if (ax < 0.5) return tophat_series(x);
  if (ax > 5e4) return 0.;
  double s, c;
  sincos(x, &s, &c);
  return (3. / (x * x * x)) * (s - x * c);
}

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
  return t;
}

// the shipped scheme: threads stride over k, sincos + division per term
__global__ void __launch_bounds__(256) sigma_direct_kernel(const double* __restrict__ kgrid, const double* __restrict__ Pk,
                                                           double dlnk, int nkg, const double* __restrict__ R, int nR,
                                                           double* __restrict__ out) {
  __shared__ double scratch[32];
  const int iR0 = blockIdx.x * NR, gs = blockIdx.y;
  const double* P = Pk + (size_t)gs * nkg;
  double r[NR], acc[NR];
  for (int i = 0; i < NR; ++i) { r[i] = R[(size_t)gs * nR + ((iR0 + i < nR) ? iR0 + i : nR - 1)]; acc[i] = 0.; }
  for (int j = threadIdx.x; j < nkg; j += blockDim.x) {
    const double kj = kgrid[j];
    const double w = dlnk * kj * (P[j] * (kj * kj));
#pragma unroll
    for (int i = 0; i < NR; ++i) { const double T = tophat_direct(kj * r[i]); acc[i] = fma(w, T * T, acc[i]); }
  }
  for (int i = 0; i < NR; ++i) {
    const double tot = block_sum(acc[i], scratch);
    if (threadIdx.x == 0 && iR0 + i < nR) out[(size_t)gs * nR + iR0 + i] = sqrt(tot / (2. * M_PI * M_PI));
  }
}

// angle addition: thread t owns the k blocks t, t + blockDim, ... of BLK contiguous points
__global__ void __launch_bounds__(256) sigma_rot_kernel(const double* __restrict__ kgrid, const double* __restrict__ Pk,
                                                        double dlnk, int nkg, const double* __restrict__ R, int nR,
                                                        double* __restrict__ out) {
  __shared__ double scratch[32];
  const int iR0 = blockIdx.x * NR, gs = blockIdx.y;
  const double* P = Pk + (size_t)gs * nkg;
  double r[NR], ir3[NR], acc[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    r[i] = R[(size_t)gs * nR + ((iR0 + i < nR) ? iR0 + i : nR - 1)];
    ir3[i] = 3. / (r[i] * r[i] * r[i]);
    acc[i] = 0.;
  }
  const int nblocks = (nkg + BLK - 1) / BLK;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    const int j0 = b * BLK, j1 = (j0 + BLK < nkg) ? j0 + BLK : nkg;
    double s[NR], c[NR];
    double kj = kgrid[j0];
#pragma unroll
    for (int i = 0; i < NR; ++i) sincos(kj * r[i], &s[i], &c[i]);       // seed (cheap: once per BLK points)
    for (int j = j0; j < j1; ++j) {
      const double pj = P[j];
      const double w = dlnk * kj * (pj * (kj * kj));
      const double ik3 = 1. / (kj * kj * kj);                            // one division per k, shared by the radii
      const double kn = (j + 1 < j1) ? kgrid[j + 1] : kj;
      const double dk = kn - kj;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const double x = kj * r[i];
        double T;
        if (x < 0.5) T = tophat_series(x);
        else if (x > 5e4) T = 0.;
        else T = (ir3[i] * ik3) * (s[i] - x * c[i]);
        acc[i] = fma(w, T * T, acc[i]);
        // advance (sin, cos) to the next point
        if (x > XROT) {
          if (x <= 5e4) sincos(kn * r[i], &s[i], &c[i]);
        } else {
          const double d = dk * r[i], d2 = d * d;
          double sd = fma(d2, 1. / 362880., -1. / 5040.);                // sin d = d (1 - d2/6 + d2^2/120 - ...)
          sd = fma(sd, d2, 1. / 120.);
          sd = fma(sd, d2, -1. / 6.);
          sd = fma(sd * d2, d, d);
          double cd = fma(d2, -1. / 3628800., 1. / 40320.);              // cos d = 1 - d2/2 + d2^2/24 - ...
          cd = fma(cd, d2, -1. / 720.);
          cd = fma(cd, d2, 1. / 24.);
          cd = fma(cd, d2, -0.5);
          cd = fma(cd, d2, 1.);
          const double sn = fma(s[i], cd, c[i] * sd), cn = fma(c[i], cd, -(s[i] * sd));
          s[i] = sn; c[i] = cn;
        }
      }
      kj = kn;
    }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const double tot = block_sum(acc[i], scratch);
    if (threadIdx.x == 0 && iR0 + i < nR) out[(size_t)gs * nR + iR0 + i] = sqrt(tot / (2. * M_PI * M_PI));
  }
}

int main() {
  const int nkg = 100000, nR = 1025, G = 16;
  std::vector<double> k(nkg), P((size_t)G * nkg), R((size_t)G * nR);
  const double dlnk = log(1e5 / 1e-5) / (nkg - 1);
  for (int j = 0; j < nkg; ++j) k[j] = 1e-5 * exp(dlnk * j);
  for (int g = 0; g < G; ++g) {
    for (int j = 0; j < nkg; ++j) {       // BBKS-like shape: k^ns T(k)^2
      const double q = k[j] / (0.2 + 0.01 * g);
      const double T = log(1. + 2.34 * q) / (2.34 * q) * pow(1. + 3.89 * q + pow(16.1 * q, 2) + pow(5.46 * q, 3) + pow(6.71 * q, 4), -0.25);
      P[(size_t)g * nkg + j] = 2e6 * pow(k[j], 0.96) * T * T;
    }
    for (int i = 0; i < nR; ++i) R[(size_t)g * nR + i] = 0.05 * pow(1e17 / 1e9, (double)i / (3. * (nR - 1))) * (1. + 0.01 * g);
  }
  double *dk, *dP, *dR, *o1, *o2;
  cudaMalloc(&dk, nkg * 8); cudaMalloc(&dP, (size_t)G * nkg * 8); cudaMalloc(&dR, (size_t)G * nR * 8);
  cudaMalloc(&o1, (size_t)G * nR * 8); cudaMalloc(&o2, (size_t)G * nR * 8);
  cudaMemcpy(dk, k.data(), nkg * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dP, P.data(), (size_t)G * nkg * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dR, R.data(), (size_t)G * nR * 8, cudaMemcpyHostToDevice);
  dim3 grid((nR + NR - 1) / NR, G);
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  sigma_direct_kernel<<<grid, 256>>>(dk, dP, dlnk, nkg, dR, nR, o1);
  sigma_rot_kernel<<<grid, 256>>>(dk, dP, dlnk, nkg, dR, nR, o2);
  cudaEventRecord(e0);
  sigma_direct_kernel<<<grid, 256>>>(dk, dP, dlnk, nkg, dR, nR, o1);
  cudaEventRecord(e1);
  sigma_rot_kernel<<<grid, 256>>>(dk, dP, dlnk, nkg, dR, nR, o2);
  cudaEventRecord(e2);
  cudaError_t err = cudaDeviceSynchronize();
  float t1 = 0.f, t2 = 0.f;
  cudaEventElapsedTime(&t1, e0, e1); cudaEventElapsedTime(&t2, e1, e2);
  std::vector<double> a((size_t)G * nR), b((size_t)G * nR);
  cudaMemcpy(a.data(), o1, a.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), o2, b.size() * 8, cudaMemcpyDeviceToHost);
  double worst = 0.;
  for (size_t i = 0; i < a.size(); ++i) worst = fmax(worst, fabs(b[i] / a[i] - 1.));
  const double terms = (double)G * nR * nkg;
  printf("direct: %.3f ms (%.2e terms/s)   rotation: %.3f ms (%.2e terms/s)   speed-up %.2fx   max |rot/direct - 1| = %.2e   %s\n",
         t1, terms / (t1 * 1e-3), t2, terms / (t2 * 1e-3), t1 / t2, worst, cudaGetErrorString(err));
  return err != cudaSuccess || !(worst < 1e-12);
}
