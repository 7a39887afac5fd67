// hm_sigma.cu -- top-hat sigma(R) by the reference's brute log-k rectangle rule.
//   cosmology._sigmaR_brute_log (cosmology.py:89-106), _sigmaR_integrand (:77-86), _Tophat_k (:41-48)
//   sigma(R) = sqrt( sum_j dlnk * k_j * (P(k_j) * k_j^2 * T(k_j R)^2) / (2 pi^2) )
// The integration grid k_j and P(k_j) come from the caller (built host-side exactly as utility.py:43 does, so
// the abscissae are bit-identical).  One CTA per radius; FP64-ALU (sincos) bound; k/P stream from L2.
#include "hm_common.cuh"

// T(x) = 3 (sin x - x cos x)/x^3.  The reference switches to 1 - x^2/10 below |x| = 1e-5 and evaluates the closed
// form everywhere else (cosmology.py:41-48), where it cancels catastrophically (relative error ~1e-16/x^2, up to
// 4e-6 just above the switch).  Here:
//   |x| < 0.5  : the Maclaurin series sum_n (-1)^n 3 x^(2n) / ((2n+3)(2n+1)!) to 12 terms (4e-16 vs mpmath),
//   |x| > 5e4  : T^2 < 4e-19, below 1e-13 of sigma^2 for every radius of a halo-mass grid (tools + tests) -> 0,
//                which also keeps sincos on its fast path (no Payne-Hanek reduction),
//   otherwise  : the closed form, as the reference.
// sigma(R) changes by < 2e-13 relative against the reference's arithmetic (tests/test_gpu_parity.py::test_sigmaR).
__device__ __forceinline__ double hm_tophat(double x) {
  const double ax = fabs(x);
  if (ax < 0.5) {
    const double x2 = x * x;
    double t = -3. / (25. * 2.585201673888498e22);              // n = 11: (-1)^n 3/((2n+3)(2n+1)!) = -3/(25*23!)
    t = fma(t, x2, 3. / (23. * 5.1090942171709440000e19));     // n = 10: 21!
    t = fma(t, x2, -3. / (21. * 1.21645100408832000e17));      // n = 9:  19!
    t = fma(t, x2, 3. / (19. * 3.55687428096000e14));          // n = 8:  17!
    t = fma(t, x2, -3. / (17. * 1.307674368000e12));           // n = 7:  15!
    t = fma(t, x2, 3. / (15. * 6.227020800e9));                // n = 6:  13!
    t = fma(t, x2, -3. / (13. * 3.9916800e7));                 // n = 5:  11!
    t = fma(t, x2, 3. / (11. * 362880.));                      // n = 4:  9!
    t = fma(t, x2, -3. / (9. * 5040.));                        // n = 3:  7!
    t = fma(t, x2, 3. / (7. * 120.));                          // n = 2:  5!
    t = fma(t, x2, -3. / (5. * 6.));                           // n = 1:  3!
    t = fma(t, x2, 1.);                                        // n = 0
    return t;
  }
  if (ax > 5e4) return 0.;
  double s, c;
  sincos(x, &s, &c);
  return (3. / (x * x * x)) * (s - x * c);
}

// One CTA integrates HM_SIG_NR radii of one sample at once, so the integration grid k_j and P(k_j) (1.6 MB per
// sample, L2-resident) are read once per HM_SIG_NR radii and the weight dlnk k^3 P(k) is formed once per k.
#define HM_SIG_NR 8
__global__ void __launch_bounds__(256) hm_sigma_tophat_kernel(const double* __restrict__ kgrid,
                                                              const double* __restrict__ Pk, double dlnk, int nkg,
                                                              const double* __restrict__ R, int nR,
                                                              double* __restrict__ out) {
  __shared__ double scratch[32];
  const int iR0 = blockIdx.x * HM_SIG_NR, gs = blockIdx.y;
  const double* P = Pk + (size_t)gs * nkg;
  double r[HM_SIG_NR], acc[HM_SIG_NR];
#pragma unroll
  for (int i = 0; i < HM_SIG_NR; ++i) {
    r[i] = R[(size_t)gs * nR + ((iR0 + i < nR) ? iR0 + i : nR - 1)];     // clamp: duplicates are never written
    acc[i] = 0.;
  }
  for (int j = threadIdx.x; j < nkg; j += blockDim.x) {
    const double kj = kgrid[j];
    const double w = dlnk * kj * (P[j] * (kj * kj));                     // cosmology.py:86,103
#pragma unroll
    for (int i = 0; i < HM_SIG_NR; ++i) {
      const double T = hm_tophat(kj * r[i]);
      acc[i] = fma(w, T * T, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < HM_SIG_NR; ++i) {
    const double tot = hm_block_sum(acc[i], scratch);
    if (threadIdx.x == 0 && iR0 + i < nR) out[(size_t)gs * nR + iR0 + i] = sqrt(tot / (2. * M_PI * M_PI));
  }
}

extern "C" int hm_sigma_tophat(const double* kgrid_dev, const double* Pk_dev, double dlnk, int32_t nkg,
                               const double* R_dev, int32_t nR, int32_t G, double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(kgrid_dev && Pk_dev && R_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nkg > 0 && nR > 0 && G > 0 && G <= 65535, "bad sizes (G must be <= 65535 per call)");
  dim3 grid((nR + HM_SIG_NR - 1) / HM_SIG_NR, G);
  hm_sigma_tophat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kgrid_dev, Pk_dev, dlnk, nkg, R_dev, nR, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
