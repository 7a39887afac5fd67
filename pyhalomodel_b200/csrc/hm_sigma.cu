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

// ---------------------------------------------------------------------------------------------------------------
// The kernel.  A CTA integrates 32*NRL radii of one sample: a LANE owns NRL consecutive radii, a WARP walks blocks of
// HM_SIG_BLK contiguous wavenumbers (blocks warp, warp + 8, ...).  Per block the warp first stages the per-wavenumber
// quantities in its slice of shared memory -- k_j, w_j = dlnk k_j^3 P(k_j), u_j = dlnk P(k_j)/k_j^3 and the second
// difference of the grid -- so every division and every load of k, P happens once per (block, 32*NRL radii) and the
// loops below read them with broadcast loads.  Each lane then classifies the block by the range of x = k R it spans
// for its radii and runs one of four loops:
//   skip      every x > 5e4: the terms are dropped (see hm_tophat);
//   series    every x < 2: Maclaurin series of T (13 terms reach 1e-21 at x = 2; 8 terms below x = 0.25, 4 below 0.03),
//             no trigonometry at all;
//   rotation  every x in [0.5, 5e4] and a smooth grid: (sin x, cos x) are seeded by ONE sincos per radius at the block
//             start and advanced by angle addition, sin(x+d) = sin x cos d + cos x sin d, with the increment
//             d_j = (k_{j+1} - k_j) R.  On a logarithmic grid d_j grows with x (up to 11 rad at x = 5e4), so
//             (sin d, cos d) are themselves advanced by angle addition with THEIR increment e_j = (k_{j+2} - 2 k_{j+1}
//             + k_j) R <= 2.6e-3, whose sine and cosine are three-term Taylor polynomials.  T^2 w is accumulated as
//             u_j (sin x - x cos x)^2 and scaled by 9/R^6 once per block.  19 FP64 operations per term instead of ~50
//             (sincos + division); rounding grows by ~1e-16 per step: < 1e-14 per 64-point block (host prototype
//             tools/proto_sigma_rotation.py: 8e-15 on sigma);
//   direct    anything else (blocks straddling a region boundary, coarse or irregular user grids): hm_tophat per term,
//             the arithmetic of round 1.
// The partial sums of a radius are combined in a fixed order that does not depend on NRL (bitwise reproducible).
// ---------------------------------------------------------------------------------------------------------------
#ifndef HM_SIG_BLK
#define HM_SIG_BLK 64
#endif
#define HM_SIG_WARPS 8
#define HM_SIG_THREADS (32 * HM_SIG_WARPS)

// T(x) = sum_n (-1)^n 3 x^(2n) / ((2n+3)(2n+1)!) truncated after NT terms: 13 terms reach 1e-21 at x = 2, 8 terms
// 1e-24 at x = 0.25, 4 terms 5e-19 at x = 0.03.
template <int NT>
__device__ __forceinline__ double hm_tophat_series_n(double x2) {
  constexpr double cf[13] = {1.,
                             -3. / (5. * 6.),
                             3. / (7. * 120.),
                             -3. / (9. * 5040.),
                             3. / (11. * 362880.),
                             -3. / (13. * 3.9916800e7),
                             3. / (15. * 6.227020800e9),
                             -3. / (17. * 1.307674368000e12),
                             3. / (19. * 3.55687428096000e14),
                             -3. / (21. * 1.21645100408832000e17),
                             3. / (23. * 5.1090942171709440000e19),
                             -3. / (25. * 2.585201673888498e22),
                             3. / (27. * 1.5511210043330986e25)};
  double t = cf[NT - 1];
#pragma unroll
  for (int n = NT - 2; n >= 0; --n) t = fma(t, x2, cf[n]);
  return t;
}

template <int NRL, int NT>
__device__ __forceinline__ void hm_sigma_series_block(const double* ks, const double* ws, int nj, const double (&r)[NRL],
                                                      double (&acc)[NRL]) {
  for (int j = 0; j < nj; ++j) {
    const double kj = ks[j], w = ws[j];
#pragma unroll
    for (int i = 0; i < NRL; ++i) {
      const double x = kj * r[i];
      const double T = hm_tophat_series_n<NT>(x * x);
      acc[i] = fma(w, T * T, acc[i]);
    }
  }
}

// Angle-addition loop of one block (see the kernel's comment).  SMALL_E: every |e| <= 1e-4, so sin e = e - e^3/6 and
// cos e = 1 - e^2/2 hold 1e-21.
template <int NRL, bool SMALL_E>
__device__ __forceinline__ void hm_sigma_rotation_block(const double* ks, const double* us, const double* dds, int nj,
                                                        const double (&r)[NRL], double (&acc)[NRL]) {
  double s[NRL], c[NRL], sd[NRL], cd[NRL], part[NRL];
  const double k0 = ks[0], dk0 = ks[1] - ks[0];
#pragma unroll
  for (int i = 0; i < NRL; ++i) {
    sincos(k0 * r[i], &s[i], &c[i]);                                     // the seeds: the only sincos of the block
    sincos(dk0 * r[i], &sd[i], &cd[i]);
    part[i] = 0.;
  }
  for (int j = 0; j < nj; ++j) {
    const double kj = ks[j], u = us[j], ddk = dds[j];
#pragma unroll
    for (int i = 0; i < NRL; ++i) {
      const double x = kj * r[i];
      const double t = fma(-x, c[i], s[i]);                              // sin x - x cos x
      part[i] = fma(u, t * t, part[i]);
      const double sn = fma(s[i], cd[i], c[i] * sd[i]);                  // x -> x + d
      const double cn = fma(c[i], cd[i], -(s[i] * sd[i]));
      s[i] = sn; c[i] = cn;
      const double e = ddk * r[i], e2 = e * e;                           // d -> d + e
      const double se = SMALL_E ? fma(e * e2, -1. / 6., e) : fma(e * e2, fma(e2, 1. / 120., -1. / 6.), e);
      const double ce = SMALL_E ? fma(e2, -0.5, 1.) : fma(e2, fma(e2, 1. / 24., -0.5), 1.);
      const double sdn = fma(sd[i], ce, cd[i] * se);
      const double cdn = fma(cd[i], ce, -(sd[i] * se));
      sd[i] = sdn; cd[i] = cdn;
    }
  }
#pragma unroll
  for (int i = 0; i < NRL; ++i) {
    const double r3 = r[i] * r[i] * r[i];
    acc[i] = fma(part[i], 9. / (r3 * r3), acc[i]);                       // T^2 w = (9/R^6) u (sin x - x cos x)^2
  }
}

template <int NRL>
__global__ void __launch_bounds__(HM_SIG_THREADS) hm_sigma_tophat_kernel(const double* __restrict__ kgrid,
                                                                         const double* __restrict__ Pk, double dlnk, int nkg,
                                                                         const double* __restrict__ R, int nR,
                                                                         double* __restrict__ out) {
  constexpr int BLK = HM_SIG_BLK;
  __shared__ double sk[HM_SIG_WARPS][BLK + 2], sw[HM_SIG_WARPS][BLK], su[HM_SIG_WARPS][BLK], sdd[HM_SIG_WARPS][BLK];
  __shared__ double sred[HM_SIG_WARPS][NRL][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int iR0 = blockIdx.x * (32 * NRL) + lane * NRL, gs = blockIdx.y;
  const double* P = Pk + (size_t)gs * nkg;
  double r[NRL], acc[NRL];
  double rmin = 1e300, rmax = 0.;
#pragma unroll
  for (int i = 0; i < NRL; ++i) {
    r[i] = R[(size_t)gs * nR + ((iR0 + i < nR) ? iR0 + i : nR - 1)];     // clamp: duplicates are never written
    acc[i] = 0.;
    rmin = fmin(rmin, fabs(r[i]));
    rmax = fmax(rmax, fabs(r[i]));
  }
  double* ks = sk[warp];
  double* ws = sw[warp];
  double* us = su[warp];
  double* dds = sdd[warp];
  const int nblocks = (nkg + BLK - 1) / BLK;
  for (int b = warp; b < nblocks; b += HM_SIG_WARPS) {
    const int j0 = b * BLK, nj = (j0 + BLK < nkg) ? BLK : nkg - j0;
    // ---- stage the block: k (two points beyond its end), w, u and the second differences ----
    __syncwarp();
    double emax = 0.;
    for (int j = lane; j < BLK + 2; j += 32) {
      const int jj = (j0 + j < nkg) ? j0 + j : nkg - 1;
      ks[j] = kgrid[jj];
    }
    __syncwarp();
    for (int j = lane; j < BLK; j += 32) {
      const double kj = ks[j];
      const double pj = (j < nj) ? P[j0 + j] : 0.;
      ws[j] = dlnk * kj * (pj * (kj * kj));                              // cosmology.py:86,103
      us[j] = dlnk * pj / (kj * kj * kj);                                // w_j / k_j^6
      const double dd = (ks[j + 2] - ks[j + 1]) - (ks[j + 1] - kj);
      dds[j] = dd;
      if (j < nj) emax = fmax(emax, fabs(dd));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, off));
    __syncwarp();
    const double kfirst = ks[0], klast = ks[nj - 1];
    const bool increasing = kfirst < klast || nj == 1;
    const double xlo = fmin(kfirst, klast) * rmin, xhi = fmax(kfirst, klast) * rmax;
    if (xlo > 5e4 && increasing && kfirst >= 0.) continue;               // ---- skip ----
    if (xhi < 2. && kfirst >= 0. && klast >= 0.) {                       // ---- series ----
      if (xhi < 0.03) hm_sigma_series_block<NRL, 4>(ks, ws, nj, r, acc);
      else if (xhi < 0.25) hm_sigma_series_block<NRL, 8>(ks, ws, nj, r, acc);
      else hm_sigma_series_block<NRL, 13>(ks, ws, nj, r, acc);
      continue;
    }
    // the rotation needs every x inside [0.5, 5e4], an increasing grid with two more points for the second difference,
    // and small second differences (the Taylor polynomials of sin e, cos e hold 1e-17 up to |e| = 3e-3)
    if (xlo >= 0.5 && xhi <= 5e4 && increasing && nj >= 8 && j0 + nj + 1 < nkg && emax * rmax <= 2.5e-3) {   // ---- rotation ----
      if (emax * rmax <= 1e-4) hm_sigma_rotation_block<NRL, true>(ks, us, dds, nj, r, acc);
      else hm_sigma_rotation_block<NRL, false>(ks, us, dds, nj, r, acc);
      continue;
    }
    for (int j = 0; j < nj; ++j) {                                       // ---- direct ----
      const double kj = ks[j], w = ws[j];
#pragma unroll
      for (int i = 0; i < NRL; ++i) {
        const double T = hm_tophat(kj * r[i]);
        acc[i] = fma(w, T * T, acc[i]);
      }
    }
  }
  // sum over the warps in order, one thread per radius
#pragma unroll
  for (int i = 0; i < NRL; ++i) sred[warp][i][lane] = acc[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NRL; ++i) {
      double tot = 0.;
#pragma unroll
      for (int w = 0; w < HM_SIG_WARPS; ++w) tot += sred[w][i][lane];
      if (iR0 + i < nR) out[(size_t)gs * nR + iR0 + i] = sqrt(tot / (2. * M_PI * M_PI));
    }
  }
}

extern "C" int hm_sigma_tophat(const double* kgrid_dev, const double* Pk_dev, double dlnk, int32_t nkg,
                               const double* R_dev, int32_t nR, int32_t G, double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(kgrid_dev && Pk_dev && R_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nkg > 0 && nR > 0 && G > 0 && G <= 65535, "bad sizes (G must be <= 65535 per call)");
  cudaStream_t st = (cudaStream_t)stream;
  // four radii per lane when that still leaves a few CTAs per SM, one per lane for small problems
  const long long ctas4 = (long long)((nR + 127) / 128) * G;
  if (ctas4 >= 2LL * hm_num_sms())
    hm_sigma_tophat_kernel<4><<<dim3((nR + 127) / 128, G), HM_SIG_THREADS, 0, st>>>(kgrid_dev, Pk_dev, dlnk, nkg, R_dev, nR, out_dev);
  else
    hm_sigma_tophat_kernel<1><<<dim3((nR + 31) / 32, G), HM_SIG_THREADS, 0, st>>>(kgrid_dev, Pk_dev, dlnk, nkg, R_dev, nR, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
