// hm_sigma.cu -- top-hat sigma(R) by the reference's brute log-k rectangle rule.
//   cosmology._sigmaR_brute_log (cosmology.py:89-106), _sigmaR_integrand (:77-86), _Tophat_k (:41-48)
//   sigma(R) = sqrt( sum_j dlnk * k_j * (P(k_j) * k_j^2 * T(k_j R)^2) / (2 pi^2) )
// The integration grid k_j and P(k_j) come from the caller (built host-side exactly as utility.py:43 does, so
// the abscissae are bit-identical).  One CTA per radius; FP64-ALU (sincos) bound; k/P stream from L2.
#include "hm_common.cuh"

__device__ __forceinline__ double hm_tophat(double x) {
  if (fabs(x) < 1e-5) return 1. - x * x / 10.;          // cosmology.py:15,47
  double s, c;
  sincos(x, &s, &c);
  return (3. / (x * x * x)) * (s - x * c);
}

__global__ void __launch_bounds__(256) hm_sigma_tophat_kernel(const double* __restrict__ kgrid,
                                                              const double* __restrict__ Pk, double dlnk, int nkg,
                                                              const double* __restrict__ R, int nR,
                                                              double* __restrict__ out) {
  __shared__ double scratch[32];
  const int iR = blockIdx.x, gs = blockIdx.y;
  const double r = R[(size_t)gs * nR + iR];
  const double* P = Pk + (size_t)gs * nkg;
  double acc = 0.;
  for (int j = threadIdx.x; j < nkg; j += blockDim.x) {
    const double kj = kgrid[j];
    const double T = hm_tophat(kj * r);
    acc += dlnk * kj * (P[j] * (kj * kj) * (T * T));     // cosmology.py:86,103
  }
  const double tot = hm_block_sum(acc, scratch);
  if (threadIdx.x == 0) out[(size_t)gs * nR + iR] = sqrt(tot / (2. * M_PI * M_PI));
}

extern "C" int hm_sigma_tophat(const double* kgrid_dev, const double* Pk_dev, double dlnk, int32_t nkg,
                               const double* R_dev, int32_t nR, int32_t G, double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(kgrid_dev && Pk_dev && R_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nkg > 0 && nR > 0 && G > 0 && G <= 65535, "bad sizes (G must be <= 65535 per call)");
  dim3 grid(nR, G);
  hm_sigma_tophat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kgrid_dev, Pk_dev, dlnk, nkg, R_dev, nR, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
