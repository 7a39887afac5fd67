// hm_massfn.cu -- element-wise mass function / bias and halo averages.
//   hm_massfn_nu : model._mass_function_nu / _linear_bias_nu  (pyhalomodel.py:217-284)
//   hm_average   : model.average                              (pyhalomodel.py:335-359)
#include "hm_common.cuh"

__global__ void hm_massfn_nu_kernel(const hm_model_desc d, const double* __restrict__ nu, long long n,
                                    double* __restrict__ g, double* __restrict__ b) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = nu[i];
    if (g) g[i] = hm_g_nu(d, x);
    if (b) b[i] = hm_b_nu(d, x);
  }
}

extern "C" int hm_massfn_nu(const hm_model_desc* model_host, const double* nu_dev, int64_t n, double* g_dev,
                            double* b_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(model_host && nu_dev && n >= 0, "NULL pointer or negative length");
  HM_REQUIRE(model_host->family >= HM_PS74 && model_host->family <= HM_TINKER10_PBS, "unknown family");
  if (n == 0) return HM_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)hm_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  hm_massfn_nu_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*model_host, nu_dev, n, g_dev, b_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

// <f_j> = rhom * sum_m tw[m] (f_j[m]/M[m]) g(nu[m]), tw = trapezoid-in-nu weights.  One CTA, fixed-order reduction.
__global__ void __launch_bounds__(256) hm_average_kernel(const hm_model_desc d, const double* __restrict__ M,
                                                         const double* __restrict__ sigma, int nM,
                                                         const double* __restrict__ f, int nf,
                                                         double* __restrict__ out) {
  __shared__ double scratch[32];
  const int j = blockIdx.x;
  double acc = 0.;
  for (int m = threadIdx.x; m < nM; m += blockDim.x) {
    const double nu = d.dc / sigma[m];
    const double nu_hi = (m + 1 < nM) ? d.dc / sigma[m + 1] : nu;
    const double nu_lo = (m > 0) ? d.dc / sigma[m - 1] : nu;
    acc += (0.5 * (nu_hi - nu_lo)) * ((f[(size_t)j * nM + m] / M[m]) * hm_g_nu(d, nu));
  }
  const double tot = hm_block_sum(acc, scratch);
  if (threadIdx.x == 0) out[j] = tot * d.rhom;
}

extern "C" int hm_average(const hm_model_desc* model_host, const double* M_dev, const double* sigma_dev,
                          int32_t nM, const double* f_dev, int32_t nf, double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(model_host && M_dev && sigma_dev && f_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nM >= 2 && nf >= 1, "need nM >= 2 and nf >= 1");
  hm_average_kernel<<<nf, 256, 0, (cudaStream_t)stream>>>(*model_host, M_dev, sigma_dev, nM, f_dev, nf, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

// Batched variant for sweeps: out[b] = <f_{b/F}> under model b.  One CTA per model.
__global__ void __launch_bounds__(256) hm_average_batch_kernel(const hm_model_desc* __restrict__ models,
                                                               const double* __restrict__ M,
                                                               const double* __restrict__ sigma, int nM, int F,
                                                               const double* __restrict__ f, double* __restrict__ out) {
  __shared__ double scratch[32];
  const int b = blockIdx.x, s = b / F;
  const hm_model_desc d = models[b];
  const double* sig = sigma + (size_t)s * nM;
  const double* fs = f + (size_t)s * nM;
  double acc = 0.;
  for (int m = threadIdx.x; m < nM; m += blockDim.x) {
    const double nu = d.dc / sig[m];
    const double nu_hi = (m + 1 < nM) ? d.dc / sig[m + 1] : nu;
    const double nu_lo = (m > 0) ? d.dc / sig[m - 1] : nu;
    acc += (0.5 * (nu_hi - nu_lo)) * ((fs[m] / M[m]) * hm_g_nu(d, nu));
  }
  const double tot = hm_block_sum(acc, scratch);
  if (threadIdx.x == 0) out[b] = tot * d.rhom;
}

extern "C" int hm_average_batch(const hm_model_desc* models_dev, const double* M_dev, const double* sigma_dev,
                                int32_t nM, int32_t n_samples, int32_t models_per_sample, const double* f_dev,
                                double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(models_dev && M_dev && sigma_dev && f_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nM >= 2 && n_samples >= 1 && models_per_sample >= 1, "bad sizes");
  hm_average_batch_kernel<<<n_samples * models_per_sample, 256, 0, (cudaStream_t)stream>>>(
      models_dev, M_dev, sigma_dev, nM, models_per_sample, f_dev, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

// Generic weighted row sums: out[j][k] = sum_m w[j][m] * grid[k][m]  (the building block of the halo-model mass
// integrals once the trapezoid rule and the mass function are folded into w; used by cross_halo_spectrum,
// pyhalomodel.py:573-704).  One CTA per grid row, fixed-order reduction.
__global__ void __launch_bounds__(256) hm_rowsum_kernel(const double* __restrict__ grid, const double* __restrict__ w,
                                                        int nk, int nM, int nw, double* __restrict__ out) {
  __shared__ double scratch[32];
  const int ik = blockIdx.x;
  const double* row = grid + (size_t)ik * nM;
  for (int j = 0; j < nw; ++j) {
    double acc = 0.;
    for (int m = threadIdx.x; m < nM; m += blockDim.x) acc = fma(w[(size_t)j * nM + m], row[m], acc);
    const double tot = hm_block_sum(acc, scratch);
    if (threadIdx.x == 0) out[(size_t)j * nk + ik] = tot;
  }
}

extern "C" int hm_weighted_rowsum(const double* grid_dev, const double* w_dev, int32_t nk, int32_t nM, int32_t nw,
                                  double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(grid_dev && w_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nk > 0 && nM > 0 && nw > 0, "bad sizes");
  hm_rowsum_kernel<<<nk, 256, 0, (cudaStream_t)stream>>>(grid_dev, w_dev, nk, nM, nw, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
