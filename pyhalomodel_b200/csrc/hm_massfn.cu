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

// ---------------------------------------------------------------------------------------------------------------
// model.multiplicity_function / mass_function (pyhalomodel.py:286-315): F = g(nu) dnu/dlnM with
// dnu/dlnM = -(nu/6) 2 dln(sigma)/dln(R) and the derivative of the quadratic through the three samples around each
// point (utility.derivative_from_samples, utility.py:15-35; the end points use the first / last three samples),
// in closed Lagrange form.  One thread per mass.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hm_multiplicity_kernel(const hm_model_desc d, const double* __restrict__ M,
                                                              const double* __restrict__ sigma, int nM,
                                                              double* __restrict__ F, double* __restrict__ n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nM) return;
  const int j = (i < 1) ? 1 : ((i > nM - 2) ? nM - 2 : i);
  const double c = 3. / (4. * M_PI * d.rhom);                      // R = cbrt(3M/(4 pi rhom)), cosmology.py:148-155
  const double x0 = log(cbrt(c * M[j - 1])), x1 = log(cbrt(c * M[j])), x2 = log(cbrt(c * M[j + 1]));
  const double f0 = log(sigma[j - 1]), f1 = log(sigma[j]), f2 = log(sigma[j + 1]);
  const double x = (i == j) ? x1 : ((i < j) ? x0 : x2);
  const double deriv = f0 * (2. * x - x1 - x2) / ((x0 - x1) * (x0 - x2)) + f1 * (2. * x - x0 - x2) / ((x1 - x0) * (x1 - x2)) +
                       f2 * (2. * x - x0 - x1) / ((x2 - x0) * (x2 - x1));
  const double nu = d.dc / sigma[i];
  const double Fi = hm_g_nu(d, nu) * (-(nu / 6.) * (2. * deriv));
  if (F) F[i] = Fi;
  if (n) n[i] = Fi * d.rhom / (M[i] * M[i]);
}

extern "C" int hm_multiplicity_function(const hm_model_desc* model_host, const double* M_dev, const double* sigma_dev,
                                        int32_t nM, double* F_dev, double* n_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(model_host && M_dev && sigma_dev && (F_dev || n_dev), "NULL pointer");
  HM_REQUIRE(nM >= 3, "need at least three masses");
  hm_multiplicity_kernel<<<(nM + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*model_host, M_dev, sigma_dev, nM, F_dev, n_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// model.cross_halo_spectrum (pyhalomodel.py:573-704): every tracer crossed with haloes of ONE mass M_h.
//   stage 1 (one CTA): nu, g, b, the trapezoid weights w2 = tw b g / M, A = 1 - trapz(g b) (a trapezoid here, not the
//                      integral to infinity: :624-625), nu(M_h) = sum L nu and b(M_h) (:631-639; L are the weights of
//                      the cubic interpolation in ln M, linear in the samples, supplied by the caller);
//   stage 2 (CTA per (k, tracer)): I_u(k), W_u(k, M_h) = sum L W_u(k, .), the beta_NL integral (:691-704) and the
//                      three spectra (:641-677).
// Workspace: (nM + 4) doubles.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hm_cross_prep_kernel(const hm_model_desc d, const double* __restrict__ M,
                                                            const double* __restrict__ sigma, const double* __restrict__ L,
                                                            int nM, double* __restrict__ w2, double* __restrict__ scal) {
  __shared__ double scratch[32];
  double accA = 0., accnu = 0.;
  for (int m = threadIdx.x; m < nM; m += blockDim.x) {
    const double nu = d.dc / sigma[m];
    const double nu_hi = (m + 1 < nM) ? d.dc / sigma[m + 1] : nu;
    const double nu_lo = (m > 0) ? d.dc / sigma[m - 1] : nu;
    const double tw = 0.5 * (nu_hi - nu_lo);
    const double bg = hm_b_nu(d, nu) * hm_g_nu(d, nu);
    w2[m] = tw * (bg / M[m]);
    accA += tw * bg;
    accnu += L[m] * nu;
  }
  const double totA = hm_block_sum(accA, scratch);
  const double nuh = hm_block_sum(accnu, scratch);
  if (threadIdx.x == 0) {
    scal[0] = 1. - totA;
    scal[1] = nuh;
    scal[2] = hm_b_nu(d, nuh);
  }
}

struct hm_cross_args {
  int nk, nM, simple_twohalo;
  double k_trunc, rhom;
  const double* k;
  const double* Pk_lin;
  const double* M;
  const double* L;
  const double* beta;                 // [nk][nM] or NULL
  const double* w2;
  const double* scal;
  const double* gridW[HM_MAX_TRACERS];
  const double* amplitude[HM_MAX_TRACERS];
  double norm[HM_MAX_TRACERS];
  int mode[HM_MAX_TRACERS], mass[HM_MAX_TRACERS];
  double* out;                        // [P][3][nk]
};

__global__ void __launch_bounds__(128) hm_cross_kernel(const hm_cross_args a) {
  __shared__ double scratch[32];
  const int ik = blockIdx.x, p = blockIdx.y;
  const double* row = a.gridW[p] + (size_t)ik * a.nM;
  const double* brow = a.beta ? a.beta + (size_t)ik * a.nM : nullptr;
  const bool uk = a.mode[p] == HM_GRID_UK;
  const double A = a.scal[0], bh = a.scal[2], M0 = a.M[0];
  double aI = 0., aW = 0., aS = 0., aB = 0.;
  for (int m = threadIdx.x; m < a.nM; m += blockDim.x) {
    const double cW = uk ? a.amplitude[p][m] / a.norm[p] : 1.;         // W = cW * grid (Wk = amplitude Uk / norm)
    const double G = row[m];
    double e = a.w2[m] * cW;
    if (m == 0 && a.mass[p]) e += (A / M0) * cW;                      // A W(k, M0)/M0 (:541-542)
    aI = fma(e, G, aI);
    aW = fma(a.L[m] * cW, G, aW);
    if (a.simple_twohalo) aS = fma(a.w2[m], a.amplitude[p][m] / a.norm[p], aS);
    if (brow) aB = fma(brow[m] * (G * cW), a.w2[m], aB);
  }
  const double I = hm_block_sum(aI, scratch);
  const double W = hm_block_sum(aW, scratch);
  const double Ssum = a.simple_twohalo ? hm_block_sum(aS, scratch) : 0.;
  const double Bsum = brow ? hm_block_sum(aB, scratch) : 0.;
  if (threadIdx.x == 0) {
    double Iu = I * a.rhom;
    if (a.simple_twohalo) {                                            // (:648-650)
      const double low = a.mass[p] ? A * (a.amplitude[p][0] / a.norm[p]) / M0 : 0.;
      Iu = a.rhom * (Ssum + low);
    }
    double I_NL = 0.;
    if (brow) {                                                        // _I_beta_hu (:691-704)
      const double cW0 = uk ? a.amplitude[p][0] / a.norm[p] : 1.;
      double integral = Bsum;
      if (a.mass[p]) integral += A * brow[0] * (row[0] * cW0) / M0;
      I_NL = bh * integral * a.rhom;
    }
    const double p2h = (bh * Iu + I_NL) * a.Pk_lin[ik];                // (:689)
    double p1h = W;                                                    // the profile at M = M_h (:662)
    if (a.k_trunc > 0.) {
      const double q = a.k[ik] / a.k_trunc;
      p1h *= 1. - exp(-(q * q));
    }
    double* o = a.out + (size_t)p * 3 * a.nk + ik;
    o[0] = p2h;
    o[a.nk] = p1h;
    o[2 * (size_t)a.nk] = p2h + p1h;
  }
}

extern "C" size_t hm_cross_halo_workspace(int32_t nM) { return ((size_t)nM + 4) * sizeof(double); }

extern "C" int hm_cross_halo_spectrum(const hm_model_desc* model_host, const double* k_dev, const double* Pk_lin_dev,
                                      const double* M_dev, const double* sigma_dev, int32_t nk, int32_t nM,
                                      const hm_tracer* tracers, const double* normalisation_host, int32_t n_tracers,
                                      const double* interp_weights_dev, const double* beta_dev, int32_t simple_twohalo,
                                      double k_trunc, double* out_dev, double* halo_bias_nu_dev, void* ws, size_t ws_bytes,
                                      void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(model_host && k_dev && Pk_lin_dev && M_dev && sigma_dev && tracers && normalisation_host && interp_weights_dev &&
                 out_dev && ws, "NULL pointer");
  HM_REQUIRE(nk > 0 && nM >= 2 && n_tracers >= 1 && n_tracers <= HM_MAX_TRACERS, "bad sizes (1..8 tracers per call)");
  HM_REQUIRE(!(beta_dev && simple_twohalo), "a simple two-halo term is not compatible with beta_NL");
  if (ws_bytes < hm_cross_halo_workspace(nM)) {
    hm_set_error("hm_cross_halo_spectrum: workspace needs %zu bytes", hm_cross_halo_workspace(nM));
    return HM_ERR_WORKSPACE;
  }
  hm_cross_args a;
  a.nk = nk; a.nM = nM; a.simple_twohalo = simple_twohalo; a.k_trunc = k_trunc; a.rhom = model_host->rhom;
  a.k = k_dev; a.Pk_lin = Pk_lin_dev; a.M = M_dev; a.L = interp_weights_dev; a.beta = beta_dev;
  a.w2 = (const double*)ws; a.scal = (const double*)ws + nM; a.out = out_dev;
  for (int p = 0; p < HM_MAX_TRACERS; ++p) {
    const bool on = p < n_tracers;
    a.gridW[p] = on ? (tracers[p].grid_mode == HM_GRID_SEPARATE ? tracers[p].wk_dev : tracers[p].grid_dev) : nullptr;
    a.amplitude[p] = on ? tracers[p].amplitude_dev : nullptr;
    a.norm[p] = on ? normalisation_host[p] : 1.;
    a.mode[p] = on ? tracers[p].grid_mode : 0;
    a.mass[p] = on ? tracers[p].mass_tracer : 0;
    if (on) {
      HM_REQUIRE(a.gridW[p], "tracer grid is NULL");
      HM_REQUIRE(a.amplitude[p] || !(a.mode[p] == HM_GRID_UK || simple_twohalo), "amplitude is required");
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  hm_cross_prep_kernel<<<1, 256, 0, st>>>(*model_host, M_dev, sigma_dev, interp_weights_dev, nM, (double*)ws, (double*)ws + nM);
  HM_LAUNCH_CHECK();
  hm_cross_kernel<<<dim3(nk, n_tracers), 128, 0, st>>>(a);
  HM_LAUNCH_CHECK();
  if (halo_bias_nu_dev) HM_CUDA_TRY(cudaMemcpyAsync(halo_bias_nu_dev, (const double*)ws + nM, 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return HM_OK;
}
