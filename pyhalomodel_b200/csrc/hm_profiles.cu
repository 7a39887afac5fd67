// hm_profiles.cu -- Fourier-space halo windows with a fused device Si/Ci.
//   hm_sici              : scipy.special.sici as used by the reference (pyhalomodel.py:1033-1035, 1044-1049)
//   hm_window_nfw        : _win_NFW         (pyhalomodel.py:1040-1062)
//   hm_window_isothermal : _win_isothermal  (pyhalomodel.py:1029-1037)
// FP64-ALU bound (two pairs of 12-term series + three sincos per NFW element); writes are coalesced along M.
#include "hm_common.cuh"
#include "hm_tables.h"

__constant__ double c_sici_map[HM_SICI_NI][2] = HM_SICI_MAP_INIT;
__constant__ double c_sici_A[HM_SICI_NI][HM_SICI_NT] = HM_SICI_A_INIT;
__constant__ double c_sici_B[HM_SICI_NI][HM_SICI_NT] = HM_SICI_B_INIT;

// Shared-memory copy of the tables, transposed to [term][interval] (padded to 32) so that lanes whose arguments
// fall in different intervals read different banks and lanes in the same interval broadcast.
#define HM_SICI_PAD 32
static_assert(HM_SICI_NI == 24, "hm_sici_interval decodes the 24 intervals tools/gen_tables.py lays out");
struct hm_sici_tables {
  double A[HM_SICI_NT][HM_SICI_PAD];
  double B[HM_SICI_NT][HM_SICI_PAD];
  double map[2][HM_SICI_PAD];
};

__device__ __forceinline__ void hm_sici_load_tables(hm_sici_tables& T) {
  for (int i = threadIdx.x; i < HM_SICI_NT * HM_SICI_PAD; i += blockDim.x) {
    const int n = i / HM_SICI_PAD, iv = i % HM_SICI_PAD;
    T.A[n][iv] = (iv < HM_SICI_NI) ? c_sici_A[iv][n] : 0.;
    T.B[n][iv] = (iv < HM_SICI_NI) ? c_sici_B[iv][n] : 0.;
  }
  for (int i = threadIdx.x; i < 2 * HM_SICI_PAD; i += blockDim.x) {
    const int j = i / HM_SICI_PAD, iv = i % HM_SICI_PAD;
    T.map[j][iv] = (iv < HM_SICI_NI) ? c_sici_map[iv][j] : 0.;
  }
  __syncthreads();
}

// Interval of x >= 0: [0, 4] | eight of width 1/2 in (4, 8) | eight of width 1 in [8, 16) | four of width 4 in [16, 32) |
// [32, 64) | [64, 128) | [128, inf) -- the edges sit on the binary exponent and the leading mantissa bits of x, so the
// index comes from the high word instead of a chain of comparisons (the finer the intervals, the shorter the series:
// 12 terms instead of the 20 a single [4, 6] interval needs).
__device__ __forceinline__ int hm_sici_interval(double x) {
  if (x <= 4.) return 0;
  const int hi = __double2hiint(x);
  const int e = (hi >> 20) - 1023;                   // 2 for [4, 8), ... (x > 4 here; inf and nan give 1024)
  if (e == 2) return 1 + ((hi >> 17) & 7);
  if (e == 3) return 9 + ((hi >> 17) & 7);
  if (e == 4) return 17 + ((hi >> 18) & 3);
  return (e == 5) ? 21 : (e == 6) ? 22 : 23;
}

// Si(x), Ci(x) for x >= 0 (odd / even continuation for x < 0, like Cephes).
//   x <= 4: Si = x A(t), Ci = gamma + ln x + x^2 B(t), t affine in x^2
//   x >  4: f = A(t)/x, g = B(t)/x^2, t affine in 1/x^2 on 23 intervals;  Si = pi/2 - f cos x - g sin x, Ci = f sin x - g cos x
// A(t), B(t) are short power series in t (Chebyshev fits converted exactly, tools/gen_tables.py: as accurate as the
// Clenshaw sum, half its operations); one Horner loop serves every interval, so mixed warps do not diverge.  Max error vs mpmath on [1e-9,1e3]:
// Si 4.4e-16 relative, Ci 3.6e-15 absolute -- the same class as scipy's routine (tools/gen_tables.py, tests/).
// CI = 0 skips what only Ci needs (the logarithm of the small-argument branch); CI = 2 returns, for x <= 4, only the series
// part x^2 B(t) of Ci and sets `small`, so that a caller who needs a DIFFERENCE of two small-argument Ci values can add
// ln(x1 / x2) once instead of two logarithms (hm_window_nfw_kernel: the ratio is 1 + c); CI = 3 returns Si(x) / x = A(t) in
// `ci` for x <= 4 (no Ci at all); HAVE_SC = true takes sin x and
// cos x from the caller (sx, cx: of |x|) instead of calling sincos; rx returns 1/|x| (always with NEED_RX, otherwise only
// where the large-argument branch needs it: the division is skipped for x <= 4).
template <int CI, bool HAVE_SC, bool NEED_RX>
__device__ __forceinline__ void hm_sici_core(double xin, const hm_sici_tables& T, double sx, double cx, double& si, double& ci,
                                             double& rx, bool& small) {
  constexpr bool WANT_CI = CI == 1 || CI == 2;
  const double x = fabs(xin);
  const int iv = hm_sici_interval(x);
  const double x2 = x * x;
  double r2 = 0.;
  rx = 0.;
  if (NEED_RX || iv != 0) {
    rx = 1. / x;                                     // one division serves 1/x, 1/x^2 and the caller
    r2 = rx * rx;
  }
  const double u = (iv == 0) ? x2 : r2;
  const double t = T.map[0][iv] * u + T.map[1][iv];
  double Av = T.A[HM_SICI_NT - 1][iv], Bv = T.B[HM_SICI_NT - 1][iv];
#pragma unroll
  for (int n = HM_SICI_NT - 2; n >= 0; --n) {        // Horner in t: one FMA per term and series
    Av = fma(Av, t, T.A[n][iv]);
    Bv = fma(Bv, t, T.B[n][iv]);
  }
  ci = 0.;
  small = iv == 0;
  if (iv == 0) {
    si = x * Av;
    if (CI == 1) ci = 0.57721566490153286061 + log(x) + x2 * Bv;   // -inf at x = 0
    if (CI == 2) ci = x2 * Bv;
    if (CI == 3) ci = Av;                              // Si(x) / x itself (hm_window_iso_kernel)
  } else {
    double s = sx, c = cx;
    if (!HAVE_SC) sincos(x, &s, &c);
    const double f = Av * rx, g = Bv * r2;
    si = M_PI_2 - f * c - g * s;
    if (WANT_CI) ci = f * s - g * c;
    if (isinf(x)) { si = M_PI_2; ci = 0.; }
  }
  if (xin < 0.) si = -si;
  if (isnan(xin)) { si = xin; ci = xin; }
}

__device__ __forceinline__ void hm_sici_eval(double xin, const hm_sici_tables& T, double& si, double& ci) {
  double rx;
  bool small;
  hm_sici_core<1, false, false>(xin, T, 0., 0., si, ci, rx, small);
}

__global__ void __launch_bounds__(256) hm_sici_kernel(const double* __restrict__ x, long long n,
                                                      double* __restrict__ si, double* __restrict__ ci) {
  __shared__ hm_sici_tables T;
  hm_sici_load_tables(T);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s, c;
    hm_sici_eval(x[i], T, s, c);
    if (si) si[i] = s;
    if (ci) ci[i] = c;
  }
}

// U_NFW(k, M) = [cos(ks)(Ci(ks+kv)-Ci(ks)) + sin(ks)(Si(ks+kv)-Si(ks)) - sin(kv)/(ks+kv)] / (ln(1+c) - c/(1+c))
// CTA = (sample, 256 masses); a thread keeps its mass and walks the wavenumbers, so r_s, the NFW factor and its
// reciprocal are computed once per thread instead of once per element, and the stores stay coalesced along M.  Two
// sincos per element serve both Si/Ci evaluations (round 1 called sincos up to three times and sin once).
// WITH_ISO: the isothermal window Si(kv) / kv of the same radii as well (matter NFW + isothermal galaxies is the sweep's
// pair of tracers): it shares the loop, the wavenumber loads and the sine of kv with the NFW element.
template <bool WITH_ISO>
__global__ void __launch_bounds__(256) hm_window_nfw_kernel(const double* __restrict__ k, const double* __restrict__ rv,
                                                            const double* __restrict__ conc, int nk, int nM,
                                                            long long total, double* __restrict__ out,
                                                            double* __restrict__ out_iso) {
  __shared__ hm_sici_tables T;
  hm_sici_load_tables(T);
  const int mblocks = (nM + 255) / 256;
  const long long gs = blockIdx.x / mblocks;
  const int m = (int)(blockIdx.x - gs * mblocks) * 256 + threadIdx.x;
  if (m >= nM || gs * (long long)nk * nM >= total) return;
  const double r_v = rv[gs * nM + m], c = conc[gs * nM + m];
  const double r_s = r_v / c;                         // :1045
  const double l1pc = log(1. + c);
  const double f4 = l1pc - c / (1. + c);              // :1062
  const double rf4 = 1. / f4;
  double* o = out + gs * (long long)nk * nM + m;
  double* oi = WITH_ISO ? out_iso + gs * (long long)nk * nM + m : nullptr;
  const int kper = (nk + (int)gridDim.y - 1) / (int)gridDim.y;        // wavenumbers per CTA row (small batches split k too)
  const int k0 = (int)blockIdx.y * kper, k1 = (k0 + kper < nk) ? k0 + kper : nk;
#pragma unroll 2
  for (int ik = k0; ik < k1; ++ik) {
    const double kk = k[ik];
    const double kv = kk * r_v, ks = kk * r_s;        // :1046-1047
    const double ksv = ks + kv;
    double s1, c1, s12 = 0., c12 = 0., Si_sv, Ci_sv, Si_s, Ci_s, r12, r1;
    sincos(ks, &s1, &c1);
    if (ksv > 4.) sincos(ksv, &s12, &c12);            // only the large-argument branch of Si/Ci uses them (a warp holds
                                                      // neighbouring masses of one wavenumber: the branch is mostly uniform)
    bool small_sv, small_s;
    hm_sici_core<2, true, true>(ksv, T, s12, c12, Si_sv, Ci_sv, r12, small_sv);      // :1048
    hm_sici_core<2, true, false>(ks, T, s1, c1, Si_s, Ci_s, r1, small_s);            // :1049
    // Ci(ks + kv) - Ci(ks): with both arguments on the small branch the logarithms combine to ln((ks + kv) / ks) = ln(1 + c),
    // a per-mass constant; otherwise (ks <= 4 < ks + kv) the small one gets its gamma + ln x back
    double dCi;
    if (small_sv) dCi = l1pc + (Ci_sv - Ci_s);          // (ks < ks + kv: both small)
    else {
      if (small_s) Ci_s = 0.57721566490153286061 + log(ks) + Ci_s;
      dCi = Ci_sv - Ci_s;
    }
    const double f1 = c1 * dCi;
    const double f2 = s1 * (Si_sv - Si_s);
    double sv, cv = 0.;
    if (WITH_ISO && kv > 4.) sincos(kv, &sv, &cv);
    else sv = sin(kv);
    const double f3 = sv * r12;                       // sin(kv)/(ks+kv) with the reciprocal of the Si/Ci evaluation (an
                                                      // angle-subtraction sin(kv) costs 1e-11 where U cancels to 3e-5)
    o[(size_t)ik * nM] = (f1 + f2 - f3) * rf4;
    if (WITH_ISO) {                                   // Si(kv) / kv (:1036), as in hm_window_iso_kernel
      double si, ci, rx;
      bool small;
      hm_sici_core<3, true, false>(kv, T, sv, cv, si, ci, rx, small);
      oi[(size_t)ik * nM] = small ? ci : si * rx;
    }
  }
}

__global__ void __launch_bounds__(256) hm_window_iso_kernel(const double* __restrict__ k, const double* __restrict__ rv,
                                                            int nk, int nM, long long total, double* __restrict__ out) {
  __shared__ hm_sici_tables T;
  hm_sici_load_tables(T);
  const int mblocks = (nM + 255) / 256;
  const long long gs = blockIdx.x / mblocks;
  const int m = (int)(blockIdx.x - gs * mblocks) * 256 + threadIdx.x;
  if (m >= nM || gs * (long long)nk * nM >= total) return;
  const double r_v = rv[gs * nM + m];
  double* o = out + gs * (long long)nk * nM + m;
  const int kper = (nk + (int)gridDim.y - 1) / (int)gridDim.y;
  const int k0 = (int)blockIdx.y * kper, k1 = (k0 + kper < nk) ? k0 + kper : nk;
#pragma unroll 2
  for (int ik = k0; ik < k1; ++ik) {
    const double kv = k[ik] * r_v;
    double si, ci, rx;
    bool small;
    hm_sici_core<3, false, false>(kv, T, 0., 0., si, ci, rx, small);
    // Si(kv) / kv (:1036): the series itself for kv <= 4, else with the reciprocal the evaluation already has
    o[(size_t)ik * nM] = small ? ci : si * rx;
  }
}

// (sample, 256-mass block) CTAs along x; the wavenumber axis is split along y as far as needed for a launch of many waves of
// CTAs (a thread's per-mass constants are cheap next to the wavenumbers it still walks)
static dim3 hm_window_grid(int nk, int nM, int G) {
  const long long bx = (long long)G * ((nM + 255) / 256);
#ifndef HM_WINDOW_WAVES
#define HM_WINDOW_WAVES 16       // CTAs per resident CTA slot (4 per SM): enough of them that the last wave is a small part --
#endif                           // 2048 equal CTAs on 592 slots are 3.46 waves, i.e. four, 14 % of them empty
  long long by = ((long long)hm_num_sms() * 4 * HM_WINDOW_WAVES + bx - 1) / bx;
  if (by > nk) by = nk;
  if (by < 1) by = 1;
  return dim3((unsigned)bx, (unsigned)by);
}

static int hm_grid_for(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)hm_num_sms() * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

extern "C" int hm_sici(const double* x_dev, int64_t n, double* si_dev, double* ci_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(x_dev && n >= 0, "NULL pointer or negative length");
  if (n == 0) return HM_OK;
  hm_sici_kernel<<<hm_grid_for(n), 256, 0, (cudaStream_t)stream>>>(x_dev, n, si_dev, ci_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

extern "C" int hm_window_nfw(const double* k_dev, const double* rv_dev, const double* c_dev, int32_t nk, int32_t nM,
                             int32_t G, double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(k_dev && rv_dev && c_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nk > 0 && nM > 0 && G > 0, "empty grid");
  const long long total = (long long)G * nk * nM;
  hm_window_nfw_kernel<false><<<hm_window_grid(nk, nM, G), 256, 0, (cudaStream_t)stream>>>(k_dev, rv_dev, c_dev, nk, nM, total,
                                                                                          out_dev, nullptr);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

extern "C" int hm_window_nfw_isothermal(const double* k_dev, const double* rv_dev, const double* c_dev, int32_t nk, int32_t nM,
                                        int32_t G, double* nfw_out_dev, double* iso_out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(k_dev && rv_dev && c_dev && nfw_out_dev && iso_out_dev, "NULL pointer");
  HM_REQUIRE(nk > 0 && nM > 0 && G > 0, "empty grid");
  const long long total = (long long)G * nk * nM;
  hm_window_nfw_kernel<true><<<hm_window_grid(nk, nM, G), 256, 0, (cudaStream_t)stream>>>(k_dev, rv_dev, c_dev, nk, nM, total,
                                                                                         nfw_out_dev, iso_out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

extern "C" int hm_window_isothermal(const double* k_dev, const double* rv_dev, int32_t nk, int32_t nM, int32_t G,
                                    double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(k_dev && rv_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nk > 0 && nM > 0 && G > 0, "empty grid");
  const long long total = (long long)G * nk * nM;
  hm_window_iso_kernel<<<hm_window_grid(nk, nM, G), 256, 0, (cudaStream_t)stream>>>(k_dev, rv_dev, nk, nM, total, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Configuration-space profiles: W(k, M) = int_0^rv j0(k r) P(r; M) dr on the reference's fixed radial grid
// R = linspace(0, rv, nr) with a sample-based rule (profile.configuration / _halo_window, pyhalomodel.py:886-961,
// win_integration in {trapezoid, simps, romb}: all three are fixed linear combinations of the samples, passed
// here as quadrature weights per unit spacing).  P is sampled by the caller (it is a Python callable there).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hm_window_config_kernel(const double* __restrict__ k, const double* __restrict__ rv,
                                                               const double* __restrict__ Psamp,
                                                               const double* __restrict__ wq, int nk, int nM, int nr,
                                                               double* __restrict__ out) {
  const long long total = (long long)nk * nM;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ik = (int)(i / nM), m = (int)(i - (long long)ik * nM);
    const double kk = k[ik], r_v = rv[m];
    const double dR = r_v / (double)(nr - 1);
    const double* P = Psamp + (size_t)m * nr;
    double acc = 0.;
    for (int j = 0; j < nr; ++j) {
      const double R = (j == nr - 1) ? r_v : dR * (double)j;       // np.linspace hits the end point exactly
      const double x = kk * R;
      const double j0 = (x == 0.) ? 1. : sin(x) / x;               // spherical_jn(0, x)
      acc = fma(wq[j], j0 * P[j], acc);
    }
    out[i] = acc * dR;
  }
}

extern "C" int hm_window_configuration(const double* k_dev, const double* rv_dev, const double* profile_samples_dev,
                                       const double* weights_dev, int32_t nk, int32_t nM, int32_t nr,
                                       double* out_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(k_dev && rv_dev && profile_samples_dev && weights_dev && out_dev, "NULL pointer");
  HM_REQUIRE(nk > 0 && nM > 0 && nr >= 2, "bad sizes");
  hm_window_config_kernel<<<hm_grid_for((long long)nk * nM), 256, 0, (cudaStream_t)stream>>>(
      k_dev, rv_dev, profile_samples_dev, weights_dev, nk, nM, nr, out_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
