/*
 * halob200.h -- C ABI of libhalob200.so, the B200 (sm_100a) halo-model engine.
 *
 * The reference (alexander-mead/pyhalomodel) has no FFI: its boundary is the Python API
 * (pyhalomodel/__init__.py:2-3).  This header is the boundary a maintainer would bind with ctypes to replace
 * the Python loops of that API; each entry point cites the reference lines it replaces.  INTEGRATION.md shows
 * the binding.
 *
 * Conventions
 *   - all floating-point data are IEEE float64; all array arguments are BORROWED pointers: the library never
 *     frees or retains them, and allocates nothing after the first call;
 *   - pointers named *_dev are CUDA device pointers, *_host are host pointers (pinned memory recommended);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); entry points only enqueue
 *     work unless their comment says they synchronise;
 *   - every function returns 0 on success or a negative hm_status; hm_last_error() describes the last failure
 *     of the calling thread.  Nothing throws or aborts across this boundary.
 *   - row-major layouts; a profile grid is [nk][nM] with M contiguous (pyhalomodel.py:872).
 */
#ifndef HALOB200_H
#define HALOB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HM_ABI_VERSION 1

typedef enum hm_status {
  HM_OK = 0,
  HM_ERR_INVALID = -1,     /* bad argument (NULL pointer, size, family, unsupported P) */
  HM_ERR_CUDA = -2,        /* a CUDA runtime call or kernel launch failed              */
  HM_ERR_WORKSPACE = -3,   /* workspace too small / misaligned                         */
  HM_ERR_NODEVICE = -4     /* no CUDA device: there is NO CPU fallback                 */
} hm_status;

/* Mass-function families, pyhalomodel.py:62-67 (names) and :92-190 (constants). */
enum {
  HM_PS74 = 0,        /* 'Press & Schecter (1974)'   */
  HM_ST99 = 1,        /* 'Sheth & Tormen (1999)'     */
  HM_SMT01 = 2,       /* 'Sheth, Mo & Tormen (2001)' */
  HM_TINKER10 = 3,    /* 'Tinker et al. (2010)', calibrated bias (:273-282) */
  HM_DESPALI16 = 4,   /* 'Despali et al. (2016)'     */
  HM_TINKER10_PBS = 5 /* Tinker10 with the peak-background-split bias (module switch Tinker_PBS, :22,:264-271) */
};

/* Scalar state of one reference `model` instance (pyhalomodel.py:73-79 and :101-184).
 *   ST99/SMT01/DESPALI16: par[0]=A_ST par[1]=q_ST par[2]=p_ST ; SMT01 also par[3]=a par[4]=b par[5]=c
 *   TINKER10(_PBS):       par[0..4]=alpha,beta,gamma,phi,eta ; par[5..10]=A,a,B,b,C,c of the calibrated bias */
typedef struct hm_model_desc {
  int32_t family;
  int32_t reserved;
  double dc;          /* linear collapse threshold, nu = dc/sigma (pyhalomodel.py:706-711) */
  double rhom;        /* comoving matter density rho_crit*Om_m (cosmology.py:28-34)        */
  double par[12];
} hm_model_desc;

/* How the two grids of a tracer relate (data semantics of profile.Fourier, pyhalomodel.py:874-881). */
enum {
  HM_GRID_SEPARATE = 0,  /* profile(k,M,Uk,Wk,...) called directly: Wk and Uk are independent grids (:812-819)   */
  HM_GRID_UK = 1,        /* Fourier(amplitude given): grid = Uk; Wk = (amplitude*Uk)/normalisation (:880-881)     */
  HM_GRID_WK = 2         /* Fourier(amplitude=None): grid = Wk (dimensionful input); amplitude = grid[0,:],
                            Uk = grid/amplitude (:874-878)                                                        */
};

/* One halo profile ("tracer") for a batch of models.  G = number of samples (grid sets). */
typedef struct hm_tracer {
  const double* grid_dev;       /* [G][nk][nM]  Uk (HM_GRID_UK, HM_GRID_SEPARATE) or Wk (HM_GRID_WK)              */
  const double* wk_dev;         /* [G][nk][nM]  Wk, only for HM_GRID_SEPARATE, else NULL                          */
  const double* amplitude_dev;  /* [G][nM]      profile.amplitude (for HM_GRID_WK this is grid[0,:], may be NULL
                                                 and is then read from the grid)                                   */
  const double* variance_dev;   /* [G][nM] or NULL (profile.variance, :472-473)                                   */
  const double* normalisation_dev; /* [B] per model (e.g. rho_gal = average(N) depends on the mass function)      */
  int32_t grid_mode;            /* HM_GRID_*                                                                      */
  int32_t mass_tracer;          /* adds A*W(k,M0)/M0 to the two-halo integral (:541-542)                          */
  int32_t discrete_tracer;      /* N(N-1) and shot-noise handling (:466-491)                                      */
  int32_t reserved;
} hm_tracer;

/* A batch of halo-model evaluations: B = G*models_per_sample models; model b uses sample b/models_per_sample
 * (e.g. the 5 mass functions evaluated on one cosmology share its grids, sigma(M) and P_lin). */
typedef struct hm_problem {
  int32_t nk, nM;
  int32_t n_tracers;            /* P, 1..HM_MAX_TRACERS */
  int32_t n_samples;            /* G */
  int32_t models_per_sample;    /* F >= 1 */
  int32_t simple_twohalo;       /* power_spectrum(..., simple_twohalo=)      pyhalomodel.py:442-446 */
  int32_t subtract_shotnoise;   /* power_spectrum(..., subtract_shotnoise=)  :484-491               */
  int32_t correct_discrete;     /* power_spectrum(..., correct_discrete=)    :466-471               */
  double k_trunc;               /* <= 0 means None (:494-495)                                       */
  const double* k_dev;          /* [nk]   only read when k_trunc > 0                                */
  const double* M_dev;          /* [nM]   strictly increasing (checked by the caller, :387-389)     */
  const double* sigma_dev;      /* [G][nM]                                                          */
  const double* Pk_lin_dev;     /* [G][nk]                                                          */
  const hm_model_desc* models_dev; /* [B] device array, or NULL to use `model_host` (B must be 1)   */
  hm_model_desc model_host;
  hm_tracer tracers[8];
  /* optional non-linear halo bias (power_spectrum(beta=), model._I_beta, pyhalomodel.py:546-571) */
  const double* beta_dev;       /* [G][nk][nM][nM] beta_NL(k, M1, M2), or NULL                              */
  int32_t beta_do_I11;          /* module switch do_I11      (pyhalomodel.py:45)                             */
  int32_t beta_do_I12_I21;      /* module switch do_I12_I21  (pyhalomodel.py:46)                             */
  /* out_dev lives on a peer GPU (gather fused into the epilogue's stores): the epilogue then ends with a
   * system-scope fence so that a barrier after the kernel is enough for the reader */
  int32_t output_is_peer;
  int32_t reserved2;
} hm_problem;
#define HM_MAX_TRACERS 8

int hm_abi_version(void);
const char* hm_last_error(void);
/* Number of CUDA devices visible (0 if none); never fails. */
int hm_device_count(void);

/* ---- mass function and bias -------------------------------------------------------------------------------
 * g(nu) and b(nu) element-wise (model._mass_function_nu :217-240, model._linear_bias_nu :242-284).
 * g_dev / b_dev may be NULL to skip one of them. */
int hm_massfn_nu(const hm_model_desc* model_host, const double* nu_dev, int64_t n,
                 double* g_dev, double* b_dev, void* stream);

/* A[b] = 1 - int_{nu0[b]}^{inf} g(nu) b(nu) dnu (power_spectrum :413-414; QUADPACK there, fixed composite
 * Gauss-Legendre here).  models_dev: [B] device array. */
int hm_low_mass_correction(const hm_model_desc* models_dev, const double* nu0_dev, int32_t B,
                           double* A_dev, void* stream);

/* <f_j> = rhom*trapz((f_j/M) g(nu), nu) for nf functions of one model (model.average :335-359).
 * f_dev: [nf][nM]; out_dev: [nf]. */
int hm_average(const hm_model_desc* model_host, const double* M_dev, const double* sigma_dev, int32_t nM,
               const double* f_dev, int32_t nf, double* out_dev, void* stream);

/* Batched model.average for sweeps: out[b] = <f_s> under model b, with sample s = b / models_per_sample.
 * models_dev: [B]; sigma_dev, f_dev: [n_samples][nM]; out_dev: [B]. */
int hm_average_batch(const hm_model_desc* models_dev, const double* M_dev, const double* sigma_dev, int32_t nM,
                     int32_t n_samples, int32_t models_per_sample, const double* f_dev, double* out_dev, void* stream);

/* out[j][k] = sum_m w[j][m] grid[k][m]: weighted row sums of one (k,M) grid for nw weight vectors (used by
 * model.cross_halo_spectrum, pyhalomodel.py:573-704).  grid_dev: [nk][nM]; w_dev: [nw][nM]; out_dev: [nw][nk]. */
int hm_weighted_rowsum(const double* grid_dev, const double* w_dev, int32_t nk, int32_t nM, int32_t nw,
                       double* out_dev, void* stream);

/* M^2 n(M)/rhobar and n(M) (model.multiplicity_function / mass_function, pyhalomodel.py:286-315, with the three-point
 * derivative of ln sigma against ln R of utility.derivative_from_samples, utility.py:15-35).  F_dev / n_dev: [nM],
 * either may be NULL. */
int hm_multiplicity_function(const hm_model_desc* model_host, const double* M_dev, const double* sigma_dev,
                             int32_t nM, double* F_dev, double* n_dev, void* stream);

/* Every tracer crossed with haloes of one mass M_h (model.cross_halo_spectrum, pyhalomodel.py:573-704).
 * tracers: HOST array [n_tracers] (device pointers inside; grids [nk][nM], amplitudes [nM]); normalisation_host:
 * [n_tracers]; interp_weights_dev: [nM] weights L with f(M_h) = sum_m L[m] f(M[m]) (the reference's cubic
 * interpolation in ln M, :631-639, is linear in the samples); beta_dev: [nk][nM] beta_NL(k, M) at M_h or NULL;
 * k_trunc <= 0 means None.  out_dev: [n_tracers][3][nk] (2h, 1h, hm).  halo_bias_nu_dev (optional, [3]) receives
 * A, nu(M_h), b(M_h).  Workspace: hm_cross_halo_workspace(nM) bytes. */
size_t hm_cross_halo_workspace(int32_t nM);
int hm_cross_halo_spectrum(const hm_model_desc* model_host, const double* k_dev, const double* Pk_lin_dev,
                           const double* M_dev, const double* sigma_dev, int32_t nk, int32_t nM,
                           const hm_tracer* tracers, const double* normalisation_host, int32_t n_tracers,
                           const double* interp_weights_dev, const double* beta_dev, int32_t simple_twohalo,
                           double k_trunc, double* out_dev, double* halo_bias_nu_dev, void* workspace_dev,
                           size_t workspace_bytes, void* stream);

/* ---- the hot path: model.power_spectrum (pyhalomodel.py:361-513, beta=None) --------------------------------
 * out_dev: [B][S][3][nk] with S = P(P+1)/2 unique pairs ordered (0,0),(0,1),..,(0,P-1),(1,1),.. and the three
 * terms ordered (Pk_2h, Pk_1h, Pk_hm).  lowmass_dev (optional, [B]) receives A so the caller can raise the
 * reference's RuntimeWarning for A<0 (:418-420).  Workspace: hm_power_spectrum_workspace() bytes, 256-aligned. */
size_t hm_power_spectrum_workspace(const hm_problem* prob);
int hm_power_spectrum(const hm_problem* prob, double* out_dev, double* lowmass_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* The two stages of hm_power_spectrum separately (for profiling / reuse of the weights across calls):
 * stage 1 builds the per-model mass-quadrature weight vectors, stage 2 streams the (k,M) grids. */
int hm_halo_weights(const hm_problem* prob, double* lowmass_dev, void* workspace_dev, size_t workspace_bytes,
                    void* stream);
int hm_mass_quadrature(const hm_problem* prob, double* out_dev, const void* workspace_dev, size_t workspace_bytes,
                       void* stream);
/* Same, launching only part of stage 2 (for timing the dominant kernel alone): bit 0 = the streaming kernel that
 * reduces the (k,M) grids to partial row sums, bit 1 = the epilogue kernel that forms the spectra. */
int hm_mass_quadrature_stages(const hm_problem* prob, double* out_dev, const void* workspace_dev,
                              size_t workspace_bytes, void* stream, int stages);

/* ---- many tracers: the one-halo cross terms as a per-wavenumber Gram contraction on the FP64 tensor cores ----------
 * Same semantics and output layout as hm_power_spectrum for 2..HM_GRAM_MAX_TRACERS single-grid tracers
 * (grid_mode HM_GRID_UK or HM_GRID_WK).  `tracers` is a HOST array of n_tracers descriptors whose pointer fields
 * are device pointers.  Replaces the P(P+1)/2 x nk Python iterations of pyhalomodel.py:433-507. */
#define HM_GRAM_MAX_TRACERS 64
typedef struct hm_gram_problem {
  int32_t nk, nM;
  int32_t n_tracers;
  int32_t n_samples;
  int32_t models_per_sample;
  int32_t simple_twohalo, subtract_shotnoise, correct_discrete;
  double k_trunc;
  const double* k_dev;
  const double* M_dev;
  const double* sigma_dev;
  const double* Pk_lin_dev;
  const hm_model_desc* models_dev;
  hm_model_desc model_host;
  const hm_tracer* tracers;      /* host array [n_tracers] */
  int32_t output_is_peer;        /* as in hm_problem */
  int32_t reserved2;
} hm_gram_problem;
size_t hm_gram_workspace(const hm_gram_problem* prob);
int hm_gram_power_spectrum(const hm_gram_problem* prob, double* out_dev, double* lowmass_dev,
                           void* workspace_dev, size_t workspace_bytes, void* stream);
/* Same, launching only part of the pipeline (for timing the contraction alone): bit 0 = the weight kernels,
 * bit 1 = the Gram (FP64 DMMA) kernel, bit 2 = the epilogue. */
int hm_gram_power_spectrum_stages(const hm_gram_problem* prob, double* out_dev, double* lowmass_dev,
                                  void* workspace_dev, size_t workspace_bytes, void* stream, int stages);

/* ---- Fourier profiles ------------------------------------------------------------------------------------
 * Si(x), Ci(x) element-wise (scipy.special.sici as used at pyhalomodel.py:1033-1035,1044-1049). */
int hm_sici(const double* x_dev, int64_t n, double* si_dev, double* ci_dev, void* stream);
/* U(k,M) for 'NFW' (_win_NFW :1040-1055), 'isothermal' (_win_isothermal :1029-1037).  rv_dev, c_dev: [G][nM];
 * out_dev: [G][nk][nM].  k_dev: [nk]. */
int hm_window_nfw(const double* k_dev, const double* rv_dev, const double* c_dev, int32_t nk, int32_t nM,
                  int32_t G, double* out_dev, void* stream);
int hm_window_isothermal(const double* k_dev, const double* rv_dev, int32_t nk, int32_t nM, int32_t G,
                         double* out_dev, void* stream);
/* Both windows of the same radii in one pass (window_function(k, rv, c, 'NFW') and window_function(k, rv,
 * profile='isothermal'), pyhalomodel.py:1006-1062): the matter + galaxy pair of a sweep. */
int hm_window_nfw_isothermal(const double* k_dev, const double* rv_dev, const double* c_dev, int32_t nk, int32_t nM,
                             int32_t G, double* nfw_out_dev, double* iso_out_dev, void* stream);

/* W(k,M) = int_0^rv j0(k r) P(r;M) dr on R = linspace(0, rv, nr) by a sample rule (profile.configuration with
 * win_integration in {trapezoid, simps, romb}; _halo_window, pyhalomodel.py:930-961).  profile_samples_dev:
 * [nM][nr] = P(R_j; M); weights_dev: [nr] quadrature weights per unit spacing; out_dev: [nk][nM]. */
int hm_window_configuration(const double* k_dev, const double* rv_dev, const double* profile_samples_dev,
                            const double* weights_dev, int32_t nk, int32_t nM, int32_t nr, double* out_dev,
                            void* stream);

/* ---- sigma(R) ----------------------------------------------------------------------------------------------
 * sigma(R) = sqrt( sum_j dlnk k_j^3 P(k_j) T(k_j R)^2 / (2 pi^2) ) on the caller's integration grid
 * (cosmology._sigmaR_brute_log, cosmology.py:89-106; T with the Taylor switch of :41-48).
 * kgrid_dev: [nkg]; Pk_dev: [G][nkg]; R_dev: [G][nR]; out_dev: [G][nR]. */
int hm_sigma_tophat(const double* kgrid_dev, const double* Pk_dev, double dlnk, int32_t nkg,
                    const double* R_dev, int32_t nR, int32_t G, double* out_dev, void* stream);

/* ---- measurement ------------------------------------------------------------------------------------------
 * FP64 roofline denominators of the current device, in TFLOP/s (2 flops per FMA): DFMA issue throughput and DMMA
 * (mma.sync m8n8k4 f64) throughput.  Host pointers; synchronises `stream`; allocates and frees a scratch buffer. */
int hm_measure_fp64_peaks(double* dfma_tflops_host, double* dmma_tflops_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HALOB200_H */
