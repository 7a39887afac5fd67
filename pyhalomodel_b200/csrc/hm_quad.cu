// hm_quad.cu -- the hot path of model.power_spectrum (pyhalomodel.py:361-513, beta=None) as two kernels:
//
//   hm_weights_kernel   one CTA per model: nu = dc/sigma, g(nu), b(nu), trapezoid-in-nu weights, the low-mass
//                       correction A, and from those the P+S per-mass weight vectors that turn every integral of
//                       the reference into a plain weighted row sum over M:
//                         I_u(k)   = rhom * sum_m  e_u[m]  * G_u[k,m]                      (:535-544, incl. A W(k,M0)/M0)
//                         P1h_uu   = rhom * sum_m  d_u[m]  * G_u[k,m]^2                    (:465-477, :526-533)
//                         P1h_uv   = rhom * sum_m  q_uv[m] * G_u[k,m] * G_v[k,m]           (:479)
//                       where G_u is the single grid a Fourier profile was built from (Uk, or Wk when
//                       amplitude=None; :874-881), so each (k,M) grid is read from HBM exactly once.
//   hm_quad_kernel      streams the grids: a CTA takes a tile of R wavenumber rows, its threads split the mass
//                       axis with 16-byte coalesced loads, keep R*(P+S) partial sums in registers, reduce them
//                       with a transposing warp-shuffle network plus one shared-memory pass across warps, and the
//                       epilogue forms P_2h = I_u I_v P_lin, the discrete/shot-noise/k_trunc fix-ups and P_hm.
//
// HBM-bound: algorithmic bytes per model = 8*(P*nk*nM + (P+S)*nM + nk + 3*S*nk).
#include "hm_common.cuh"
#include "hm_tables.h"

#define HM_MAXP HM_MAX_TRACERS

__constant__ double c_gl_x[16] = HM_GL16_X_INIT;
__constant__ double c_gl_w[16] = HM_GL16_W_INIT;

// ---------------------------------------------------------------------------------------------------------------
// Low-mass correction: A = 1 - int_{nu0}^{inf} g b dnu                                   pyhalomodel.py:413-414
// Composite 16-point Gauss-Legendre on 10 geometrically growing panels covering [nu0, nu0+40]; the integrand is
// < 1e-300 beyond.  Agrees with scipy.integrate.quad to ~1e-14 absolute for all families (tests/).
// Thread t in [0,160) owns one node; returns that node's contribution (0 for other threads).
// ---------------------------------------------------------------------------------------------------------------
#define HM_GL_PANELS 10
#define HM_GL_NODES (16 * HM_GL_PANELS)
__device__ __forceinline__ double hm_lowmass_node(const hm_model_desc& d, double nu0, int t) {
  if (t >= HM_GL_NODES) return 0.;
  const int panel = t >> 4, node = t & 15;
  const double unit = 40. / (double)((1 << HM_GL_PANELS) - 1);
  const double a = nu0 + unit * (double)((1 << panel) - 1);
  const double b = nu0 + unit * (double)((1 << (panel + 1)) - 1);
  const double half = 0.5 * (b - a);
  const double x = 0.5 * (a + b) + half * c_gl_x[node];
  return half * c_gl_w[node] * hm_g_nu(d, x) * hm_b_nu(d, x);
}

__global__ void hm_lowmass_kernel(const hm_model_desc* __restrict__ models, const double* __restrict__ nu0,
                                  int B, double* __restrict__ A) {
  __shared__ double scratch[32];
  const int b = blockIdx.x;
  if (b >= B) return;
  const hm_model_desc d = models[b];
  const double part = hm_lowmass_node(d, nu0[b], threadIdx.x);
  const double tot = hm_block_sum(part, scratch);
  if (threadIdx.x == 0) A[b] = 1. - tot;
}

// ---------------------------------------------------------------------------------------------------------------
// Workspace layout (doubles): weights [B][NW][nMp] then scalars [B][HM_NSCAL]
// ---------------------------------------------------------------------------------------------------------------
#define HM_NSCAL (2 + 2 * HM_MAXP)   // A, rhom, SN_p[8], Isimple_p[8]

struct hm_layout {
  int P, S, NW, nMp;
  size_t weights_doubles, scal_offset, total_bytes;
};

static hm_layout hm_make_layout(const hm_problem* pr) {
  hm_layout L;
  L.P = pr->n_tracers;
  L.S = L.P * (L.P + 1) / 2;
  L.NW = L.P + L.S;
  L.nMp = (pr->nM + 1) & ~1;
  const size_t B = (size_t)pr->n_samples * (size_t)pr->models_per_sample;
  L.weights_doubles = B * (size_t)L.NW * (size_t)L.nMp;
  L.scal_offset = (L.weights_doubles + 31) & ~(size_t)31;
  L.total_bytes = (L.scal_offset + B * HM_NSCAL) * sizeof(double);
  return L;
}

struct hm_weights_args {
  int nM, nMp, P, NW, F, B;
  int simple_twohalo, correct_discrete;
  const double* M;
  const double* sigma;             // [G][nM]
  const hm_model_desc* models;     // [B] or NULL
  hm_model_desc single;
  const double* amplitude[HM_MAXP];  // [G][nM] (or grid row 0 for HM_GRID_WK with NULL amplitude)
  size_t amp_stride[HM_MAXP];        // doubles between samples
  const double* variance[HM_MAXP];
  const double* norm[HM_MAXP];       // [B]
  int mode[HM_MAXP], mass[HM_MAXP], discrete[HM_MAXP];
  double* weights;
  double* scal;
  double* lowmass;                   // optional [B]
};

__global__ void __launch_bounds__(256) hm_weights_kernel(const hm_weights_args a) {
  __shared__ double scratch[32];
  const int b = blockIdx.x;
  const int s = b / a.F;
  const hm_model_desc d = a.models ? a.models[b] : a.single;
  const double* sig = a.sigma + (size_t)s * a.nM;
  const double M0 = a.M[0];

  // A = 1 - int g b dnu from nu[0] (pyhalomodel.py:413-414)
  const double nu_first = d.dc / sig[0];
  const double A = 1. - hm_block_sum(hm_lowmass_node(d, nu_first, threadIdx.x), scratch);
  const double Aterm = A / M0;   // multiplies W(k, M[0]) for mass tracers (:541-542)

  double sn[HM_MAXP], simp[HM_MAXP];
#pragma unroll
  for (int p = 0; p < HM_MAXP; ++p) sn[p] = simp[p] = 0.;

  double* wout = a.weights + (size_t)b * a.NW * a.nMp;
  for (int m = threadIdx.x; m < a.nMp; m += blockDim.x) {
    if (m >= a.nM) {   // padding column
      for (int i = 0; i < a.NW; ++i) wout[(size_t)i * a.nMp + m] = 0.;
      continue;
    }
    const double nu = d.dc / sig[m];
    const double nu_hi = (m + 1 < a.nM) ? d.dc / sig[m + 1] : nu;
    const double nu_lo = (m > 0) ? d.dc / sig[m - 1] : nu;
    const double tw = 0.5 * (nu_hi - nu_lo);   // trapezoid rule in nu as a weight (scipy trapezoid, :27)
    const double g = hm_g_nu(d, nu);
    const double bg = hm_b_nu(d, nu) * g;
    const double Mm = a.M[m];
    const double w1 = tw * (g / Mm);           // one-halo weight  (:530)
    const double w2 = tw * (bg / Mm) + ((m == 0) ? Aterm : 0.);   // two-halo weight; low-mass term only if mass tracer
    const double w2_plain = tw * (bg / Mm);

    double cW[HM_MAXP];
#pragma unroll
    for (int p = 0; p < HM_MAXP; ++p) {
      if (p < a.P) {
        const double amp = a.amplitude[p][(size_t)s * a.amp_stride[p] + m];
        const double norm = a.norm[p][b];
        // <N(N-1)> or <N^2> plus variance (:466-473); no FMA contraction so exact zeros stay exact
        double fac = (a.correct_discrete && a.discrete[p]) ? __dmul_rn(amp, __dadd_rn(amp, -1.)) : __dmul_rn(amp, amp);
        if (a.variance[p]) fac = __dadd_rn(fac, a.variance[p][(size_t)s * a.nM + m]);
        const double an = amp / norm;
        double uS;   // U/normalisation = uS * grid
        if (a.mode[p] == HM_GRID_UK) { cW[p] = an; uS = 1. / norm; }
        else if (a.mode[p] == HM_GRID_WK) { cW[p] = 1.; uS = 1. / (amp * norm); }
        else { cW[p] = 1.; uS = 1. / norm; }
        const double w2p = a.mass[p] ? w2 : w2_plain;
        wout[(size_t)p * a.nMp + m] = w2p * cW[p];                       // e_p
        wout[(size_t)(a.P + p) * a.nMp + m] = __dmul_rn(__dmul_rn(w1, fac), uS * uS);   // d_p
        sn[p] += w1 * (amp / (norm * norm));                             // shot noise integrand (:425)
        simp[p] += w2p * an;                                             // simple two-halo (:442-446)
      }
    }
    int idx = 2 * a.P;
    for (int u = 0; u < a.P; ++u)
      for (int v = u + 1; v < a.P; ++v) wout[(size_t)(idx++) * a.nMp + m] = w1 * cW[u] * cW[v];   // q_uv
  }

  double* sc = a.scal + (size_t)b * HM_NSCAL;
  for (int p = 0; p < a.P; ++p) {
    const double t1 = hm_block_sum(sn[p], scratch);
    const double t2 = hm_block_sum(simp[p], scratch);
    if (threadIdx.x == 0) {
      sc[2 + p] = a.discrete[p] ? t1 * d.rhom : 0.;
      sc[2 + HM_MAXP + p] = t2 * d.rhom;
    }
  }
  if (threadIdx.x == 0) {
    sc[0] = A;
    sc[1] = d.rhom;
    if (a.lowmass) a.lowmass[b] = A;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The streaming quadrature kernel
// ---------------------------------------------------------------------------------------------------------------
struct hm_quad_args {
  int nk, nM, nMp, F, B, ntiles;
  long long nitems;
  int simple_twohalo, subtract_shotnoise, correct_discrete;
  double k_trunc;
  const double* k;
  const double* Pk_lin;              // [G][nk]
  const double* gridW[HM_MAXP];      // two-halo / cross source
  const double* gridU[HM_MAXP];      // auto one-halo source (== gridW unless HM_GRID_SEPARATE)
  int discrete[HM_MAXP];
  const double* weights;
  const double* scal;
  double* out;                       // [B][S][3][nk]
};

// v[0..32) per lane  ->  lane l holds sum over the warp of v[l]  (31 shuffles instead of 160).
__device__ __forceinline__ double hm_warp_transpose_sum32(double (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const double send = upper ? v[i] : v[i + off];
      const double keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int P, int R, int VEC, bool SEP>
__global__ void __launch_bounds__(256) hm_quad_kernel(const hm_quad_args a) {
  constexpr int S = P * (P + 1) / 2;
  constexpr int NW = P + S;
  constexpr int NV = R * NW;
  constexpr int NG = (NV + 31) / 32;   // groups of 32 values for the transposing reduction
  __shared__ double red[8][NG * 32];
  __shared__ double fin[NG * 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int ncol = (VEC == 2) ? (a.nM >> 1) : a.nM;
  const size_t rowstride = (size_t)a.nM;

  // contiguous chunk of (model, row-tile) items per CTA keeps the weight vectors of one model hot in L1
  const long long per = (a.nitems + gridDim.x - 1) / gridDim.x;
  long long item = (long long)blockIdx.x * per;
  const long long item_end = (item + per < a.nitems) ? item + per : a.nitems;

  for (; item < item_end; ++item) {
    const int b = (int)(item / a.ntiles);
    const int tile = (int)(item - (long long)b * a.ntiles);
    const int smp = b / a.F;
    const int row0 = tile * R;
    const double* __restrict__ w = a.weights + (size_t)b * NW * a.nMp;

    double acc[NG * 32];
#pragma unroll
    for (int i = 0; i < NG * 32; ++i) acc[i] = 0.;

    size_t rowoff[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = (row0 + r < a.nk) ? row0 + r : a.nk - 1;   // clamp: duplicates are never written
      rowoff[r] = ((size_t)smp * a.nk + row) * rowstride;
    }

    for (int c = tid; c < ncol; c += blockDim.x) {
      const int m = c * VEC;
      double x[R][P][VEC], xu[SEP ? R : 1][SEP ? P : 1][VEC];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (VEC == 2) {
            const double2 t = hm_ldg_stream2(a.gridW[p] + rowoff[r] + m);
            x[r][p][0] = t.x; x[r][p][VEC - 1] = t.y;
            if (SEP) {
              const double2 tu = hm_ldg_stream2(a.gridU[p] + rowoff[r] + m);
              xu[SEP ? r : 0][SEP ? p : 0][0] = tu.x; xu[SEP ? r : 0][SEP ? p : 0][VEC - 1] = tu.y;
            }
          } else {
            x[r][p][0] = hm_ldg_stream1(a.gridW[p] + rowoff[r] + m);
            if (SEP) xu[SEP ? r : 0][SEP ? p : 0][0] = hm_ldg_stream1(a.gridU[p] + rowoff[r] + m);
          }
        }
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double wv[VEC];
        if (VEC == 2) {
          const double2 t = *reinterpret_cast<const double2*>(w + (size_t)i * a.nMp + m);
          wv[0] = t.x; wv[VEC - 1] = t.y;
        } else {
          wv[0] = w[(size_t)i * a.nMp + m];
        }
        // which grids does weight vector i multiply?
        int u = 0, v = 0;
        if (i < P) { u = i; }
        else if (i < 2 * P) { u = v = i - P; }
        else {
          int idx = 2 * P;
#pragma unroll
          for (int uu = 0; uu < P; ++uu)
#pragma unroll
            for (int vv = uu + 1; vv < P; ++vv) {
              if (idx == i) { u = uu; v = vv; }
              ++idx;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            if (i < P) acc[r * NW + i] += wv[e] * x[r][u][e];
            else if (i < 2 * P) {
              const double xa = SEP ? xu[SEP ? r : 0][SEP ? u : 0][e] : x[r][u][e];
              acc[r * NW + i] += (wv[e] * xa) * xa;
            } else acc[r * NW + i] += (wv[e] * x[r][u][e]) * x[r][v][e];
          }
      }
    }

    // ---- reduce over the mass axis: lanes (transposing shuffle network), then warps (shared memory) ----
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      double v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = acc[g * 32 + i];
      red[warp][g * 32 + lane] = hm_warp_transpose_sum32(v, lane);
    }
    __syncthreads();
    if (tid < NG * 32) {
      double t = 0.;
      for (int wi = 0; wi < nwarps; ++wi) t += red[wi][tid];
      fin[tid] = t;
    }
    __syncthreads();

    // ---- epilogue: one thread per (row, pair)  (pyhalomodel.py:456-501) ----
    if (tid < R * S) {
      const int r = tid / S, s = tid - r * S;
      const int row = row0 + r;
      if (row < a.nk) {
        int u = 0, v = 0, cnt = 0;
#pragma unroll
        for (int uu = 0; uu < P; ++uu)
#pragma unroll
          for (int vv = uu; vv < P; ++vv) {
            if (cnt == s) { u = uu; v = vv; }
            ++cnt;
          }
        const double* sc = a.scal + (size_t)b * HM_NSCAL;
        const double rhom = sc[1];
        const double* f = fin + r * NW;
        const double Iu = a.simple_twohalo ? sc[2 + HM_MAXP + u] : f[u] * rhom;
        const double Iv = a.simple_twohalo ? sc[2 + HM_MAXP + v] : f[v] * rhom;
        const double p2h = __dmul_rn(__dadd_rn(__dmul_rn(Iu, Iv), 0.), a.Pk_lin[(size_t)smp * a.nk + row]);   // :524
        int i1 = P + u;
        if (u != v) {
          i1 = 2 * P;
          for (int uu = 0; uu < u; ++uu) i1 += P - 1 - uu;
          i1 += v - u - 1;
        }
        double p1h = f[i1] * rhom;                                                                           // :532
        if (u == v && a.discrete[u]) {                                                                       // :484-491
          if (a.correct_discrete && !a.subtract_shotnoise) p1h += sc[2 + u];
          else if (!a.correct_discrete && a.subtract_shotnoise) p1h -= sc[2 + u];
        }
        if (a.k_trunc > 0.) {                                                                                // :494-495
          const double q = a.k[row] / a.k_trunc;
          p1h *= 1. - exp(-(q * q));
        }
        double* o = a.out + ((size_t)b * S + s) * 3 * a.nk + row;
        o[0] = p2h;
        o[a.nk] = p1h;
        o[2 * (size_t)a.nk] = __dadd_rn(p2h, p1h);                                                           // :500-501
      }
    }
    // the next item's writes to red/fin happen after its own __syncthreads, which every thread reaches only
    // after finishing this epilogue
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int hm_validate(const hm_problem* pr) {
  HM_REQUIRE(pr != nullptr, "problem is NULL");
  HM_REQUIRE(pr->nk > 0 && pr->nM > 1, "need nk >= 1 and nM >= 2");
  HM_REQUIRE(pr->n_tracers >= 1 && pr->n_tracers <= HM_MAXP, "n_tracers must be in 1..8 (use the Gram path beyond)");
  HM_REQUIRE(pr->n_samples >= 1 && pr->models_per_sample >= 1, "need at least one sample and one model per sample");
  HM_REQUIRE(pr->M_dev && pr->sigma_dev && pr->Pk_lin_dev, "M, sigma and Pk_lin are required");
  HM_REQUIRE(pr->k_trunc <= 0. || pr->k_dev, "k is required when k_trunc is set");
  const long long B = (long long)pr->n_samples * pr->models_per_sample;
  HM_REQUIRE(pr->models_dev || B == 1, "models_dev is required for more than one model");
  for (int p = 0; p < pr->n_tracers; ++p) {
    const hm_tracer& t = pr->tracers[p];
    HM_REQUIRE(t.grid_dev, "tracer grid is NULL");
    HM_REQUIRE(t.grid_mode >= HM_GRID_SEPARATE && t.grid_mode <= HM_GRID_WK, "bad grid_mode");
    HM_REQUIRE(t.grid_mode != HM_GRID_SEPARATE || t.wk_dev, "HM_GRID_SEPARATE needs wk_dev");
    HM_REQUIRE(t.amplitude_dev || t.grid_mode == HM_GRID_WK, "amplitude is required unless grid_mode is HM_GRID_WK");
    HM_REQUIRE(t.normalisation_dev, "normalisation is required");
  }
  return HM_OK;
}

extern "C" size_t hm_power_spectrum_workspace(const hm_problem* pr) {
  if (!pr || pr->n_tracers < 1 || pr->n_tracers > HM_MAXP) return 0;
  return hm_make_layout(pr).total_bytes;
}

extern "C" int hm_halo_weights(const hm_problem* pr, double* lowmass_dev, void* ws, size_t ws_bytes, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  if ((rc = hm_validate(pr))) return rc;
  const hm_layout L = hm_make_layout(pr);
  if (!ws || ws_bytes < L.total_bytes || ((uintptr_t)ws & 15)) {
    hm_set_error("hm_halo_weights: workspace needs %zu bytes, 16-byte aligned (got %zu)", L.total_bytes, ws_bytes);
    return HM_ERR_WORKSPACE;
  }
  hm_weights_args a;
  a.nM = pr->nM; a.nMp = L.nMp; a.P = L.P; a.NW = L.NW; a.F = pr->models_per_sample;
  a.B = pr->n_samples * pr->models_per_sample;
  a.simple_twohalo = pr->simple_twohalo; a.correct_discrete = pr->correct_discrete;
  a.M = pr->M_dev; a.sigma = pr->sigma_dev; a.models = pr->models_dev; a.single = pr->model_host;
  for (int p = 0; p < HM_MAXP; ++p) {
    const bool on = p < L.P;
    const hm_tracer& t = pr->tracers[p];
    a.amplitude[p] = on ? (t.amplitude_dev ? t.amplitude_dev : t.grid_dev) : nullptr;
    a.amp_stride[p] = on ? (t.amplitude_dev ? (size_t)pr->nM : (size_t)pr->nk * pr->nM) : 0;
    a.variance[p] = on ? t.variance_dev : nullptr;
    a.norm[p] = on ? t.normalisation_dev : nullptr;
    a.mode[p] = on ? t.grid_mode : 0; a.mass[p] = on ? t.mass_tracer : 0; a.discrete[p] = on ? t.discrete_tracer : 0;
  }
  a.weights = (double*)ws;
  a.scal = (double*)ws + L.scal_offset;
  a.lowmass = lowmass_dev;
  hm_weights_kernel<<<a.B, 256, 0, (cudaStream_t)stream>>>(a);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

template <int P, int R>
static int hm_launch_quad(const hm_quad_args& a, bool vec2, bool sep, cudaStream_t st) {
  const int ncol = vec2 ? a.nM / 2 : a.nM;
  int threads = ((ncol + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 64) threads = 64;
  const int ctas_per_sm = 2048 / threads > 8 ? 8 : 2048 / threads;
  long long grid = (long long)hm_num_sms() * ctas_per_sm;
  if (grid > a.nitems) grid = a.nitems;
  if (vec2) {
    if (sep) hm_quad_kernel<P, R, 2, true><<<(int)grid, threads, 0, st>>>(a);
    else hm_quad_kernel<P, R, 2, false><<<(int)grid, threads, 0, st>>>(a);
  } else {
    if (sep) hm_quad_kernel<P, R, 1, true><<<(int)grid, threads, 0, st>>>(a);
    else hm_quad_kernel<P, R, 1, false><<<(int)grid, threads, 0, st>>>(a);
  }
  HM_LAUNCH_CHECK();
  return HM_OK;
}

extern "C" int hm_mass_quadrature(const hm_problem* pr, double* out_dev, const void* ws, size_t ws_bytes,
                                  void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  if ((rc = hm_validate(pr))) return rc;
  HM_REQUIRE(out_dev, "out_dev is NULL");
  const hm_layout L = hm_make_layout(pr);
  if (!ws || ws_bytes < L.total_bytes || ((uintptr_t)ws & 15)) {
    hm_set_error("hm_mass_quadrature: workspace needs %zu bytes, 16-byte aligned (got %zu)", L.total_bytes, ws_bytes);
    return HM_ERR_WORKSPACE;
  }
  hm_quad_args a;
  a.nk = pr->nk; a.nM = pr->nM; a.nMp = L.nMp; a.F = pr->models_per_sample;
  a.B = pr->n_samples * pr->models_per_sample;
  a.simple_twohalo = pr->simple_twohalo; a.subtract_shotnoise = pr->subtract_shotnoise;
  a.correct_discrete = pr->correct_discrete; a.k_trunc = pr->k_trunc; a.k = pr->k_dev; a.Pk_lin = pr->Pk_lin_dev;
  bool sep = false, aligned = (pr->nM % 2 == 0);
  for (int p = 0; p < HM_MAXP; ++p) {
    const bool on = p < L.P;
    const hm_tracer& t = pr->tracers[p];
    a.gridU[p] = on ? t.grid_dev : nullptr;
    a.gridW[p] = on ? (t.grid_mode == HM_GRID_SEPARATE ? t.wk_dev : t.grid_dev) : nullptr;
    a.discrete[p] = on ? t.discrete_tracer : 0;
    if (on && t.grid_mode == HM_GRID_SEPARATE) sep = true;
    if (on && (((uintptr_t)a.gridU[p] | (uintptr_t)a.gridW[p]) & 15)) aligned = false;
  }
  a.weights = (const double*)ws;
  a.scal = (const double*)ws + L.scal_offset;
  a.out = out_dev;
  cudaStream_t st = (cudaStream_t)stream;
#define HM_CASE(PP, RR)                                   \
  case PP:                                                \
    a.ntiles = (pr->nk + RR - 1) / RR;                    \
    a.nitems = (long long)a.B * a.ntiles;                 \
    return hm_launch_quad<PP, RR>(a, aligned, sep, st);
  switch (L.P) {
    HM_CASE(1, 8)
    HM_CASE(2, 4)
    HM_CASE(3, 3)
    HM_CASE(4, 2)
    HM_CASE(5, 1)
    HM_CASE(6, 1)
    HM_CASE(7, 1)
    HM_CASE(8, 1)
  }
#undef HM_CASE
  hm_set_error("hm_mass_quadrature: unsupported tracer count %d", L.P);
  return HM_ERR_INVALID;
}

extern "C" int hm_power_spectrum(const hm_problem* pr, double* out_dev, double* lowmass_dev, void* ws,
                                 size_t ws_bytes, void* stream) {
  int rc = hm_halo_weights(pr, lowmass_dev, ws, ws_bytes, stream);
  if (rc) return rc;
  return hm_mass_quadrature(pr, out_dev, ws, ws_bytes, stream);
}

extern "C" int hm_low_mass_correction(const hm_model_desc* models_dev, const double* nu0_dev, int32_t B,
                                      double* A_dev, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(models_dev && nu0_dev && A_dev && B > 0, "NULL pointer or empty batch");
  hm_lowmass_kernel<<<B, 160, 0, (cudaStream_t)stream>>>(models_dev, nu0_dev, B, A_dev);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
