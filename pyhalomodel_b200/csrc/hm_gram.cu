// hm_gram.cu -- many-tracer halo-model spectra: the one-halo cross terms as a per-wavenumber Gram contraction on
// the FP64 tensor cores (BASELINE config 4: 64 HOD tracers, 2080 unique pairs per k).
//
// For P tracers the cross one-halo integrals of model.power_spectrum (pyhalomodel.py:479, :526-533) are
//     P1h_uv(k) = rhom * sum_m w1[m] * W_u(k,m) * W_v(k,m),      W_p(k,m) = c_p[m] * grid_p[k,m]
// i.e. per k the upper triangle of (w1 o W) W^T, a P x P x nM product.  With P(P+1)/2 pairs the arithmetic
// intensity (~6 flop/B at P=64) is beyond what the streaming-quadrature kernel's per-pair register accumulators
// can hold, so this path tiles it for mma.sync.m8n8k4.f64 (SASS DMMA; there is no tcgen05 FP64 MMA):
//
//   hm_gram_weights_kernel  w1, w2, per-tracer c_p (W = c_p*grid) and d_p (auto one-halo weight) per mass; A
//   hm_gram_scalars_kernel  shot noise and simple-two-halo integrals per (tracer, model)
//   hm_gram_kernel<RK>      CTA = (model, RK wavenumbers, mass part).  Per 64-mass chunk: (1) all warps load the
//                           tracer rows with 16-byte streaming loads, scale them to W, store the W tile to shared
//                           memory and accumulate the diagonal work (two-halo sums sum_m w2 W_p and auto one-halo
//                           sums sum_m d_p grid_p^2) with a transposing shuffle reduction; (2) each warp runs DMMA
//                           over its share of the 8x8 upper-triangle fragments.  Two CTAs per SM overlap (1) and (2).
//   hm_gram_epilogue_kernel adds the mass parts and finishes P_2h, P_1h, P_hm for every unique pair (:456-501)
//
// The diagonal (auto) terms are NOT Gram entries: they use Uk, amplitude, variance (pyhalomodel.py:465-477).
#include "hm_common.cuh"

#define HM_GMAXP HM_GRAM_MAX_TRACERS
#define HM_GMCH 64                 // masses per chunk
#define HM_GLD (HM_GMCH + 4)       // padded row stride of the W tile (bank-conflict-free DMMA fragment loads)
#define HM_GNSCAL 4
#define HM_GWARPS 8

struct hm_gram_layout {
  int P, NB, PT, S, NW, nMp, nchunks, nparts, MP, RK, ntiles;
  size_t B, gw_stride, scal_offset, tscal_offset, partial_offset, total_bytes;
};

static hm_gram_layout hm_gram_make_layout(const hm_gram_problem* pr) {
  hm_gram_layout L;
  L.P = pr->n_tracers;
  L.NB = (L.P + 7) / 8;
  L.PT = 8 * L.NB;
  L.S = L.P * (L.P + 1) / 2;
  L.NW = L.P + L.S;                       // [2h sums P][auto 1h P][cross 1h P(P-1)/2]
  L.nMp = (pr->nM + 1) & ~1;
  L.nchunks = (L.nMp + 255) / 256;
  L.RK = 2;
  L.ntiles = (pr->nk + L.RK - 1) / L.RK;
  L.B = (size_t)pr->n_samples * (size_t)pr->models_per_sample;
  // split the mass axis so that there are a few waves of CTAs (2 resident per SM); parts are multiples of 64 masses
  const long long rows = (long long)L.B * L.ntiles;
  int nparts = 1;
  const int maxparts = (pr->nM + 4 * HM_GMCH - 1) / (4 * HM_GMCH);       // at least 256 masses per part
  while (nparts < maxparts && rows * nparts < 6LL * 148) nparts *= 2;
  if (nparts > maxparts) nparts = maxparts;
  L.MP = (((pr->nM + nparts - 1) / nparts) + HM_GMCH - 1) / HM_GMCH * HM_GMCH;
  L.nparts = (pr->nM + L.MP - 1) / L.MP;
  auto align = [](size_t x) { return (x + 31) & ~(size_t)31; };
  L.gw_stride = (size_t)(2 + 2 * L.P) * L.nMp;
  L.scal_offset = align(L.B * L.gw_stride);
  L.tscal_offset = align(L.scal_offset + L.B * HM_GNSCAL);
  L.partial_offset = align(L.tscal_offset + L.B * (size_t)L.P * 2);
  L.total_bytes = (L.partial_offset + L.B * (size_t)L.nparts * L.NW * (size_t)pr->nk) * sizeof(double);
  return L;
}

// ---------------------------------------------------------------------------------------------------------------
struct hm_gram_wargs {
  int nM, nMp, P, F;
  int correct_discrete;
  size_t gw_stride;
  const double* M;
  const double* sigma;
  const hm_model_desc* models;
  hm_model_desc single;
  const double* amplitude[HM_GMAXP];
  int amp_stride_is_grid[HM_GMAXP];   // amplitude read from row 0 of the grid (HM_GRID_WK without amplitude)
  size_t grid_sample_stride;
  const double* variance[HM_GMAXP];
  const double* norm[HM_GMAXP];
  signed char mode[HM_GMAXP], mass[HM_GMAXP], discrete[HM_GMAXP];
  double* gw;
  double* scal;
  double* tscal;
  double* lowmass;
};

__global__ void __launch_bounds__(256) hm_gram_weights_kernel(const hm_gram_wargs a) {
  __shared__ double scratch[32];
  const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int s = b / a.F;
  const hm_model_desc d = a.models ? a.models[b] : a.single;
  const double* sig = a.sigma + (size_t)s * a.nM;
  if (chunk == 0) {   // A = 1 - int g b dnu (pyhalomodel.py:413-414)
    const double tot = hm_block_sum(hm_lowmass_node(d, d.dc / sig[0], tid), scratch);
    if (tid == 0) {
      a.scal[(size_t)b * HM_GNSCAL] = 1. - tot;
      a.scal[(size_t)b * HM_GNSCAL + 1] = d.rhom;
      if (a.lowmass) a.lowmass[b] = 1. - tot;
    }
  }
  double* gw = a.gw + (size_t)b * a.gw_stride;
  const int m = chunk * 256 + tid;
  if (m >= a.nMp) return;
  if (m >= a.nM) {
    for (int i = 0; i < 2 + 2 * a.P; ++i) gw[(size_t)i * a.nMp + m] = 0.;
    return;
  }
  const double nu = d.dc / sig[m];
  const double nu_hi = (m + 1 < a.nM) ? d.dc / sig[m + 1] : nu;
  const double nu_lo = (m > 0) ? d.dc / sig[m - 1] : nu;
  const double tw = 0.5 * (nu_hi - nu_lo);
  const double g = hm_g_nu(d, nu);
  const double bias = hm_b_nu(d, nu);
  const double Mm = a.M[m];
  const double w1 = tw * (g / Mm);             // one-halo weight (:530)
  const double w2 = tw * ((bias * g) / Mm);    // two-halo weight (:539)
  gw[m] = w1;
  gw[(size_t)a.nMp + m] = w2;
  for (int p = 0; p < a.P; ++p) {
    const size_t aoff = a.amp_stride_is_grid[p] ? (size_t)s * a.grid_sample_stride : (size_t)s * a.nM;
    const double amp = a.amplitude[p][aoff + m];
    const double norm = a.norm[p][b];
    double fac = (a.correct_discrete && a.discrete[p]) ? __dmul_rn(amp, __dadd_rn(amp, -1.)) : __dmul_rn(amp, amp);
    if (a.variance[p]) fac = __dadd_rn(fac, a.variance[p][(size_t)s * a.nM + m]);   // :466-473
    double c, uS;
    if (a.mode[p] == HM_GRID_UK) { c = amp / norm; uS = 1. / norm; }
    else { c = 1.; uS = 1. / (amp * norm); }                                        // HM_GRID_WK
    gw[(size_t)(2 + p) * a.nMp + m] = c;
    gw[(size_t)(2 + a.P + p) * a.nMp + m] = __dmul_rn(__dmul_rn(w1, fac), uS * uS);  // d_p (:475-477)
  }
}

// shot noise P_SN = rhom trapz((amp/norm^2) g/M) (:423-426) and the simple two-halo integral (:442-446)
__global__ void __launch_bounds__(256) hm_gram_scalars_kernel(const hm_gram_wargs a) {
  __shared__ double scratch[32];
  const int p = blockIdx.x, b = blockIdx.y, s = b / a.F;
  const double* gw = a.gw + (size_t)b * a.gw_stride;
  const size_t aoff = a.amp_stride_is_grid[p] ? (size_t)s * a.grid_sample_stride : (size_t)s * a.nM;
  const double norm = a.norm[p][b];
  double sn = 0., si = 0.;
  for (int m = threadIdx.x; m < a.nM; m += blockDim.x) {
    const double amp = a.amplitude[p][aoff + m];
    sn += gw[m] * (amp / (norm * norm));
    si += gw[(size_t)a.nMp + m] * (amp / norm);
  }
  sn = hm_block_sum(sn, scratch);
  si = hm_block_sum(si, scratch);
  if (threadIdx.x == 0) {
    const double A = a.scal[(size_t)b * HM_GNSCAL], rhom = a.scal[(size_t)b * HM_GNSCAL + 1];
    if (a.mass[p]) si += (A / a.M[0]) * (a.amplitude[p][aoff] / norm);
    double* t = a.tscal + ((size_t)b * a.P + p) * 2;
    t[0] = a.discrete[p] ? sn * rhom : 0.;
    t[1] = si * rhom;
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct hm_gram_args {
  int nk, nM, nMp, P, NB, PT, NW, F, ntiles, nparts, MP;
  size_t gw_stride;
  const double* grid[HM_GMAXP];
  const double* gw;
  double* partial;          // [B][nparts][NW][nk]
};

__device__ __forceinline__ void hm_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#define HM_GMAXFRAG 5     // ceil(36 / 8): upper-triangle 8x8 fragments per warp at P = 64

template <int RK>
__global__ void __launch_bounds__(32 * HM_GWARPS, 2) hm_gram_kernel(const hm_gram_args a) {
  extern __shared__ __align__(16) double gsm[];
  double* Ys = gsm;                                     // [RK][PT][HM_GLD]   W tile of the current mass chunk
  double* w1s = Ys + (size_t)RK * a.PT * HM_GLD;        // [HM_GMCH]
  double* ssum = w1s + HM_GMCH;                         // [RK][2][PT]        two-halo and auto one-halo sums
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = a.P, PT = a.PT, NB = a.NB;

  // item = (model, tile of RK wavenumbers, mass part)
  int item = blockIdx.x;
  const int part = item % a.nparts; item /= a.nparts;
  const int tile = item % a.ntiles;
  const int b = item / a.ntiles;
  const int smp = b / a.F;
  const int row0 = tile * RK;
  const int mbeg = part * a.MP;
  const int mend = (mbeg + a.MP < a.nM) ? mbeg + a.MP : a.nM;
  const double* __restrict__ gw = a.gw + (size_t)b * a.gw_stride;

  // my share of the upper-triangle fragments (i <= j), contiguous in row-major order so that A is reused
  const int nfrag = NB * (NB + 1) / 2;
  const int per = (nfrag + HM_GWARPS - 1) / HM_GWARPS;
  int fi[HM_GMAXFRAG], fj[HM_GMAXFRAG], nmine = 0;
  {
    int f = 0;
    for (int i = 0; i < NB; ++i)
      for (int j = i; j < NB; ++j, ++f)
        if (f >= warp * per && f < (warp + 1) * per && nmine < HM_GMAXFRAG) {
#pragma unroll
          for (int t = 0; t < HM_GMAXFRAG; ++t)
            if (t == nmine) { fi[t] = i; fj[t] = j; }
          ++nmine;
        }
  }
  double acc[RK][HM_GMAXFRAG][2];
#pragma unroll
  for (int r = 0; r < RK; ++r)
#pragma unroll
    for (int t = 0; t < HM_GMAXFRAG; ++t) acc[r][t][0] = acc[r][t][1] = 0.;
  for (int i = tid; i < RK * 2 * PT; i += blockDim.x) ssum[i] = 0.;

  size_t rowoff[RK];
#pragma unroll
  for (int r = 0; r < RK; ++r) {
    const int row = (row0 + r < a.nk) ? row0 + r : a.nk - 1;     // clamp: duplicates are never written
    rowoff[r] = ((size_t)smp * a.nk + row) * (size_t)a.nM;
  }

  for (int m0 = mbeg; m0 < mend; m0 += HM_GMCH) {
    // ---- phase 1: load, scale to W, stage, and the diagonal work.  Warp w owns tracers w, w+8, ... ----------
    const int m = m0 + 2 * lane;
    const bool in = m < mend;                          // nM and MP are even: a lane's two masses are in or out together
    const double2 w2v = in ? *reinterpret_cast<const double2*>(gw + (size_t)a.nMp + m) : make_double2(0., 0.);
    if (warp == 0) {
      const double2 w1v = in ? *reinterpret_cast<const double2*>(gw + m) : make_double2(0., 0.);
      *reinterpret_cast<double2*>(w1s + 2 * lane) = w1v;
    }
    double vals[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) vals[i] = 0.;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int p = warp + 8 * t;
      if (p < PT) {                                    // warp-uniform
        double2 cc = make_double2(0., 0.), dd = make_double2(0., 0.);
        double2 x[RK];
#pragma unroll
        for (int r = 0; r < RK; ++r) x[r] = make_double2(0., 0.);
        if (p < P && in) {
          cc = *reinterpret_cast<const double2*>(gw + (size_t)(2 + p) * a.nMp + m);
          dd = *reinterpret_cast<const double2*>(gw + (size_t)(2 + P + p) * a.nMp + m);
#pragma unroll
          for (int r = 0; r < RK; ++r) x[r] = hm_ldg_stream2(a.grid[p] + rowoff[r] + m);
        }
#pragma unroll
        for (int r = 0; r < RK; ++r) {
          const double2 y = make_double2(cc.x * x[r].x, cc.y * x[r].y);
          *reinterpret_cast<double2*>(Ys + ((size_t)r * PT + p) * HM_GLD + 2 * lane) = y;
          if (RK * 2 * t + 2 * r + 1 < 32) {
            vals[(t * RK + r) * 2] = fma(w2v.x, y.x, w2v.y * y.y);                           // sum_m w2 W_p   (:539)
            vals[(t * RK + r) * 2 + 1] = fma(dd.x * x[r].x, x[r].x, (dd.y * x[r].y) * x[r].y);   // sum_m d_p U^2  (:475-477)
          }
        }
      }
    }
    {
      const double tot = hm_warp_transpose_sum32(vals, lane);     // lane j now holds value j of this warp
      const int t = lane / (2 * RK), r = (lane >> 1) % RK, kind = lane & 1;
      const int p = warp + 8 * t;
      if (lane < 16 * RK && p < PT) ssum[((size_t)r * 2 + kind) * PT + p] += tot;   // only this warp touches these p
    }
    __syncthreads();

    // ---- phase 2: DMMA over my fragments: C[u,v] += sum_m (w1 W_u) W_v ---------------------------------------
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll 4
    for (int s4 = 0; s4 < HM_GMCH / 4; ++s4) {
      const double w1e = w1s[4 * s4 + fk];
#pragma unroll
      for (int r = 0; r < RK; ++r) {
        const double* Yr = Ys + (size_t)r * PT * HM_GLD + 4 * s4 + fk;
        int last_i = -1;
        double afrag = 0.;
#pragma unroll
        for (int t = 0; t < HM_GMAXFRAG; ++t) {
          if (t < nmine) {                             // warp-uniform
            if (fi[t] != last_i) {
              last_i = fi[t];
              afrag = w1e * Yr[(size_t)(8 * fi[t] + fr) * HM_GLD];
            }
            const double bfrag = Yr[(size_t)(8 * fj[t] + fr) * HM_GLD];
            hm_dmma884(acc[r][t][0], acc[r][t][1], afrag, bfrag);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- write this part's sums: partial[b][part][value][row] -----------------------------------------------------
  double* dst = a.partial + ((size_t)b * a.nparts + part) * (size_t)a.NW * a.nk;
  for (int i = tid; i < RK * 2 * P; i += blockDim.x) {
    const int r = i / (2 * P), rem = i - r * 2 * P, kind = rem / P, p = rem - kind * P;
    if (row0 + r < a.nk) dst[(size_t)(kind * P + p) * a.nk + row0 + r] = ssum[((size_t)r * 2 + kind) * PT + p];
  }
  // C fragment layout (m8n8k4.f64): c0/c1 = C[8i + lane/4][8j + 2*(lane%4) + {0,1}]
#pragma unroll
  for (int t = 0; t < HM_GMAXFRAG; ++t) {
    if (t < nmine) {
      const int u = 8 * fi[t] + (lane >> 2);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int v = 8 * fj[t] + 2 * (lane & 3) + e;
        if (u < v && v < P) {
          // strict upper triangle, row-major: index of (u, v)
          const size_t idx = (size_t)2 * P + (size_t)u * P - (size_t)u * (u + 1) / 2 + (v - u - 1);
#pragma unroll
          for (int r = 0; r < RK; ++r)
            if (row0 + r < a.nk) dst[idx * a.nk + row0 + r] = acc[r][t][e];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct hm_gram_eargs {
  int nk, nM, nMp, P, S, NW, F, nparts, B;
  int simple_twohalo, subtract_shotnoise, correct_discrete;
  double k_trunc;
  size_t gw_stride;
  const double* k;
  const double* M;
  const double* Pk_lin;
  const double* partial;
  const double* gw;
  const double* scal;
  const double* tscal;
  const double* grid[HM_GMAXP];
  signed char mass[HM_GMAXP], discrete[HM_GMAXP];
  double* out;              // [B][S][3][nk]
};

// thread per (model, unique pair, wavenumber), wavenumber fastest => coalesced reads of the partial sums and writes
__global__ void __launch_bounds__(256) hm_gram_epilogue_kernel(const hm_gram_eargs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_model = (long long)a.S * a.nk;
  if (gid >= per_model * a.B) return;
  const int b = (int)(gid / per_model);
  const long long rem = gid - (long long)b * per_model;
  const int s = (int)(rem / a.nk), row = (int)(rem - (long long)s * a.nk);
  const int smp = b / a.F, P = a.P;
  // pair s -> (u, v), u <= v, row-major
  int u = 0, left = s;
  while (left >= P - u) { left -= P - u; ++u; }
  const int v = u + left;
  const double A = a.scal[(size_t)b * HM_GNSCAL], rhom = a.scal[(size_t)b * HM_GNSCAL + 1];
  const size_t i1h = (u == v) ? (size_t)(P + u)
                              : (size_t)2 * P + (size_t)u * P - (size_t)u * (u + 1) / 2 + (v - u - 1);
  double fu = 0., fv = 0., f1 = 0.;
  for (int j = 0; j < a.nparts; ++j) {
    const double* pp = a.partial + ((size_t)b * a.nparts + j) * (size_t)a.NW * a.nk + row;
    fu += pp[(size_t)u * a.nk];
    fv += pp[(size_t)v * a.nk];
    f1 += pp[i1h * a.nk];
  }
  const double* gw = a.gw + (size_t)b * a.gw_stride;
  const double Aterm = A / a.M[0];
  const size_t goff = ((size_t)smp * a.nk + row) * (size_t)a.nM;     // W_p(k, M[0]) = c_p[0] * grid_p[k, 0]  (:541-542)
  if (a.mass[u]) fu += Aterm * (gw[(size_t)(2 + u) * a.nMp] * a.grid[u][goff]);
  if (a.mass[v]) fv += Aterm * (gw[(size_t)(2 + v) * a.nMp] * a.grid[v][goff]);
  const double Iu = a.simple_twohalo ? a.tscal[((size_t)b * P + u) * 2 + 1] : fu * rhom;
  const double Iv = a.simple_twohalo ? a.tscal[((size_t)b * P + v) * 2 + 1] : fv * rhom;
  const double p2h = __dmul_rn(__dadd_rn(__dmul_rn(Iu, Iv), 0.), a.Pk_lin[(size_t)smp * a.nk + row]);   // :524
  double p1h = f1 * rhom;                                                                              // :532
  if (u == v && a.discrete[u]) {                                                                       // :484-491
    const double sn = a.tscal[((size_t)b * P + u) * 2];
    if (a.correct_discrete && !a.subtract_shotnoise) p1h += sn;
    else if (!a.correct_discrete && a.subtract_shotnoise) p1h -= sn;
  }
  if (a.k_trunc > 0.) {                                                                                // :494-495
    const double q = a.k[row] / a.k_trunc;
    p1h *= 1. - exp(-(q * q));
  }
  double* o = a.out + ((size_t)b * a.S + s) * 3 * a.nk + row;
  o[0] = p2h;
  o[a.nk] = p1h;
  o[2 * (size_t)a.nk] = __dadd_rn(p2h, p1h);                                                           // :500-501
}

// ---------------------------------------------------------------------------------------------------------------
static int hm_gram_validate(const hm_gram_problem* pr) {
  HM_REQUIRE(pr != nullptr, "problem is NULL");
  HM_REQUIRE(pr->nk > 0 && pr->nM > 1, "need nk >= 1 and nM >= 2");
  HM_REQUIRE(pr->nM % 2 == 0, "the Gram path needs an even number of masses");
  HM_REQUIRE(pr->n_tracers >= 2 && pr->n_tracers <= HM_GMAXP, "n_tracers must be in 2..64");
  HM_REQUIRE(pr->n_samples >= 1 && pr->models_per_sample >= 1, "need at least one sample and one model per sample");
  HM_REQUIRE(pr->M_dev && pr->sigma_dev && pr->Pk_lin_dev && pr->tracers, "M, sigma, Pk_lin and tracers are required");
  HM_REQUIRE(pr->k_trunc <= 0. || pr->k_dev, "k is required when k_trunc is set");
  const long long B = (long long)pr->n_samples * pr->models_per_sample;
  HM_REQUIRE(B <= 65535, "at most 65535 models per call");
  HM_REQUIRE(pr->models_dev || B == 1, "models_dev is required for more than one model");
  for (int p = 0; p < pr->n_tracers; ++p) {
    const hm_tracer& t = pr->tracers[p];
    HM_REQUIRE(t.grid_dev && !((uintptr_t)t.grid_dev & 15), "tracer grid is NULL or not 16-byte aligned");
    HM_REQUIRE(t.grid_mode == HM_GRID_UK || t.grid_mode == HM_GRID_WK, "the Gram path needs single-grid tracers");
    HM_REQUIRE(t.amplitude_dev || t.grid_mode == HM_GRID_WK, "amplitude is required unless grid_mode is HM_GRID_WK");
    HM_REQUIRE(t.normalisation_dev, "normalisation is required");
  }
  return HM_OK;
}

extern "C" size_t hm_gram_workspace(const hm_gram_problem* pr) {
  if (!pr || pr->n_tracers < 2 || pr->n_tracers > HM_GMAXP) return 0;
  return hm_gram_make_layout(pr).total_bytes;
}

extern "C" int hm_gram_power_spectrum(const hm_gram_problem* pr, double* out_dev, double* lowmass_dev, void* ws,
                                      size_t ws_bytes, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  if ((rc = hm_gram_validate(pr))) return rc;
  HM_REQUIRE(out_dev, "out_dev is NULL");
  const hm_gram_layout L = hm_gram_make_layout(pr);
  if (!ws || ws_bytes < L.total_bytes || ((uintptr_t)ws & 15)) {
    hm_set_error("hm_gram_power_spectrum: workspace needs %zu bytes, 16-byte aligned (got %zu)", L.total_bytes, ws_bytes);
    return HM_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  double* wsd = (double*)ws;

  hm_gram_wargs w;
  w.nM = pr->nM; w.nMp = L.nMp; w.P = L.P; w.F = pr->models_per_sample; w.correct_discrete = pr->correct_discrete;
  w.gw_stride = L.gw_stride; w.M = pr->M_dev; w.sigma = pr->sigma_dev; w.models = pr->models_dev;
  w.single = pr->model_host; w.grid_sample_stride = (size_t)pr->nk * pr->nM;
  for (int p = 0; p < HM_GMAXP; ++p) {
    const bool on = p < L.P;
    const hm_tracer* t = on ? &pr->tracers[p] : nullptr;
    w.amplitude[p] = on ? (t->amplitude_dev ? t->amplitude_dev : t->grid_dev) : nullptr;
    w.amp_stride_is_grid[p] = on ? (t->amplitude_dev ? 0 : 1) : 0;
    w.variance[p] = on ? t->variance_dev : nullptr;
    w.norm[p] = on ? t->normalisation_dev : nullptr;
    w.mode[p] = on ? (signed char)t->grid_mode : 0;
    w.mass[p] = on ? (signed char)(t->mass_tracer != 0) : 0;
    w.discrete[p] = on ? (signed char)(t->discrete_tracer != 0) : 0;
  }
  w.gw = wsd; w.scal = wsd + L.scal_offset; w.tscal = wsd + L.tscal_offset; w.lowmass = lowmass_dev;
  hm_gram_weights_kernel<<<dim3(L.nchunks, (unsigned)L.B), 256, 0, st>>>(w);
  HM_LAUNCH_CHECK();
  hm_gram_scalars_kernel<<<dim3(L.P, (unsigned)L.B), 256, 0, st>>>(w);
  HM_LAUNCH_CHECK();

  hm_gram_args g;
  g.nk = pr->nk; g.nM = pr->nM; g.nMp = L.nMp; g.P = L.P; g.NB = L.NB; g.PT = L.PT; g.NW = L.NW;
  g.F = pr->models_per_sample; g.ntiles = L.ntiles; g.nparts = L.nparts; g.MP = L.MP; g.gw_stride = L.gw_stride;
  for (int p = 0; p < HM_GMAXP; ++p) g.grid[p] = (p < L.P) ? pr->tracers[p].grid_dev : nullptr;
  g.gw = wsd; g.partial = wsd + L.partial_offset;
  const size_t smem = ((size_t)L.RK * L.PT * HM_GLD + HM_GMCH + (size_t)L.RK * 2 * L.PT) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gram_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  const long long items = (long long)L.B * L.ntiles * L.nparts;
  HM_REQUIRE(items < (1LL << 31), "too many work items for one launch");
  hm_gram_kernel<2><<<(unsigned)items, 32 * HM_GWARPS, smem, st>>>(g);
  HM_LAUNCH_CHECK();

  hm_gram_eargs e;
  e.nk = pr->nk; e.nM = pr->nM; e.nMp = L.nMp; e.P = L.P; e.S = L.S; e.NW = L.NW; e.F = pr->models_per_sample;
  e.nparts = L.nparts; e.B = (int)L.B;
  e.simple_twohalo = pr->simple_twohalo; e.subtract_shotnoise = pr->subtract_shotnoise;
  e.correct_discrete = pr->correct_discrete; e.k_trunc = pr->k_trunc; e.gw_stride = L.gw_stride;
  e.k = pr->k_dev; e.M = pr->M_dev; e.Pk_lin = pr->Pk_lin_dev; e.partial = g.partial; e.gw = wsd;
  e.scal = w.scal; e.tscal = w.tscal;
  for (int p = 0; p < HM_GMAXP; ++p) {
    e.grid[p] = g.grid[p];
    e.mass[p] = w.mass[p];
    e.discrete[p] = w.discrete[p];
  }
  e.out = out_dev;
  const long long nthreads = (long long)L.B * L.S * pr->nk;
  hm_gram_epilogue_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(e);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
