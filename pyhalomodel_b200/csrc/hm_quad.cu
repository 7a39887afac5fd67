// hm_quad.cu -- the hot path of model.power_spectrum (pyhalomodel.py:361-513) as a short kernel pipeline:
//
//   hm_weights_kernel      (256-mass chunk, model) CTAs: nu = dc/sigma, g(nu), b(nu), trapezoid-in-nu weights, the
//                          low-mass correction A, and from those the P+S per-mass weight vectors that turn every
//                          integral of the reference into a plain weighted row sum over M:
//                            I_u(k)  = rhom * sum_m  e_u[m]  * G_u[k,m]                  (:535-544, incl. A W(k,M0)/M0)
//                            P1h_uu  = rhom * sum_m  d_u[m]  * G_u[k,m]^2                (:465-477, :526-533)
//                            P1h_uv  = rhom * sum_m  q_uv[m] * G_u[k,m] * G_v[k,m]       (:479)
//                          where G_u is the single grid a Fourier profile was built from (Uk, or Wk when
//                          amplitude=None; :874-881), so each (k,M) grid is read from HBM exactly once.
//   hm_quad_stream_kernel  the HBM-bound kernel.  Persistent CTAs (one per SM); a producer warp feeds a shared-
//                          memory ring (221 KB: 6 stages, or 3 of twice the width) with TMA bulk copies
//                          (cp.async.bulk + mbarrier complete_tx) of R-row x 512*H-mass tiles of every tracer
//                          grid; 8 independent consumer warps keep the weights of their 2*H mass columns in
//                          registers, accumulate R*(P+S) sums, reduce them over the warp with a transposing
//                          shuffle network and store the warp's partial row sums (no CTA-wide synchronisation).
//   hm_quad_rows_kernel    sweeps with two to five models per sample on one to three tracers: a lane owns two
//                          wavenumber rows (one at three tracers) for the whole mass axis, the weights are broadcast
//                          from shared memory, and nothing is reduced across lanes (see the comment above the kernel).
//   hm_quad_kernel         fallback of the same arithmetic with plain (scalar or 16-byte) loads for odd nM,
//                          unaligned grids or profiles with separate Uk/Wk grids.
//   hm_beta2h_kernel       optional non-linear halo bias: I_NL = rhom^2 a_u^T beta_k a_v + edge terms (:546-571)
//   hm_epilogue_kernel     adds the mass parts in fixed order and forms P_2h = (I_u I_v + I_NL) P_lin, the
//                          discrete/shot-noise/k_trunc fix-ups and P_hm (:456-501).
//
// Algorithmic bytes per launch of the streaming kernel: 8*(G*P*nk*nM + B*(P+S)*nM).
#include <stdlib.h>

#include <cuda.h>      // CUtensorMap: types only, the encoder is fetched through cudaGetDriverEntryPoint

#include "hm_common.cuh"

#define HM_MAXP HM_MAX_TRACERS

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
// Workspace layout (doubles):
//   weights [B][NW][nMp]            per-mass weight vectors
//   scal    [B][HM_NSCAL]           A, rhom, w2[0], -, c_p[0] (8): the last two feed the beta_NL edge terms; then the
//                                   per-(model, tracer) factors of the row-per-lane kernel (see hm_weights_kernel)
//   inl     [B][S][nk]              non-linear-bias two-halo term I_NL (only when beta is given)
//   cpart   [B][nchunks+1][2*HM_MAXP] per-mass-chunk partial sums of the shot-noise and simple-two-halo integrals
//                                   (last slot: the low-mass term of the latter); only written when flags need them
//   partial [B][nparts][nk][NW]     partial row sums over mass sub-ranges written by the quadrature kernel
//                                   (streaming kernel: one part per consumer warp per 512-mass slice)
// ---------------------------------------------------------------------------------------------------------------
#define HM_NSCAL 24          // per model: A, rhom, w2[0], -, c_p[0] (8); row-per-lane sweeps: s_p (3), rn_p^2 (3), low-mass term (3), - (3)
#define HM_RSCAL 12          // first of the row-per-lane kernel's per-(model, tracer) scalars
#define HM_WCHUNK 256        // masses per CTA of the weights kernel
#define HM_MC 512            // masses per slice of the streaming kernel per pair of columns a consumer thread owns
                             // (256 consumer threads x double2); slices are HM_MC or 2*HM_MC wide
#define HM_CONSUMER_WARPS 8
#define HM_STREAM_THREADS (32 * (HM_CONSUMER_WARPS + 1))
#define HM_ROWS_CS 32        // row-per-lane kernel: mass columns per stage (four per consumer warp)
#define HM_ROWS_THREADS (32 * (HM_CONSUMER_WARPS + 4))   // two consumer warpgroups + the producer's warpgroup (setmaxnreg)
#define HM_ROWS_PRODUCER_REGS 24
#ifndef HM_ROWS_LAUNCH_THREADS
#define HM_ROWS_LAUNCH_THREADS HM_ROWS_THREADS   // (experiment: 512 makes ptxas launch the CTA with 128 registers per thread,
#endif                                           //  which leaves room for a co-resident weights CTA)
#ifndef HM_ROWS_CONSUMER_REGS
#define HM_ROWS_CONSUMER_REGS 232
#endif
#define HM_ROWS_MAXF 5       // ... for 2..5 models per sample on one to three tracers
#define HM_ROWS_MAXP 3

struct hm_layout {
  int P, S, NW, nMp, nchunks, nslices, nparts, MC;
  int NVEC;                  // row-per-lane sweeps: factored weight vectors per SAMPLE (2 F + 2 P)
  bool stream, rows;
  size_t B, scal_offset, cpart_offset, partial_offset, inl_offset, bpart_offset, bcount_offset, counter_offset, total_bytes;
  int bsplit;                // beta_NL: row chunks per (model, wavenumber) matrix, one CTA each
};

static bool hm_can_stream(const hm_problem* pr) {
  if (pr->nM % 2) return false;
  for (int p = 0; p < pr->n_tracers; ++p) {
    if (pr->tracers[p].grid_mode == HM_GRID_SEPARATE) return false;
    if ((uintptr_t)pr->tracers[p].grid_dev & 15) return false;
  }
  return true;
}

// The shot-noise integral only enters a discrete auto spectrum when exactly one of correct_discrete /
// subtract_shotnoise is set (:484-491), the k-independent two-halo integral only with simple_twohalo (:442-446).
static bool hm_needs_cpart(const hm_problem* pr) {
  return pr->simple_twohalo || ((pr->correct_discrete != 0) != (pr->subtract_shotnoise != 0));
}

static hm_layout hm_make_layout(const hm_problem* pr) {
  hm_layout L;
  L.P = pr->n_tracers;
  L.S = L.P * (L.P + 1) / 2;
  L.NW = L.P + L.S;
  L.nMp = (pr->nM + 1) & ~1;
  L.nchunks = (L.nMp + HM_WCHUNK - 1) / HM_WCHUNK;
  L.stream = hm_can_stream(pr);
  // wide slices (4 mass columns per consumer thread) halve the partial sums and the reductions per byte; they fit
  // the register budget up to 3 tracers and are used when the mass axis is long enough to fill them
  // two to five mass functions per sample on one to three tracers (BASELINE config 3 has five on two): the row-per-lane
  // kernel reads the shared grids once for all of them and combines its column groups itself (one part)
  L.rows = L.stream && pr->models_per_sample >= 2 && pr->models_per_sample <= HM_ROWS_MAXF && L.P <= HM_ROWS_MAXP && !pr->beta_dev;
  L.NVEC = 2 * pr->models_per_sample + 2 * L.P;
  L.MC = (pr->nM >= 2 * HM_MC && L.P <= 3 && !L.rows) ? 2 * HM_MC : HM_MC;
  L.nslices = L.stream ? (pr->nM + L.MC - 1) / L.MC : 1;
  L.nparts = L.stream ? L.nslices * HM_CONSUMER_WARPS : 1;
  if (L.rows) { L.nslices = 1; L.nparts = 1; }
  L.B = (size_t)pr->n_samples * (size_t)pr->models_per_sample;
  auto align = [](size_t x) { return (x + 31) & ~(size_t)31; };
  const size_t wvecs = L.rows ? (size_t)pr->n_samples * L.NVEC : L.B * (size_t)L.NW;
  L.scal_offset = align(wvecs * (size_t)L.nMp);
  L.cpart_offset = align(L.scal_offset + L.B * HM_NSCAL);
  L.partial_offset = align(L.cpart_offset + L.B * (size_t)(L.nchunks + 1) * 2 * HM_MAXP);
  L.inl_offset = align(L.partial_offset + L.B * (size_t)L.nparts * L.NW * (size_t)pr->nk);
  // beta_NL: a (model, k) matrix is one CTA's worth of streaming (8 MB at 1024 masses), and B nk CTAs seldom divide evenly
  // over the SMs (256 over 148: the SMs with two take twice as long, 0.81 of the HBM peak); cut every matrix into row chunks
  // so that there are at least sixteen CTAs per SM (measured: 449 / 377 / 354 / 337 / 356 us at 3 / 6 / 9 / 16 / 24 for
  // 256 x 1024^2), each with its own partial sums, finished by the last chunk to arrive
  L.bsplit = 1;
  if (pr->beta_dev) {
    const long long mats = (long long)L.B * pr->nk;
    static const char* env = getenv("HALOB200_BETA_CTAS_PER_SM");      // (A/B switch for timing runs)
    const long long per_sm = env ? atoll(env) : 16;
    long long sp = (per_sm * hm_num_sms() + mats - 1) / mats;
    const long long most = (pr->nM + 63) / 64;                     // at least 64 rows per chunk
    sp = sp > most ? most : sp;
    L.bsplit = (int)(sp < 1 ? 1 : (sp > 16 ? 16 : sp));
  }
  L.bpart_offset = align(L.inl_offset + (pr->beta_dev ? L.B * (size_t)L.S * pr->nk : 0));
  L.bcount_offset = align(L.bpart_offset + (pr->beta_dev ? L.B * (size_t)pr->nk * L.bsplit * (L.S + 2 * L.P) : 0));
  L.counter_offset = align(L.bcount_offset + (pr->beta_dev ? (L.B * (size_t)pr->nk + 1) / 2 : 0));      // sweep kernel: item queue
  L.total_bytes = (L.counter_offset + 32) * sizeof(double);
  return L;
}

struct hm_weights_args {
  int nM, nMp, P, NW, F, nchunks;
  int rows, NVEC;                    // factored layout of the row-per-lane sweep kernel: NVEC vectors per sample
  int correct_discrete;
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
  double* cpart;                     // [B][nchunks + 1][2*HM_MAXP], written only by hm_weights_kernel<true>
  double* lowmass;                   // optional [B]
};

// grid (nchunks + 1, G): CTA (chunk, s) builds the weight vectors of 256 masses for ALL F models of sample s -- the
// loads of sigma, M, amplitudes and variances, the three divisions behind nu and ln(nu) are shared by the F mass
// functions (hm_gb_nu_shared) -- and CTA (nchunks, s) integrates the low-mass correction A of each model
// (pyhalomodel.py:413-414) and owns everything that needs it: e_p[0] with the A W(k,M0)/M0 term, the scalars, A.
// Factored layout (a.rows; the row-per-lane sweep kernel): every weight vector of the reference is a product of a
// mass-function part, a tracer part and a per-(model, tracer) scalar,
//   e_p^f = w2^f * A_p * s_p^f,   d_p^f = w1^f * T_p * (rn_p^f)^2,   q_uv^f = w1^f * A_u A_v * s_u^f s_v^f,
// with A_p = amplitude (1 for HM_GRID_WK), T_p = <N(N-1)> etc. (divided by amplitude^2 for HM_GRID_WK), rn = 1 /
// normalisation and s_p = rn (1 for HM_GRID_WK).  The kernel then writes 2 F + 2 P vectors per SAMPLE -- w2^f, w1^f,
// A_p, T_p -- instead of F (P + S) per sample, and the scalars and the low-mass term A W(k, M0) / M0 (:541-542) go to
// scal[HM_RSCAL ...]: the sweep kernel applies them once per row at the end of the mass axis.
// RED: also reduce the shot-noise and simple-two-halo integrals (:423-426, :442-446); only non-default flags read
// them (hm_epilogue_kernel), so the default path skips the 16 block reductions per model altogether.
// PT: compile-time bound on the number of tracers (2, 4 or 8).  The kernel is bound by the latency of its FP64
// chains (exp, log, division), so what counts is resident warps: PT = 2 fits 64 registers, 32 warps per SM.
template <int PT>
struct hm_fam_consts { hm_model_desc d; hm_family_consts c; double rnorm[PT]; };

template <bool RED, int PT, bool ROWS>
#ifndef HM_WMINB
#define HM_WMINB 4
#endif
__global__ void __launch_bounds__(HM_WCHUNK, (PT <= 2) ? HM_WMINB : (PT <= 4 ? 3 : 2)) hm_weights_kernel(const hm_weights_args a) {
  constexpr int NRED = 2 * PT;                          // shot noise [PT], simple two-halo [PT]
  constexpr int FG = 8;                                 // models per group (shared constants, low-mass nodes)
  __shared__ double sred[RED ? HM_WCHUNK / 32 : 1][RED ? NRED : 1];
  __shared__ hm_fam_consts<PT> sfc[FG];
  const int s = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* sig = a.sigma + (size_t)s * a.nM;
  const double M0 = a.M[0];

  if (chunk == a.nchunks) {
    // ---- low-mass CTA: A = 1 - int g b dnu from nu[0] by composite Gauss-Legendre.  Warp w integrates model w, w + 8,
    // ...: a lane evaluates five of the 160 nodes (independent chains), the warp adds them in a fixed order, and lane 0
    // finishes the model's column m = 0. ----
    const double sg0 = sig[0], sg1 = sig[1];
    for (int f = warp; f < a.F; f += HM_WCHUNK / 32) {
      const int b = s * a.F + f;
      const hm_model_desc d = a.models ? a.models[b] : a.single;
      const hm_family_consts c = hm_make_family_consts(d);
      const double nu = d.dc / sg0;
      double part = 0.;
#pragma unroll 1
      for (int j = 0; j < HM_GL_NODES / 32; ++j) part += hm_lowmass_node_shared(d, c, nu, lane + 32 * j);
      const double A = 1. - hm_warp_sum(part);
      if (lane == 0) {
        const double Aterm = A / M0;                    // multiplies W(k, M[0]) for mass tracers (:541-542)
        const double nu_hi = d.dc / sg1;
        const double tw = 0.5 * (nu_hi - nu);
        double g, bg;
        hm_gb_nu_shared(d, c, nu, log(nu), g, bg);      // the same arithmetic as the other masses' weights
        const double w2_plain = tw * (bg * (1. / M0));
        double* wout = a.weights + (size_t)b * a.NW * a.nMp;
        double* sc = a.scal + (size_t)b * HM_NSCAL;
        double* cp = a.cpart + ((size_t)b * (a.nchunks + 1) + a.nchunks) * 2 * HM_MAXP;
        for (int p = 0; p < HM_MAXP; ++p) {
          double cW = 0., low = 0.;
          if (p < a.P) {
            const double an = a.amplitude[p][(size_t)s * a.amp_stride[p]] * (1. / a.norm[p][b]);
            cW = (a.mode[p] == HM_GRID_UK) ? an : 1.;
            if (!ROWS) wout[(size_t)p * a.nMp] = a.mass[p] ? (w2_plain + Aterm) * cW : w2_plain * cW;   // e_p[0]
            if (a.mass[p]) low = Aterm * an;
            if (ROWS && p < HM_ROWS_MAXP) {
              const double rn = 1. / a.norm[p][b];
              sc[HM_RSCAL + p] = (a.mode[p] == HM_GRID_UK) ? rn : 1.;       // s_p
              sc[HM_RSCAL + HM_ROWS_MAXP + p] = rn * rn;                               // rn_p^2
              sc[HM_RSCAL + 2 * HM_ROWS_MAXP + p] = a.mass[p] ? Aterm * cW : 0.;           // multiplies G_p(k, M0) in I_p
            }
          }
          sc[4 + p] = cW;                               // w2[0] and c_p[0]: edge terms of the beta_NL integral
          if (RED) { cp[p] = 0.; cp[HM_MAXP + p] = low; }
        }
        sc[0] = A;
        sc[1] = d.rhom;
        sc[2] = w2_plain;
        if (a.lowmass) a.lowmass[b] = A;
      }
    }
    return;
  }

  const int m = chunk * HM_WCHUNK + tid;
  const bool live = m < a.nM;
  const bool pad = !live && m < a.nMp;                  // padding column: zero weights
  double sg = 1., sg_hi = 1., sg_lo = 1., rM = 1.;
  double amp[PT], var[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) amp[p] = var[p] = 0.;
  if (live) {
    sg = sig[m];
    sg_hi = (m + 1 < a.nM) ? sig[m + 1] : sg;
    sg_lo = (m > 0) ? sig[m - 1] : sg;
    rM = 1. / a.M[m];
#pragma unroll
    for (int p = 0; p < PT; ++p)
      if (p < a.P) {
        amp[p] = a.amplitude[p][(size_t)s * a.amp_stride[p] + m];
        if (a.variance[p]) var[p] = a.variance[p][(size_t)s * a.nM + m];
      }
  }
  if (ROWS && (live || pad)) {
    // family-independent tracer vectors A_p, T_p of this sample
    double* wv = a.weights + ((size_t)s * a.NVEC + 2 * a.F) * a.nMp + m;
#pragma unroll
    for (int p = 0; p < PT; ++p)
      if (p < a.P) {
        double fac = (a.correct_discrete && a.discrete[p]) ? __dmul_rn(amp[p], __dadd_rn(amp[p], -1.))
                                                           : __dmul_rn(amp[p], amp[p]);
        if (a.variance[p]) fac = __dadd_rn(fac, var[p]);
        double A = amp[p], T = fac;
        if (a.mode[p] != HM_GRID_UK) {
          const double ra = 1. / amp[p];
          A = 1.; T = __dmul_rn(fac, ra * ra);
        }
        wv[(size_t)p * a.nMp] = live ? A : 0.;
        wv[(size_t)(a.P + p) * a.nMp] = live ? T : 0.;
      }
  }
  double dc_prev = 0., nu = 0., lnnu = 0., tw = 0.;
#pragma unroll 1
  for (int f = 0; f < a.F; ++f) {
    const int b = s * a.F + f;
    if ((f & (FG - 1)) == 0) {
      // per-model constants of the next group of models: parameter logarithms and reciprocal normalisations,
      // computed once per CTA instead of once per thread
      __syncthreads();                                  // the previous group is done with sfc
      if (tid < FG && f + tid < a.F) {
        const hm_model_desc dm = a.models ? a.models[b + tid] : a.single;
        sfc[tid].d = dm;
        sfc[tid].c = hm_make_family_consts(dm);
#pragma unroll
        for (int p = 0; p < PT; ++p) sfc[tid].rnorm[p] = (p < a.P) ? 1. / a.norm[p][b + tid] : 0.;
      }
      __syncthreads();
    }
    const hm_fam_consts<PT>& fc = sfc[f & (FG - 1)];
    double* wout = a.weights + (size_t)b * a.NW * a.nMp;
    double red[RED ? NRED : 1];
#pragma unroll
    for (int i = 0; i < (RED ? NRED : 1); ++i) red[i] = 0.;
    double* wrow = a.weights + ((size_t)s * a.NVEC + f) * a.nMp + m;     // factored layout: w2^f, and w1^f F vectors on
    if (pad) {
      if (ROWS) { wrow[0] = 0.; wrow[(size_t)a.F * a.nMp] = 0.; }
      else
        for (int i = 0; i < a.NW; ++i) wout[(size_t)i * a.nMp + m] = 0.;
    } else if (live) {
      const hm_model_desc& d = fc.d;
      const double dc = d.dc;
      if (f == 0 || dc != dc_prev) {                    // models of one sample usually share delta_c
        nu = dc / sg;
        const double nu_hi = (m + 1 < a.nM) ? dc / sg_hi : nu;
        const double nu_lo = (m > 0) ? dc / sg_lo : nu;
        tw = 0.5 * (nu_hi - nu_lo);                     // trapezoid rule in nu as a weight (scipy trapezoid, :27)
        lnnu = log(nu);
        dc_prev = dc;
      }
      double g, bg;
      hm_gb_nu_shared(d, fc.c, nu, lnnu, g, bg);
      const double w1 = tw * (g * rM);                  // one-halo weight  (:530)
      const double w2_plain = tw * (bg * rM);           // two-halo weight  (:539)
      if (ROWS) { wrow[0] = w2_plain; wrow[(size_t)a.F * a.nMp] = w1; }
      double cW[PT];
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        cW[p] = 0.;
        if (p < a.P && (RED || !ROWS)) {
          const double rnorm = fc.rnorm[p];
          // <N(N-1)> or <N^2> plus variance (:466-473); no FMA contraction so exact zeros stay exact
          double fac = (a.correct_discrete && a.discrete[p]) ? __dmul_rn(amp[p], __dadd_rn(amp[p], -1.))
                                                             : __dmul_rn(amp[p], amp[p]);
          if (a.variance[p]) fac = __dadd_rn(fac, var[p]);
          const double an = amp[p] * rnorm;
          double uS;   // U/normalisation = uS * grid
          if (a.mode[p] == HM_GRID_UK) { cW[p] = an; uS = rnorm; }
          else if (a.mode[p] == HM_GRID_WK) { cW[p] = 1.; uS = rnorm / amp[p]; }
          else { cW[p] = 1.; uS = rnorm; }
          if (!ROWS) {
            if (m != 0) wout[(size_t)p * a.nMp + m] = w2_plain * cW[p];                    // e_p (e_p[0]: low-mass CTA)
            wout[(size_t)(a.P + p) * a.nMp + m] = __dmul_rn(__dmul_rn(w1, fac), uS * uS);   // d_p
          }
          if (RED) {
            red[p] = w1 * (an * rnorm);                                                    // shot noise integrand (:425)
            red[PT + p] = w2_plain * an;                                                   // simple two-halo (:442-446)
          }
        }
      }
      int idx = 2 * a.P;
#pragma unroll
      for (int u = 0; u < PT; ++u)
#pragma unroll
        for (int v = u + 1; v < PT; ++v)
          if (v < a.P && !ROWS) wout[(size_t)(idx++) * a.nMp + m] = w1 * cW[u] * cW[v];         // q_uv
    }
    if (RED) {
      // one reduction phase for the 2 PT sums: warp butterflies, then a fixed-order sum over the 8 warps
#pragma unroll
      for (int i = 0; i < NRED; ++i) {
        const bool used = (i % PT) < a.P;               // the others are zero
        const double t = used ? hm_warp_sum(red[i]) : 0.;
        if (lane == 0) sred[warp][i] = t;
      }
      __syncthreads();
      if (tid < NRED) {
        double t = 0.;
#pragma unroll
        for (int w = 0; w < HM_WCHUNK / 32; ++w) t += sred[w][tid];
        const int p = tid % PT, kind = tid / PT;        // kind 0: shot noise, 1: simple two-halo
        double* cp = a.cpart + ((size_t)b * (a.nchunks + 1) + chunk) * 2 * HM_MAXP;
        cp[kind * HM_MAXP + p] = (kind == 1 || (p < a.P && a.discrete[p])) ? t : 0.;
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weights of a sweep in the factored layout under default flags (no shot-noise / simple two-halo reductions): the
// sweep's own kernel.  hm_weights_kernel is straight-line code -- one pass over ~2300 instructions per warp, five
// inlined family bodies -- and ncu shows it waiting for instructions (27 % of its stall samples) and for the eight
// threads that prepare the per-model constants behind a CTA barrier (13 %).  Here a thread owns FOUR masses, so that
// each family body runs as a loop (fetched once, executed four times), the constants are prepared by the first F
// threads WHILE the others do the family-independent part (nu, ln nu, the trapezoid weight, the tracer vectors), and
// CTAs are small (128 threads, 512 masses): 1024 of them for 512 samples spread evenly over the SMs in one wave.
// grid (ceil(nMp / 512), G); the warps of a sample's CTAs share out its low-mass corrections once their masses are done.
// ---------------------------------------------------------------------------------------------------------------
#ifndef HM_WS_THREADS
#define HM_WS_THREADS 128
#define HM_WS_PER 4
#define HM_WS_MINB 7
#endif
__global__ void __launch_bounds__(HM_WS_THREADS, HM_WS_MINB) hm_weights_sweep_kernel(const hm_weights_args a) {
  constexpr int PT = HM_ROWS_MAXP;
  __shared__ hm_fam_consts<PT> sfc[HM_ROWS_MAXF];
  const int s = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* sig = a.sigma + (size_t)s * a.nM;
  const double M0 = a.M[0];

  if (tid < a.F) {                                      // per-model constants: ready when the others reach the barrier
    const hm_model_desc dm = a.models ? a.models[s * a.F + tid] : a.single;
    sfc[tid].d = dm;
    sfc[tid].c = hm_make_family_consts(dm);
  }
  // ---- family-independent part of my four masses: nu, ln nu and tw / M wait in shared memory (column tid: no bank
  // conflicts), so that the family loop below can run over them with a rolled loop and few registers ----
  __shared__ double s_nu[HM_WS_PER][HM_WS_THREADS], s_ln[HM_WS_PER][HM_WS_THREADS], s_tw[HM_WS_PER][HM_WS_THREADS];
  const double dc0 = a.models ? a.models[(size_t)s * a.F].dc : a.single.dc;
  const int m0 = chunk * (HM_WS_THREADS * HM_WS_PER) + tid;
  auto shared_part = [&](double dc) {                  // (rolled loops: this kernel waits for instructions, not for results)
#pragma unroll 1
    for (int j = 0; j < HM_WS_PER; ++j) {
      const int m = m0 + j * HM_WS_THREADS;
      const bool live = m < a.nM;
      const double sg = live ? sig[m] : 1.;
      const double sg_hi = (live && m + 1 < a.nM) ? sig[m + 1] : sg;
      const double sg_lo = (live && m > 0) ? sig[m - 1] : sg;
      const double rM = live ? 1. / a.M[m] : 0.;
      const double nu = dc / sg;
      // trapezoid rule in nu as a weight (scipy trapezoid, :27); one-sided at the ends, where the missing neighbour is
      // the mass itself (dc / sg is nu bit for bit)
      const double tw = 0.5 * (dc / sg_hi - dc / sg_lo);
      s_nu[j][tid] = nu;
      s_ln[j][tid] = log(nu);
      s_tw[j][tid] = tw * rM;
    }
  };
  shared_part(dc0);
  // tracer vectors A_p, T_p of this sample (see hm_weights_kernel)
#pragma unroll 1
  for (int j = 0; j < HM_WS_PER; ++j) {
    const int m = m0 + j * HM_WS_THREADS;
    if (m < a.nMp) {
      const bool live = m < a.nM;
      double* wv = a.weights + ((size_t)s * a.NVEC + 2 * a.F) * a.nMp + m;
#pragma unroll
      for (int p = 0; p < PT; ++p)
        if (p < a.P) {
          double A = 0., T = 0.;
          if (live) {
            const double amp = a.amplitude[p][(size_t)s * a.amp_stride[p] + m];
            double fac = (a.correct_discrete && a.discrete[p]) ? __dmul_rn(amp, __dadd_rn(amp, -1.)) : __dmul_rn(amp, amp);
            if (a.variance[p]) fac = __dadd_rn(fac, a.variance[p][(size_t)s * a.nM + m]);
            A = amp; T = fac;
            if (a.mode[p] != HM_GRID_UK) {
              const double ra = 1. / amp;
              A = 1.; T = __dmul_rn(fac, ra * ra);
            }
          }
          wv[(size_t)p * a.nMp] = A;
          wv[(size_t)(a.P + p) * a.nMp] = T;
        }
    }
  }
  __syncthreads();
  double dc_cur = dc0;
#pragma unroll 1
  for (int f = 0; f < a.F; ++f) {
    const hm_fam_consts<PT>& fc = sfc[f];
    const double dc = fc.d.dc;
    if (dc != dc_cur) {                                 // models of one sample usually share delta_c
      shared_part(dc);
      dc_cur = dc;
    }
    double* wrow = a.weights + ((size_t)s * a.NVEC + f) * a.nMp;
#pragma unroll 1
    for (int j = 0; j < HM_WS_PER; ++j) {
      const int m = m0 + j * HM_WS_THREADS;
      if (m < a.nM) {
        double g, bg;
        hm_gb_nu_shared(fc.d, fc.c, s_nu[j][tid], s_ln[j][tid], g, bg);
        const double twM = s_tw[j][tid];
        wrow[m] = twM * bg;                                              // w2: two-halo weight  (:539)
        wrow[(size_t)a.F * a.nMp + m] = twM * g;                         // w1: one-halo weight  (:530)
      } else if (m < a.nMp) {
        wrow[m] = 0.; wrow[(size_t)a.F * a.nMp + m] = 0.;
      }
    }
  }
  // ---- low-mass corrections A = 1 - int g b dnu (pyhalomodel.py:413-414) and the per-model scalars: the sample's models are
  // dealt over the warps of its chunk CTAs (warp w of chunk c takes models w * nchunks + c, + 4 nchunks, ...), so the launch
  // has no separate low-mass CTAs -- 1024 instead of 1536 CTAs for 512 samples, all resident at once ----
  {
    const double sg0 = sig[0], sg1 = sig[1];
    for (int f = warp * (int)gridDim.x + chunk; f < a.F; f += (HM_WS_THREADS / 32) * (int)gridDim.x) {
      const int b = s * a.F + f;
      const hm_model_desc d = a.models ? a.models[b] : a.single;
      const hm_family_consts c = hm_make_family_consts(d);
      const double nu = d.dc / sg0;
      double part = 0.;
#pragma unroll 1
      for (int j = 0; j < HM_GL_NODES / 32; ++j) part += hm_lowmass_node_shared(d, c, nu, lane + 32 * j);
      const double A = 1. - hm_warp_sum(part);
      if (lane == 0) {
        const double Aterm = A / M0;                    // multiplies W(k, M[0]) for mass tracers (:541-542)
        const double nu_hi = d.dc / sg1;
        const double tw = 0.5 * (nu_hi - nu);
        double g, bg;
        hm_gb_nu_shared(d, c, nu, log(nu), g, bg);
        double* sc = a.scal + (size_t)b * HM_NSCAL;
        for (int p = 0; p < HM_MAXP; ++p) {
          double cW = 0.;
          if (p < a.P) {
            const double rn = 1. / a.norm[p][b];
            const double an = a.amplitude[p][(size_t)s * a.amp_stride[p]] * rn;
            cW = (a.mode[p] == HM_GRID_UK) ? an : 1.;
            if (p < HM_ROWS_MAXP) {
              sc[HM_RSCAL + p] = (a.mode[p] == HM_GRID_UK) ? rn : 1.;       // s_p
              sc[HM_RSCAL + HM_ROWS_MAXP + p] = rn * rn;                               // rn_p^2
              sc[HM_RSCAL + 2 * HM_ROWS_MAXP + p] = a.mass[p] ? Aterm * cW : 0.;           // multiplies G_p(k, M0) in I_p
            }
          }
          sc[4 + p] = cW;
        }
        sc[0] = A;
        sc[1] = d.rhom;
        sc[2] = tw * (bg * (1. / M0));
        if (a.lowmass) a.lowmass[b] = A;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Shared pieces of the two quadrature kernels
// ---------------------------------------------------------------------------------------------------------------
struct hm_quad_args {
  int nk, nM, nMp, F, B, ntiles, nslices, nparts, MC;
  long long nitems;
  const double* gridW[HM_MAXP];      // two-halo / cross source
  const double* gridU[HM_MAXP];      // auto one-halo source (== gridW unless HM_GRID_SEPARATE)
  const double* weights;
  double* partial;                   // [B][nparts][nk][NW]
  // fused epilogue of the row-per-lane kernel (default flags only): spectra straight from the item's sums
  const double* scal;                // [B][HM_NSCAL]
  const double* Pk_lin;              // [G][nk]
  double* out;                       // [B][S][3][nk]
  int* item_counter;                 // row-per-lane kernel: items handed out beyond the first one per CTA (zeroed per launch)
};

// acc[r*NW + i] += w_i * (grid values the weight vector i multiplies), for one mass column.
template <int P, int R, bool SEP>
__device__ __forceinline__ void hm_accumulate(double* __restrict__ acc, const double* __restrict__ wv,
                                              const double (&x)[R][P], const double (&xu)[SEP ? R : 1][SEP ? P : 1]) {
  constexpr int NW = P + P * (P + 1) / 2;
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
      acc[r * NW + p] = fma(wv[p], x[r][p], acc[r * NW + p]);                                 // two-halo sums
      const double xa = SEP ? xu[SEP ? r : 0][SEP ? p : 0] : x[r][p];
      acc[r * NW + P + p] = fma(wv[P + p] * xa, xa, acc[r * NW + P + p]);                     // auto one-halo
    }
    int idx = 2 * P;
#pragma unroll
    for (int u = 0; u < P; ++u)
#pragma unroll
      for (int v = u + 1; v < P; ++v) {
        acc[r * NW + idx] = fma(wv[idx] * x[r][u], x[r][v], acc[r * NW + idx]);               // cross one-halo
        ++idx;
      }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fallback quadrature kernel: plain loads, one CTA per (model, R-row tile), whole mass axis (nslices == 1)
// ---------------------------------------------------------------------------------------------------------------
template <int P, int R, int VEC, bool SEP>
__global__ void __launch_bounds__(256) hm_quad_kernel(const hm_quad_args a) {
  constexpr int S = P * (P + 1) / 2;
  constexpr int NW = P + S;
  constexpr int NV = R * NW;
  constexpr int NG = (NV + 31) / 32;   // groups of 32 values for the transposing reduction
  __shared__ double red[8][NG * 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int ncol = (VEC == 2) ? (a.nM >> 1) : a.nM;
  const size_t rowstride = (size_t)a.nM;

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
      double x[VEC][R][P], xu[VEC][SEP ? R : 1][SEP ? P : 1];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (VEC == 2) {
            const double2 t = hm_ldg_stream2(a.gridW[p] + rowoff[r] + m);
            x[0][r][p] = t.x; x[VEC - 1][r][p] = t.y;
            if (SEP) {
              const double2 tu = hm_ldg_stream2(a.gridU[p] + rowoff[r] + m);
              xu[0][SEP ? r : 0][SEP ? p : 0] = tu.x; xu[VEC - 1][SEP ? r : 0][SEP ? p : 0] = tu.y;
            }
          } else {
            x[0][r][p] = hm_ldg_stream1(a.gridW[p] + rowoff[r] + m);
            if (SEP) xu[0][SEP ? r : 0][SEP ? p : 0] = hm_ldg_stream1(a.gridU[p] + rowoff[r] + m);
          }
        }
      double wv[VEC][NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        if (VEC == 2) {
          const double2 t = *reinterpret_cast<const double2*>(w + (size_t)i * a.nMp + m);
          wv[0][i] = t.x; wv[VEC - 1][i] = t.y;
        } else {
          wv[0][i] = w[(size_t)i * a.nMp + m];
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) hm_accumulate<P, R, SEP>(acc, wv[e], x[e], xu[e]);
    }

    // reduce over the mass axis: lanes (transposing shuffle network), then warps (shared memory)
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      double v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = acc[g * 32 + i];
      red[warp][g * 32 + lane] = hm_warp_transpose_sum32(v, lane);
    }
    __syncthreads();
    if (tid < NV) {
      double t = 0.;
      for (int wi = 0; wi < nwarps; ++wi) t += red[wi][tid];
      const int r = tid / NW, i = tid - r * NW;
      if (row0 + r < a.nk) a.partial[((size_t)b * a.nk + row0 + r) * NW + i] = t;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming quadrature kernel: TMA bulk copies -> shared-memory ring -> register accumulation
// ---------------------------------------------------------------------------------------------------------------
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void hm_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// 16-byte LDGSTS (global -> shared without registers, L2 only) and its completion on an mbarrier
__device__ __forceinline__ void hm_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void hm_cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

template <int P, int R, int H>      // H = pairs of mass columns per consumer thread (slice width H * HM_MC)
struct hm_stream_cfg {
  static constexpr int S = P * (P + 1) / 2;
  static constexpr int NW = P + S;
  static constexpr int NV = R * NW;
  static constexpr int NG = (NV + 31) / 32;
  static constexpr int MC = H * HM_MC;
  static constexpr int STAGE_DOUBLES = R * P * MC;
  static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
  static constexpr int BUDGET = 226 * 1024 - 256;
#ifndef HM_ROWS_NS_MAX
#define HM_ROWS_NS_MAX 8
#endif
  static constexpr int NS = (BUDGET / STAGE_BYTES) > HM_ROWS_NS_MAX ? HM_ROWS_NS_MAX : (BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = NS * STAGE_BYTES + 256;
};

// (model, mass slice, row tile) cursor over the item range of one CTA, advanced without integer division
struct hm_item_cursor {
  int b, slice, tile;
  __device__ __forceinline__ void init(long long item, int nslices, int ntiles) {
    const long long per_model = (long long)nslices * ntiles;
    b = (int)(item / per_model);
    const int rem = (int)(item - (long long)b * per_model);
    slice = rem / ntiles;
    tile = rem - slice * ntiles;
  }
  __device__ __forceinline__ void next(int nslices, int ntiles) {
    if (++tile == ntiles) {
      tile = 0;
      if (++slice == nslices) { slice = 0; ++b; }
    }
  }
};

template <int P, int R, int H>
// 9 warps: one SM sub-partition (16K registers) hosts three of them, hence at most 168 registers per thread
__global__ void __launch_bounds__(HM_STREAM_THREADS, 1) hm_quad_stream_kernel(const hm_quad_args a) {
  using C = hm_stream_cfg<P, R, H>;
  constexpr int NW = C::NW, NS = C::NS, MC = C::MC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)NS * C::STAGE_DOUBLES);    // full[NS], empty[NS]
  const uint32_t bar_full = hm_smem_addr(bars), bar_empty = hm_smem_addr(bars + NS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      hm_mbar_init(bar_full + 8 * s, 1);
      hm_mbar_init(bar_empty + 8 * s, HM_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // contiguous range of (model, mass slice, row tile) items for this CTA
  const long long per = (a.nitems + gridDim.x - 1) / gridDim.x;
  const long long it0 = (long long)blockIdx.x * per;
  const long long it1 = (it0 + per < a.nitems) ? it0 + per : a.nitems;
  if (it0 >= it1) return;
  const int nloc = (int)(it1 - it0);
  hm_item_cursor cur;
  cur.init(it0, a.nslices, a.ntiles);

  if (warp == HM_CONSUMER_WARPS) {
    // ------------------------------ producer: one lane drives the TMA engine ------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t parity = 0;
      for (int n = 0; n < nloc; ++n) {
        const int smp = cur.b / a.F, m0 = cur.slice * MC, row0 = cur.tile * R;
        const int cols = (a.nM - m0 < MC) ? a.nM - m0 : MC;
        const uint32_t bytes = (uint32_t)cols * 8u;
        hm_mbar_wait(bar_empty + 8 * s, parity ^ 1u);        // slot drained by all consumer warps
        hm_mbar_arrive_expect_tx(bar_full + 8 * s, bytes * (uint32_t)(R * P));
        const uint32_t dst0 = hm_smem_addr(ring + (size_t)s * C::STAGE_DOUBLES);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int row = (row0 + r < a.nk) ? row0 + r : a.nk - 1;   // clamp: duplicates are never written
          const size_t off = ((size_t)smp * a.nk + row) * (size_t)a.nM + m0;
#pragma unroll
          for (int p = 0; p < P; ++p)
            hm_bulk_g2s(dst0 + (uint32_t)((r * P + p) * MC * 8), a.gridW[p] + off, bytes, bar_full + 8 * s);
        }
        cur.next(a.nslices, a.ntiles);
        if (++s == NS) { s = 0; parity ^= 1u; }
      }
    }
    return;
  }

  // ------- consumers: 8 independent warps, 64*H mass columns per warp, H column pairs per thread ------------
  // No CTA-wide synchronisation: every warp reduces its own columns and writes its own partial row sums;
  // the epilogue kernel adds the parts in fixed order.
  double wv[H][2][NW];
  int cur_b = -1, cur_slice = -1;
  int s = 0;
  uint32_t parity = 0;
  for (int n = 0; n < nloc; ++n) {
    const int m0 = cur.slice * MC + 2 * tid, row0 = cur.tile * R;      // my pair h covers masses m0 + h*HM_MC + {0,1}
    if (cur.b != cur_b || cur.slice != cur_slice) {   // weights of my columns stay in registers for the slice
      cur_b = cur.b; cur_slice = cur.slice;
#pragma unroll
      for (int h = 0; h < H; ++h) {
          const int m = m0 + h * HM_MC;
          const double* __restrict__ w = a.weights + (size_t)cur.b * NW * a.nMp + m;
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            const double2 t = (m < a.nM) ? *reinterpret_cast<const double2*>(w + (size_t)i * a.nMp) : make_double2(0., 0.);
            wv[h][0][i] = t.x; wv[h][1][i] = t.y;
          }
        }
    }
    constexpr int NVF = R * NW, NGF = (NVF + 31) / 32;
    double acc[NGF * 32];
#pragma unroll
    for (int i = 0; i < NGF * 32; ++i) acc[i] = 0.;
    const double xu[1][1] = {{0.}};
    hm_mbar_wait(bar_full + 8 * s, parity);                  // TMA bytes have landed
    const double* st = ring + (size_t)s * C::STAGE_DOUBLES + 2 * tid;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const bool active = m0 + h * HM_MC < a.nM;
      double x[2][R][P];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const double2 t = active ? *reinterpret_cast<const double2*>(st + (r * P + p) * MC + h * HM_MC)
                                   : make_double2(0., 0.);
          x[0][r][p] = t.x; x[1][r][p] = t.y;
        }
      if (h == H - 1) {                                      // all my values are in registers: release the slot early
        __syncwarp();
        if (lane == 0) hm_mbar_arrive(bar_empty + 8 * s);
      }
      hm_accumulate<P, R, false>(acc, wv[h][0], x[0], xu);
      hm_accumulate<P, R, false>(acc, wv[h][1], x[1], xu);
    }

    // all R*NW sums of this tile go through ONE transposing reduction (lane j ends up with value j)
    const int nvalid = ((a.nk - row0 < R) ? a.nk - row0 : R) * NW;
#pragma unroll
    for (int g = 0; g < NGF; ++g) {
      double v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = acc[g * 32 + i];
      const double t = hm_warp_transpose_sum32(v, lane);
      const int rem = g * 32 + lane;
      if (rem < nvalid) {
        // partial[model][slice*8 + warp][row0 + r][i]: the R*NW values of a tile are contiguous -> coalesced store
        const size_t model = (size_t)cur.b;
        a.partial[((model * a.nparts + cur.slice * HM_CONSUMER_WARPS + warp) * a.nk + row0) * NW + rem] = t;
      }
    }
    cur.next(a.nslices, a.ntiles);
    if (++s == NS) { s = 0; parity ^= 1u; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row-per-lane streaming kernel for sweeps that evaluate FM = 2..5 models (mass functions) on the grids of one sample.
//
// With FM*(P+S) = 25 sums per grid row and only 16 bytes of grid per (k, M) point, the column-per-lane kernel above
// drowns in its warp reductions (one 32-value shuffle network per 50 FMAs; ncu: 423 instructions per warp tile,
// FP64 pipe 29 %).  Here a lane owns two wavenumber rows (lane and lane + 32 of a 64-row tile; ONE row of a 32-row tile
// at three tracers, whose 9 FM sums per row would not fit twice) for the whole mass axis, so its FM*(P+S) sums per row
// stay in registers from the first mass to the last and nothing is ever reduced across lanes.  The weights of a mass column are the same for every lane: they ride in the ring next to
// the grid tile and are read with broadcast 16-byte shared-memory loads, one load feeding four FMAs (two rows x two
// columns).
//
// Feeding: ONE producer lane drives the TMA unit with 2-D tensor-map copies (cp.async.bulk.tensor.2d, SASS UTMALDG):
// per 32-column stage one 64-row x 16-column box per tracer and half stage (128-byte rows, SWIZZLE_128B: the
// eight lanes of a quarter warp read eight different 16-byte bank groups without padding) plus one FM*(P+S) x 32 box of
// the weight vectors; ragged mass and wavenumber edges are zero-filled by the TMA unit.  (Round 1 fed this ring with
// four warps of LDGSTS copies after 153 one-dimensional bulk copies per stage had serialised on the TMA unit.)
//
// An item is (sample, 64-row tile); the 8 consumer warps own 4 of a stage's 32 columns each and add their sums at the
// end of the item in a binary tree through shared memory (warp w+4 -> w, w+2 -> w, 1 -> 0: a fixed summation order).
// The hand-overs of the tree are mbarrier pairs per writer, so a warp that has handed its sums over starts the next
// item's stages at once while the others are still adding: the combine overlaps the streaming.  Items are dealt
// round-robin (CTA c takes c, c + grid, ...), so the row tiles of one sample run at the same time on neighbouring
// SMs and its weight vectors come from DRAM once and from L2 for the other tiles.
// ---------------------------------------------------------------------------------------------------------------
// 16-byte shared-memory load from a 32-bit shared-window address
__device__ __forceinline__ double2 hm_lds2(uint32_t addr) {
  return *reinterpret_cast<const double2*>(__cvta_shared_to_generic(addr));
}

__device__ __forceinline__ void hm_tma_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// the same copy with an L2 eviction-priority hint (the grids are read once: evict_first keeps them from displacing the
// sample's weight vectors, which the other row tiles of the sample re-read from L2)
__device__ __forceinline__ void hm_tma_2d_hint(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar), "l"(pol) : "memory");
}

template <int P, int FM>
struct hm_rows_cfg {
  static constexpr int S = P * (P + 1) / 2;
  static constexpr int NW = P + S;
  static constexpr int NV = FM * NW;                          // sums per grid row
  static constexpr int NVEC = 2 * FM + 2 * P;                 // factored weight vectors per sample: w2^f, w1^f, A_p, T_p
  static constexpr int CS = HM_ROWS_CS;                       // mass columns per stage
  static constexpr int BOXC = 16;                             // columns per grid box: 128 bytes, the swizzle span
  static constexpr int NBOX = CS / BOXC;
  static constexpr int NR = (P <= 2) ? 2 : 1;                 // wavenumber rows per lane: 2 NR NV accumulator registers (100 at
                                                              // P = 2, FM = 5; three tracers have NV = 9 FM sums per row: one row)
  static constexpr int RT = 32 * NR;                          // rows per item
  static constexpr int BOX_BYTES = RT * BOXC * 8;
  static constexpr int DATA_BYTES = P * NBOX * BOX_BYTES;
  static constexpr int W_BYTES = NVEC * CS * 8;
  static constexpr int TX_BYTES = DATA_BYTES + W_BYTES;
  static constexpr int STAGE_BYTES = ((TX_BYTES + 1023) / 1024) * 1024;
  static constexpr int COMB_DOUBLES = RT * NV;                // one reader's buffer of the combine tree
#ifdef HM_ROWS_DEBUG_NO_COMBINE
  static constexpr int COMB_BYTES = 0;
#else
  static constexpr int COMB_BYTES = 4 * COMB_DOUBLES * 8;
#endif
  static constexpr int BAR_BYTES = 384;                       // mbarriers (2 NS + 16) and the ring of item descriptors (8 ints)
  static constexpr int BUDGET = 227 * 1024 - BAR_BYTES - COMB_BYTES - 1024;   // 1024: alignment slack of the ring
#ifndef HM_ROWS_NS_MAX
#define HM_ROWS_NS_MAX 8
#endif
  static constexpr int NS = (BUDGET / STAGE_BYTES) > HM_ROWS_NS_MAX ? HM_ROWS_NS_MAX : (BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = NS * STAGE_BYTES + COMB_BYTES + BAR_BYTES + 1024;
  static_assert(CS == 4 * HM_CONSUMER_WARPS, "a consumer warp owns four columns of a stage");
  static_assert(NS >= 2, "the ring needs two stages");
};

template <int P, int FM, bool FUSED>
#ifdef HM_ROWS_LAUNCH_REGS      // (experiment: a smaller register allocation at launch leaves room for a co-resident weights CTA)
__global__ void __maxnreg__(HM_ROWS_LAUNCH_REGS) hm_quad_rows_kernel(const hm_quad_args a,
#else
__global__ void __launch_bounds__(HM_ROWS_LAUNCH_THREADS, 1) hm_quad_rows_kernel(const hm_quad_args a,
#endif
                                                                          const __grid_constant__ CUtensorMap tmG0,
                                                                          const __grid_constant__ CUtensorMap tmG1,
                                                                          const __grid_constant__ CUtensorMap tmG2,
                                                                          const __grid_constant__ CUtensorMap tmW) {
  using C = hm_rows_cfg<P, FM>;
  constexpr int NW = C::NW, NV = C::NV, NS = C::NS, RT = C::RT, CS = C::CS, NR = C::NR;
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B boxes want 1024-byte aligned destinations
  // (the padding is made opaque to the compiler, which otherwise recomputes it -- S2UR of the shared window, three integer
  // operations and a scoreboard wait -- at the top of every stage instead of keeping one register)
  uint32_t pad = (1024u - (hm_smem_addr(smem_dyn) & 1023u)) & 1023u;
  asm volatile("" : "+r"(pad));
  unsigned char* smem_raw = smem_dyn + pad;
  double* comb = reinterpret_cast<double*>(smem_raw + (size_t)NS * C::STAGE_BYTES);        // [4 readers][NV][RT rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * C::STAGE_BYTES + C::COMB_BYTES);
  const uint32_t ring = hm_smem_addr(smem_raw);
  const uint32_t bar_full = hm_smem_addr(bars), bar_empty = hm_smem_addr(bars + NS);
  const uint32_t bar_pfull = hm_smem_addr(bars + 2 * NS), bar_pempty = hm_smem_addr(bars + 2 * NS + 8);   // per writer warp
  volatile int* desc = reinterpret_cast<volatile int*>(bars + 2 * NS + 16);                               // item queue, see below

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      hm_mbar_init(bar_full + 8 * s, 1);                  // the producer's arrive.expect_tx
      hm_mbar_init(bar_empty + 8 * s, HM_CONSUMER_WARPS);
    }
    for (int w = 0; w < 8; ++w) {
      hm_mbar_init(bar_pfull + 8 * w, 1);
      hm_mbar_init(bar_pempty + 8 * w, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // (sample, row tile) items, each walking the whole mass axis, are handed out DYNAMICALLY: CTA c starts with item c and
  // draws the next ones from a global counter, so SMs that run slower (or share their cycles with another kernel) take
  // fewer items and the launch has no long tail.  Consecutive items are the tiles of one sample, so its weight vectors
  // still come from L2 for all but the first.  The producer publishes the item of the stages it is about to fill in a
  // small ring of descriptors BEFORE it arrives on the first stage's barrier; consumers read it after their wait on that
  // barrier (release/acquire through the mbarrier).  -1 ends the kernel.
  const int nitems = (int)a.nitems;                           // (host: nitems < 2^31)
  const int nstages = (a.nM + CS - 1) / CS;
  int s = 0;
  uint32_t parity = 0;

  if (warp >= HM_CONSUMER_WARPS) {
    // ---------------------------------- producer: one lane drives the TMA unit ----------------------------------
    // Its warpgroup hands its registers back first: twelve warps at the launch bound of 168 registers fill the SM's
    // register file, and what the producer's four warps release lets the consumers grow to 232 (setmaxnreg below),
    // room for the 100 accumulators plus the operands of the next column pair in flight.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HM_ROWS_PRODUCER_REGS));
    if (warp == HM_CONSUMER_WARPS && lane == 0) {
      int it = blockIdx.x;
#ifdef HM_ROWS_EVICT_FIRST
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
      for (uint32_t seq = 0;; ++seq) {
        const bool done = it >= nitems;
        hm_mbar_wait(bar_empty + 8 * s, parity ^ 1u);       // slot drained by all consumer warps (first stage of the item)
        desc[seq & 7u] = done ? -1 : it;
        if (done) {
          hm_mbar_arrive(bar_full + 8 * s);                 // an empty stage that carries the end marker
          break;
        }
        const int smp = it / a.ntiles, tile = it - smp * a.ntiles;
        const int row = smp * a.nk + tile * RT;             // first row of the item in the [G*nk][nM] tensors
        const int wrow = smp * C::NVEC;                     // first of its weight vectors in the [G*NVEC][nMp] tensor
        for (int st = 0; st < nstages; ++st) {
          const int m0 = st * CS;
          if (st > 0) hm_mbar_wait(bar_empty + 8 * s, parity ^ 1u);
          hm_mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)C::TX_BYTES);   // OOB parts of a box count as well
          const uint32_t dst = ring + (uint32_t)s * C::STAGE_BYTES;
#pragma unroll
          for (int p = 0; p < P; ++p)
#pragma unroll
            for (int cb = 0; cb < C::NBOX; ++cb)
#ifdef HM_ROWS_EVICT_FIRST
              hm_tma_2d_hint(dst + (uint32_t)((p * C::NBOX + cb) * C::BOX_BYTES), p == 0 ? &tmG0 : (p == 1 ? &tmG1 : &tmG2),
                             m0 + cb * C::BOXC, row, bar_full + 8 * s, pol);
#else
              hm_tma_2d(dst + (uint32_t)((p * C::NBOX + cb) * C::BOX_BYTES), p == 0 ? &tmG0 : (p == 1 ? &tmG1 : &tmG2),
                        m0 + cb * C::BOXC, row, bar_full + 8 * s);
#endif
          hm_tma_2d(dst + (uint32_t)C::DATA_BYTES, &tmW, m0, wrow, bar_full + 8 * s);
          if (++s == NS) { s = 0; parity ^= 1u; }
        }
        it = (int)gridDim.x + atomicAdd(a.item_counter, 1);
      }
    }
    return;
  }

  // ------------------------------------------- consumer warps --------------------------------------------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HM_ROWS_CONSUMER_REGS));
  // my four columns of a stage: 4*warp .. 4*warp + 3 = 16-byte chunks q0, q0 + 1 of box warp / 4; inside a box the
  // chunk index is XORed with (row & 7) (SWIZZLE_128B)
  const uint32_t boxoff = (uint32_t)(warp >> 2) * C::BOX_BYTES + (uint32_t)lane * 128u;
  const uint32_t q0 = 2u * (warp & 3), sw = lane & 7u;
  uint32_t xo0 = boxoff + (((q0) ^ sw) << 4), xo1 = boxoff + (((q0 + 1u) ^ sw) << 4);
  uint32_t wcol = 4u * warp * 8u;                                  // byte offset of my columns in a weight row
  asm volatile("" : "+r"(xo0), "+r"(xo1), "+r"(wcol));            // (kept in registers: not rebuilt from %tid every stage)
  uint32_t landed = 0;                                        // the early probe of the current stage's barrier succeeded
  for (uint32_t item_no = 0;; ++item_no) {
    if (!landed) hm_mbar_wait(bar_full + 8 * s, parity);      // first stage of the next item, or the end marker
    landed = 1;
    const int it = desc[item_no & 7u];
    if (it < 0) break;
    const int smp = it / a.ntiles, tile = it - smp * a.ntiles;
    double acc[NR][NV];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int j = 0; j < NV; ++j) acc[r][j] = 0.;
#pragma unroll 1
    for (int st = 0; st < nstages; ++st) {
      // (32-bit shared-window addresses, turned into pointers only at the load: a generic stage pointer made the compiler
      // rebuild the shared window's base -- S2UR, ULEA and a scoreboard wait -- at the top of every stage)
      const uint32_t stage = ring + (uint32_t)s * C::STAGE_BYTES;
      const uint32_t wsm = stage + C::DATA_BYTES + wcol;
      if (!landed) hm_mbar_wait(bar_full + 8 * s, parity);    // the stage's boxes have landed
      const int sn = (s + 1 == NS) ? 0 : s + 1;
      const uint32_t pn = (s + 1 == NS) ? parity ^ 1u : parity;
#pragma unroll
      for (int cp = 0; cp < 2; ++cp) {
        // The barrier probe takes a few hundred cycles to answer even when the data is there: ask about the NEXT stage
        // in the middle of this one (the ring runs on across items; past the last stage the answer is never used)
        if (cp == 1) landed = hm_mbar_test(bar_full + 8 * sn, pn);
        double2 x[NR][P];                                      // [row][tracer] of this column pair
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
          for (int p = 0; p < P; ++p)
            x[r][p] = hm_lds2(stage + (uint32_t)((p * C::NBOX) * C::BOX_BYTES + r * (32 * 128)) + (cp == 0 ? xo0 : xo1));
        // basis values of my rows and two columns, in the order of the sums: A_p G_p, T_p G_p^2, (A_u G_u)(A_v G_v).
        // The tracer factors are lane-uniform: broadcast 16-byte loads, like the mass-function weights below.
        double2 v[NR][NW];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          const double2 Av = hm_lds2(wsm + (uint32_t)(((2 * FM + p) * CS + 2 * cp) * 8));
          const double2 Tv = hm_lds2(wsm + (uint32_t)(((2 * FM + P + p) * CS + 2 * cp) * 8));
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            v[r][p] = make_double2(Av.x * x[r][p].x, Av.y * x[r][p].y);
            // no FMA contraction and T first: T = 0 (e.g. central galaxies) keeps the sum an exact zero
            v[r][P + p] = make_double2(__dmul_rn(__dmul_rn(Tv.x, x[r][p].x), x[r][p].x),
                                       __dmul_rn(__dmul_rn(Tv.y, x[r][p].y), x[r][p].y));
          }
        }
        {
          int idx = 2 * P;
#pragma unroll
          for (int u = 0; u < P; ++u)
#pragma unroll
            for (int w = u + 1; w < P; ++w) {
#pragma unroll
              for (int r = 0; r < NR; ++r) v[r][idx] = make_double2(v[r][u].x * v[r][w].x, v[r][u].y * v[r][w].y);
              ++idx;
            }
        }
#pragma unroll
        for (int f = 0; f < FM; ++f) {
          const double2 w2 = hm_lds2(wsm + (uint32_t)((f * CS + 2 * cp) * 8));                          // two-halo weight
          const double2 w1 = hm_lds2(wsm + (uint32_t)(((FM + f) * CS + 2 * cp) * 8));                   // one-halo weight
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            const double2 w = (i < P) ? w2 : w1;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              acc[r][f * NW + i] = fma(w.x, v[r][i].x, acc[r][f * NW + i]);
              acc[r][f * NW + i] = fma(w.y, v[r][i].y, acc[r][f * NW + i]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) hm_mbar_arrive(bar_empty + 8 * s);       // my reads of the slot are done
      s = sn; parity = pn;
    }

    // ---- combine the eight column groups: binary tree over the warps, pipelined with the next item.  The tree runs on
    // VIRTUAL warp numbers vw = (warp - item) mod 8, so the root -- which reads three buffers, applies the per-model
    // factors and stores the item -- is a different warp every item and no warp is the slowest on average.  Reader r
    // owns buffer r; writer x signals pfull[x] once per item and its reader drains it with pempty[x], so every pair
    // barrier completes exactly one phase per item whoever plays the role.  A writer waits until the previous occupant
    // of its target buffer has been drained: the previous writer of this item, or the buffer's last writer of the
    // previous item. ----
    const uint32_t ph = item_no & 1u;
    const int vw = (int)((warp - item_no) & 7u);
    // (fused epilogue) the root warp fetches P_lin of its rows now, so that the loads fly while it waits for the other
    // warps' sums instead of serialising behind the last hand-over
    double pl[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) pl[r] = 0.;
    if constexpr (FUSED) {
      if (vw == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int row = tile * RT + r * 32 + lane;
          pl[r] = a.Pk_lin[(size_t)smp * a.nk + (row < a.nk ? row : a.nk - 1)];
        }
      }
    }
#ifdef HM_ROWS_DEBUG_NO_COMBINE
    if (false)
#endif
#pragma unroll
    for (int step = 4; step >= 1; step >>= 1) {
      if (vw < step) {
        const int xw = vw + step;
        hm_mbar_wait(bar_pfull + 8 * xw, ph);
        const double* src = comb + (size_t)vw * C::COMB_DOUBLES + lane;
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
          for (int j = 0; j < NV; ++j) acc[r][j] += src[j * RT + r * 32];
        __syncwarp();
        if (lane == 0) hm_mbar_arrive(bar_pempty + 8 * xw);
      } else if (vw < 2 * step) {
        const int t = vw - step;
        int pred;
        uint32_t pph;
        if (step == 4) { pred = (t == 0) ? 1 : (t == 1) ? 3 : vw; pph = ph ^ 1u; }
        else if (step == 2) { pred = t + 4; pph = ph; }
        else { pred = 2; pph = ph; }
        hm_mbar_wait(bar_pempty + 8 * pred, pph);
        double* dst = comb + (size_t)t * C::COMB_DOUBLES + lane;
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
          for (int j = 0; j < NV; ++j) dst[j * RT + r * 32] = acc[r][j];
        __syncwarp();
        if (lane == 0) hm_mbar_arrive(bar_pfull + 8 * vw);
      }
    }
    if (vw == 0) {      // the item's RT x NV sums, coalesced over rows
      constexpr int S = C::S;
      // first mass column of my rows: the low-mass term A W(k, M0) / M0 of mass tracers (:541-542)
      double g0[NR][P];
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int row = tile * RT + r * 32 + lane;
#pragma unroll
        for (int p = 0; p < P; ++p)
          g0[r][p] = a.gridW[p][((size_t)smp * a.nk + (row < a.nk ? row : a.nk - 1)) * (size_t)a.nM];
      }
#pragma unroll
      for (int f = 0; f < FM; ++f) {
        const size_t b = (size_t)smp * FM + f;
        const double* sc = a.scal + b * HM_NSCAL;
        const double rhom = sc[1];
        double sp[P], rn2[P], low[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
          sp[p] = sc[HM_RSCAL + p]; rn2[p] = sc[HM_RSCAL + HM_ROWS_MAXP + p]; low[p] = sc[HM_RSCAL + 2 * HM_ROWS_MAXP + p];
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int row = tile * RT + r * 32 + lane;
          // the per-(model, tracer) factors the factored weights left out (hm_weights_kernel)
          double* t = &acc[r][f * NW];
#pragma unroll
          for (int p = 0; p < P; ++p) {
            t[p] = fma(low[p], g0[r][p], t[p] * sp[p]);
            t[P + p] = __dmul_rn(t[P + p], rn2[p]);
          }
          {
            int idx = 2 * P;
#pragma unroll
            for (int u = 0; u < P; ++u)
#pragma unroll
              for (int w = u + 1; w < P; ++w) t[idx++] *= sp[u] * sp[w];
          }
          if (row < a.nk) {
            if constexpr (FUSED) {
              // default flags (no beta_NL, simple two-halo, k_trunc or shot-noise fix-up): the spectra follow from the
              // sums directly, with the arithmetic of hm_epilogue_kernel (pyhalomodel.py:456-501)
              const double plin = pl[r];
              double* o = a.out + b * S * 3 * (size_t)a.nk + row;
              int sidx = 0, icross = 2 * P;
#pragma unroll
              for (int u = 0; u < P; ++u)
#pragma unroll
                for (int v = u; v < P; ++v) {
                  const double Iu = t[u] * rhom, Iv = t[v] * rhom;                                        // :540-543
                  const double p2h = __dmul_rn(__dmul_rn(Iu, Iv), plin);                                  // :524
                  const double p1h = t[u == v ? P + u : icross] * rhom;                                   // :532
                  if (u != v) ++icross;
                  o[(size_t)(sidx * 3 + 0) * a.nk] = p2h;
                  o[(size_t)(sidx * 3 + 1) * a.nk] = p1h;
                  o[(size_t)(sidx * 3 + 2) * a.nk] = __dadd_rn(p2h, p1h);                                 // :500-501
                  ++sidx;
                }
            } else {
#pragma unroll
              for (int i = 0; i < NW; ++i) a.partial[(b * NW + i) * (size_t)a.nk + row] = t[i];           // [model][i][row]
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Non-linear halo bias: I_NL(k) = rhom^2 [ trapz2d(beta W1 W2 g1 g2 b1 b2 / M1 M2) + low-mass edge terms ]
// (model._I_beta, pyhalomodel.py:546-571).  With a_p[m] = w2[m] W_p(k,m) the double trapezoid is the bilinear form
// a_u^T beta_k a_v, and the edge terms are L_u (beta_k[0,:] . a_v) (I12), L_v (beta_k[:,0] . a_u) (I21) and
// L_u L_v beta_k[0,0] (I11) with L_p = mass_p A W_p(k,M0)/M0.  Pure HBM streaming of beta[k] (nM x nM doubles per
// wavenumber): one CTA per (model, k), a warp per matrix row, the P vectors a_v staged in shared memory.
// ---------------------------------------------------------------------------------------------------------------
struct hm_beta_args {
  int nk, nM, nMp, F, do_I11, do_I12_I21, vec2;
  const double* M;
  const double* beta;                // [G][nk][nM][nM]
  const double* gridW[HM_MAXP];
  int mass[HM_MAXP];
  const double* weights;
  const double* scal;
  double* inl;                       // [B][S][nk]
  int split;                         // row chunks per (model, k) matrix
  double* part;                      // [B nk][split][S + 2 P] partial sums of the chunks
  int* count;                        // [B nk] chunks finished (zeroed per launch)
};

template <int P>
__global__ void __launch_bounds__(256) hm_beta2h_kernel(const hm_beta_args a) {
  constexpr int S = P * (P + 1) / 2, NW = P + S;
  extern __shared__ __align__(16) double bsm[];
  double* av = bsm;                                     // [P][nMp]  a_p[m] = w2[m] W_p(k,m) (no low-mass term)
  double* red = av + (size_t)P * a.nMp;                 // [8 warps][S + P]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mat = blockIdx.x / a.split, chunk = blockIdx.x - mat * a.split;    // (model, wavenumber) matrix and my row chunk
  const int ik = mat % a.nk, b = mat / a.nk, smp = b / a.F;
  const int rows_per = (a.nM + a.split - 1) / a.split;
  const int r0 = chunk * rows_per, r1 = (r0 + rows_per < a.nM) ? r0 + rows_per : a.nM;
  const double* sc = a.scal + (size_t)b * HM_NSCAL;
  const double A = sc[0], rhom = sc[1], w20 = sc[2], M0 = a.M[0];
  const double* w = a.weights + (size_t)b * NW * a.nMp;
  const size_t goff = ((size_t)smp * a.nk + ik) * (size_t)a.nM;
  for (int m = tid; m < a.nMp; m += blockDim.x)
#pragma unroll
    for (int p = 0; p < P; ++p) {
      double v = 0.;
      if (m < a.nM) {
        const double g = a.gridW[p][goff + m];
        v = (m == 0) ? (w20 * sc[4 + p]) * g : w[(size_t)p * a.nMp + m] * g;   // e_p[0] carries the low-mass term: rebuild a_p[0]
      }
      av[(size_t)p * a.nMp + m] = v;
    }
  __syncthreads();

  const double* bk = a.beta + ((size_t)smp * a.nk + ik) * (size_t)a.nM * a.nM;
  double pair_acc = 0.;      // lane s < S: sum_m1 a_u[m1] y_v[m1];  lane S+p: sum_m1 beta[m1,0] a_p[m1]  (column 0)
  double y0 = 0.;            // lane p < P of warp 0: y_p at row 0 = beta[0,:] . a_p
  int pu = 0, pv = 0;
  {
    int cnt = 0;
#pragma unroll
    for (int uu = 0; uu < P; ++uu)
#pragma unroll
      for (int vv = uu; vv < P; ++vv) {
        if (cnt == lane) { pu = uu; pv = vv; }
        ++cnt;
      }
    if (lane >= S && lane < S + P) pu = lane - S;
  }
  for (int m1 = r0 + warp; m1 < r1; m1 += 8) {
    const double* row = bk + (size_t)m1 * a.nM;
    double y[P];
#pragma unroll
    for (int p = 0; p < P; ++p) y[p] = 0.;
    if (a.vec2) {        // even nM, 16-byte aligned rows: 8 independent 16-byte streaming loads in flight per lane
      const int npair = a.nM >> 1;
      for (int c0 = lane; c0 < npair; c0 += 32 * 8) {
        double2 bv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + 32 * j;
          bv[j] = (c < npair) ? hm_ldg_stream2(row + 2 * c) : make_double2(0., 0.);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + 32 * j;
          if (c < npair) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
              const double2 avv = *reinterpret_cast<const double2*>(av + (size_t)p * a.nMp + 2 * c);
              y[p] = fma(bv[j].x, avv.x, fma(bv[j].y, avv.y, y[p]));
            }
          }
        }
      }
    } else {
      for (int c = lane; c < a.nM; c += 32) {
        const double bv = hm_ldg_stream1(row + c);
#pragma unroll
        for (int p = 0; p < P; ++p) y[p] = fma(bv, av[(size_t)p * a.nMp + c], y[p]);
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) y[p] = hm_warp_sum(y[p]);        // every lane holds y_p[m1] = beta[m1,:] . a_p
    double yv = 0.;
#pragma unroll
    for (int p = 0; p < P; ++p) yv = (p == pv) ? y[p] : yv;
    if (lane < S) pair_acc = fma(av[(size_t)pu * a.nMp + m1], yv, pair_acc);
    else if (lane < S + P) pair_acc = fma(row[0], av[(size_t)pu * a.nMp + m1], pair_acc);
    if (m1 == 0) {
#pragma unroll
      for (int p = 0; p < P; ++p) y0 = (p == lane) ? y[p] : y0;
    }
  }
  if (lane < S + P) red[warp * (S + P) + lane] = pair_acc;
  __shared__ double y0s[HM_MAXP];
  __shared__ int s_last;
  if (warp == 0 && lane < P) y0s[lane] = y0;          // (row 0 is warp 0's first row of chunk 0)
  __syncthreads();
  // my chunk's sums: the eight warps in order; then the chunk that finds the others done adds the chunks in order
  double* mypart = a.part + ((size_t)mat * a.split + chunk) * (S + 2 * P);
  if (tid < S + P) {
    double t = 0.;
    for (int wi = 0; wi < 8; ++wi) t += red[wi * (S + P) + tid];
    mypart[tid] = t;
  } else if (tid < S + 2 * P) mypart[tid] = (chunk == 0) ? y0s[tid - S - P] : 0.;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.count + mat, 1) == a.split - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < S) {
    const double* pp = a.part + (size_t)mat * a.split * (S + 2 * P);
    double tot = 0.;
    for (int c = 0; c < a.split; ++c) tot += __ldcg(pp + (size_t)c * (S + 2 * P) + tid);
    int u = 0, v = 0, cnt = 0;
#pragma unroll
    for (int uu = 0; uu < P; ++uu)
#pragma unroll
      for (int vv = uu; vv < P; ++vv) {
        if (cnt == tid) { u = uu; v = vv; }
        ++cnt;
      }
    double col_u = 0.;
    for (int c = 0; c < a.split; ++c) col_u += __ldcg(pp + (size_t)c * (S + 2 * P) + S + u);
    const double y0v = __ldcg(pp + S + P + v);
    // L_p = A W_p(k, M0)/M0 for mass tracers (:560-570); W_p(k, M0) = c_p[0] grid_p[k, 0]
    const double Lu = a.mass[u] ? A * (sc[4 + u] * a.gridW[u][goff]) / M0 : 0.;
    const double Lv = a.mass[v] ? A * (sc[4 + v] * a.gridW[v][goff]) / M0 : 0.;
    if (a.do_I11 && a.mass[u] && a.mass[v]) tot += bk[0] * Lu * Lv;                    // :559-560
    if (a.do_I12_I21 && a.mass[u]) tot += Lu * y0v;                                    // :561-565  beta[0,:] with W_v
    if (a.do_I12_I21 && a.mass[v]) tot += Lv * col_u;                                  // :566-570  beta[:,0] with W_u
    a.inl[((size_t)b * S + tid) * a.nk + ik] = tot * rhom * rhom;                      // :571
  }
}

template <int P>
static int hm_launch_beta(const hm_beta_args& a, int B, cudaStream_t st) {
  const size_t smem = ((size_t)P * a.nMp + 8 * (P * (P + 1) / 2 + P)) * sizeof(double);
  static hm_once_per_device once;
  if (once.needed()) {
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_beta2h_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (smem > 200 * 1024) {
    hm_set_error("beta_NL: %d tracers x %d masses do not fit in shared memory", P, a.nM);
    return HM_ERR_INVALID;
  }
  HM_CUDA_TRY(cudaMemsetAsync(a.count, 0, (size_t)B * a.nk * sizeof(int), st));
  hm_beta2h_kernel<P><<<B * a.nk * a.split, 256, smem, st>>>(a);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Epilogue: one thread per (model, wavenumber)                                          pyhalomodel.py:456-501
// ---------------------------------------------------------------------------------------------------------------
struct hm_epi_args {
  int nk, B, F, P, NW, S, nparts, nchunks;
  int simple_twohalo, subtract_shotnoise, correct_discrete, fence_out;
  int use_cpart;             // the shot-noise / simple-two-halo integrals are needed (and were reduced by the weights kernel)
  int rows_layout;           // partial is [B][nparts][NW][nk] (row-per-lane kernel) instead of [B][nparts][nk][NW]
  double k_trunc;
  const double* k;
  const double* Pk_lin;      // [G][nk]
  const double* partial;
  const double* scal;
  const double* cpart;
  const double* inl;         // [B][S][nk] or NULL
  int discrete[HM_MAXP];
  double* out;               // [B][S][3][nk]
};

// One CTA per (model, 32 wavenumbers).  Phase 1: PG part-groups x 32 rows of threads add the mass parts (coalesced:
// 32 rows x NW values are contiguous); phase 2: one thread per (pair, row) finishes the spectra with coalesced
// stores.  All sums run in a fixed order => bitwise reproducible.
#define HM_EPI_ROWS 32
template <int P>
__global__ void __launch_bounds__(256) hm_epilogue_kernel(const hm_epi_args a) {
  constexpr int S = P * (P + 1) / 2, NW = P + S;
  extern __shared__ double sf[];                       // [PG][NW][32] partial sums, then 2*HM_MAXP chunk sums
  const int PG = blockDim.x / HM_EPI_ROWS;
  double* scs = sf + (size_t)PG * NW * HM_EPI_ROWS;
  const int tid = threadIdx.x, r = tid & 31, pg = tid >> 5;
  const int tiles_per_model = (a.nk + HM_EPI_ROWS - 1) / HM_EPI_ROWS;
  const int b = blockIdx.x / tiles_per_model;
  const int row0 = (blockIdx.x - b * tiles_per_model) * HM_EPI_ROWS;
  const int smp = b / a.F;
  const double rhom = a.scal[(size_t)b * HM_NSCAL + 1];
  const int row = row0 + r;

  double f[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) f[i] = 0.;
  const int rs = a.rows_layout ? 1 : NW, is = a.rows_layout ? a.nk : 1;     // strides of a row and of a value
  if (row < a.nk) {
#pragma unroll 4
    for (int j = pg; j < a.nparts; j += PG) {
      const double* pp = a.partial + ((size_t)b * a.nparts + j) * a.nk * NW + (size_t)row * rs;
#pragma unroll
      for (int i = 0; i < NW; ++i) f[i] += pp[(size_t)i * is];
    }
  }
#pragma unroll
  for (int i = 0; i < NW; ++i) sf[((size_t)pg * NW + i) * HM_EPI_ROWS + r] = f[i];
  if (tid < 2 * HM_MAXP) {     // tid < 8: shot noise of tracer tid; tid >= 8: simple two-halo integral
    const double* cp = a.cpart + (size_t)b * a.nchunks * 2 * HM_MAXP + tid;
    double t = 0.;
    if (a.use_cpart)
      for (int c = 0; c < a.nchunks; ++c) t += cp[(size_t)c * 2 * HM_MAXP];
    scs[tid] = t * rhom;       // P_SN = rhom trapz((amplitude/norm^2) g/M) (:423-426); simple two-halo (:442-446)
  }
  __syncthreads();

  for (int idx = tid; idx < S * HM_EPI_ROWS; idx += blockDim.x) {
    const int s = idx >> 5, rr = idx & 31;            // a warp works on one pair: (u, v) is warp-uniform
    const int krow = row0 + rr;
    if (krow >= a.nk) continue;
    int u = 0, v = 0, i1h = 0;
    {
      int cnt = 0, icross = 2 * P;
#pragma unroll
      for (int uu = 0; uu < P; ++uu)
#pragma unroll
        for (int vv = uu; vv < P; ++vv) {
          if (cnt == s) { u = uu; v = vv; i1h = (uu == vv) ? P + uu : icross; }
          if (uu != vv) ++icross;
          ++cnt;
        }
    }
    double fu = 0., fv = 0., f1 = 0.;
    for (int g = 0; g < PG; ++g) {
      const double* base = sf + (size_t)g * NW * HM_EPI_ROWS + rr;
      fu += base[u * HM_EPI_ROWS];
      fv += base[v * HM_EPI_ROWS];
      f1 += base[i1h * HM_EPI_ROWS];
    }
    const double Iu = a.simple_twohalo ? scs[HM_MAXP + u] : fu * rhom;                         // :540-543
    const double Iv = a.simple_twohalo ? scs[HM_MAXP + v] : fv * rhom;
    const double I_NL = a.inl ? a.inl[((size_t)b * S + s) * a.nk + krow] : 0.;                 // :520-521
    const double p2h = __dmul_rn(__dadd_rn(__dmul_rn(Iu, Iv), I_NL), a.Pk_lin[(size_t)smp * a.nk + krow]);   // :524
    double p1h = f1 * rhom;                                                                    // :532
    if (u == v && a.discrete[u]) {                                                             // :484-491
      if (a.correct_discrete && !a.subtract_shotnoise) p1h += scs[u];
      else if (!a.correct_discrete && a.subtract_shotnoise) p1h -= scs[u];
    }
    if (a.k_trunc > 0.) {                                                                      // :494-495
      const double q = a.k[krow] / a.k_trunc;
      p1h *= 1. - exp(-(q * q));
    }
    double* o = a.out + ((size_t)b * S + s) * 3 * a.nk + krow;
    o[0] = p2h;
    o[a.nk] = p1h;
    o[2 * (size_t)a.nk] = __dadd_rn(p2h, p1h);                                                 // :500-501
  }
  // `out` may live on a peer GPU (sweep.PeerGatherBuffer: the gather is fused into these stores): make them
  // visible system-wide before the kernel retires, so that a barrier after it is enough for the reader.  Fences
  // are cumulative: one system-scope fence by a thread that has synchronised with the whole CTA covers its stores.
  if (a.fence_out) {
    __syncthreads();
    if (tid == 0) __threadfence_system();
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
  HM_REQUIRE(!(pr->beta_dev && pr->simple_twohalo), "a simple two-halo term is not compatible with beta_NL");
  HM_REQUIRE(!pr->beta_dev || pr->n_tracers <= 6, "beta_NL is supported for up to 6 tracers per call");
  const long long B = (long long)pr->n_samples * pr->models_per_sample;
  HM_REQUIRE(B <= 65535, "at most 65535 models per call");
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

static int hm_check_workspace(const char* who, const hm_layout& L, const void* ws, size_t ws_bytes) {
  if (!ws || ws_bytes < L.total_bytes || ((uintptr_t)ws & 15)) {
    hm_set_error("%s: workspace needs %zu bytes, 16-byte aligned (got %zu)", who, L.total_bytes, ws_bytes);
    return HM_ERR_WORKSPACE;
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
  if ((rc = hm_check_workspace("hm_halo_weights", L, ws, ws_bytes))) return rc;
  hm_weights_args a;
  a.nM = pr->nM; a.nMp = L.nMp; a.P = L.P; a.NW = L.NW; a.F = pr->models_per_sample; a.nchunks = L.nchunks;
  a.rows = L.rows; a.NVEC = L.NVEC;
  a.correct_discrete = pr->correct_discrete;
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
  a.cpart = (double*)ws + L.cpart_offset;
  a.lowmass = lowmass_dev;
  dim3 grid(L.nchunks + 1, (unsigned)pr->n_samples);
  const bool red = hm_needs_cpart(pr);
  cudaStream_t st = (cudaStream_t)stream;
#define HM_WLAUNCH(PT, ROWS)                                                    \
  do {                                                                          \
    if (red) hm_weights_kernel<true, PT, ROWS><<<grid, HM_WCHUNK, 0, st>>>(a);  \
    else hm_weights_kernel<false, PT, ROWS><<<grid, HM_WCHUNK, 0, st>>>(a);     \
  } while (0)
  static const bool no_sweep_kernel = getenv("HALOB200_NO_SWEEP_WEIGHTS") != nullptr;      // (A/B switch for timing runs)
  if (L.rows && !red && !no_sweep_kernel) {
    dim3 gs((L.nMp + HM_WS_THREADS * HM_WS_PER - 1) / (HM_WS_THREADS * HM_WS_PER), (unsigned)pr->n_samples);
    hm_weights_sweep_kernel<<<gs, HM_WS_THREADS, 0, st>>>(a);
  } else if (L.rows && L.P <= 2) HM_WLAUNCH(2, true);
  else if (L.rows) HM_WLAUNCH(4, true);
  else if (L.P <= 2) HM_WLAUNCH(2, false);
  else if (L.P <= 4) HM_WLAUNCH(4, false);
  else HM_WLAUNCH(8, false);
#undef HM_WLAUNCH
  HM_LAUNCH_CHECK();
  return HM_OK;
}

template <int P, int R>
static int hm_launch_fallback(hm_quad_args& a, bool vec2, bool sep, cudaStream_t st) {
  a.ntiles = (a.nk + R - 1) / R;
  a.nitems = (long long)a.B * a.ntiles;
  const int ncol = vec2 ? a.nM / 2 : a.nM;
  int threads = ((ncol + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 64) threads = 64;
  long long grid = (long long)hm_num_sms() * (512 / threads);
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

template <int P, int R, int H>
static int hm_launch_stream_h(hm_quad_args& a, cudaStream_t st) {
  using C = hm_stream_cfg<P, R, H>;
  static hm_once_per_device once;
  if (once.needed()) {
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_quad_stream_kernel<P, R, H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
  }
  a.ntiles = (a.nk + R - 1) / R;
  a.nitems = (long long)a.B * a.nslices * a.ntiles;
  long long grid = hm_num_sms();
  if (grid > a.nitems) grid = a.nitems;
  hm_quad_stream_kernel<P, R, H><<<(int)grid, HM_STREAM_THREADS, C::SMEM_BYTES, st>>>(a);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*hm_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static hm_tmap_encode_fn hm_tmap_encoder() {
  static hm_tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (hm_tmap_encode_fn)p;
  }
  return fn;
}

// Row-major [rows][cols] float64 tensor with `pitch` doubles between rows, boxes of box_rows x box_cols.
static int hm_make_tmap_2d(CUtensorMap* tm, const double* base, uint64_t cols, uint64_t rows, uint64_t pitch,
                           uint32_t box_cols, uint32_t box_rows, bool swizzle128) {
  hm_tmap_encode_fn enc = hm_tmap_encoder();
  if (!enc) {
    hm_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return HM_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch * sizeof(double)};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    hm_set_error("cuTensorMapEncodeTiled failed with CUresult %d (%llu x %llu, pitch %llu)", (int)r,
                 (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch);
    return HM_ERR_CUDA;
  }
  return HM_OK;
}

template <int P, int FM>
static int hm_launch_rows(hm_quad_args& a, bool fused, cudaStream_t st) {
  using C = hm_rows_cfg<P, FM>;
  static hm_once_per_device once;
  if (once.needed()) {
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_quad_rows_kernel<P, FM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_quad_rows_kernel<P, FM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     C::SMEM_BYTES));
  }
  a.ntiles = (a.nk + C::RT - 1) / C::RT;
  const int G = a.B / FM;
  a.nitems = (long long)G * a.ntiles;                          // items enumerate (sample, row tile)
  CUtensorMap tmG[HM_ROWS_MAXP], tmW;
  int rc;
  for (int p = 0; p < HM_ROWS_MAXP; ++p)
    if ((rc = hm_make_tmap_2d(&tmG[p], a.gridW[p < P ? p : 0], (uint64_t)a.nM, (uint64_t)G * a.nk, (uint64_t)a.nM, C::BOXC,
                              C::RT, true)))
      return rc;
  if ((rc = hm_make_tmap_2d(&tmW, a.weights, (uint64_t)a.nMp, (uint64_t)G * C::NVEC, (uint64_t)a.nMp, C::CS, C::NVEC, false)))
    return rc;
  long long grid = hm_num_sms();
  if (grid > a.nitems) grid = a.nitems;
  HM_CUDA_TRY(cudaMemsetAsync(a.item_counter, 0, sizeof(int), st));
  if (fused) hm_quad_rows_kernel<P, FM, true><<<(int)grid, HM_ROWS_THREADS, C::SMEM_BYTES, st>>>(a, tmG[0], tmG[1], tmG[2], tmW);
  else hm_quad_rows_kernel<P, FM, false><<<(int)grid, HM_ROWS_THREADS, C::SMEM_BYTES, st>>>(a, tmG[0], tmG[1], tmG[2], tmW);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

template <int P, int R>
static int hm_launch_stream(hm_quad_args& a, bool rows, bool fused, cudaStream_t st) {
  // two to five models per sample on one to three tracers (the five mass functions of BASELINE config 3) take the
  // row-per-lane kernel, which reads the shared grids once; other combinations read them once per model
  if constexpr (P <= HM_ROWS_MAXP) {
    if (rows) switch (a.F) {
      case 2: return hm_launch_rows<P, 2>(a, fused, st);
      case 3: return hm_launch_rows<P, 3>(a, fused, st);
      case 4: return hm_launch_rows<P, 4>(a, fused, st);
      case 5: return hm_launch_rows<P, 5>(a, fused, st);
    }
  }
  if constexpr (P <= 3) {
    if (a.MC == 2 * HM_MC) return hm_launch_stream_h<P, R, 2>(a, st);
  }
  return hm_launch_stream_h<P, R, 1>(a, st);
}

template <int P>
static int hm_launch_epilogue(const hm_epi_args& e, cudaStream_t st) {
  constexpr int NW = P + P * (P + 1) / 2;
  // part-groups per CTA: small enough (<= 9 KB of shared memory) that epilogue CTAs co-reside with the persistent
  // streaming CTAs of the NEXT step when consecutive steps are pipelined over two streams
  const int PG = (NW <= 9) ? 4 : (NW <= 20) ? 2 : 1;
  const int tiles = (e.nk + HM_EPI_ROWS - 1) / HM_EPI_ROWS;
  const size_t smem = ((size_t)PG * NW * HM_EPI_ROWS + 2 * HM_MAXP) * sizeof(double);
  hm_epilogue_kernel<P><<<e.B * tiles, PG * HM_EPI_ROWS, smem, st>>>(e);
  HM_LAUNCH_CHECK();
  return HM_OK;
}

extern "C" int hm_mass_quadrature_stages(const hm_problem* pr, double* out_dev, const void* ws, size_t ws_bytes,
                                         void* stream, int stages) {
  int rc = hm_check_device();
  if (rc) return rc;
  if ((rc = hm_validate(pr))) return rc;
  HM_REQUIRE(out_dev, "out_dev is NULL");
  const hm_layout L = hm_make_layout(pr);
  if ((rc = hm_check_workspace("hm_mass_quadrature", L, ws, ws_bytes))) return rc;
  hm_quad_args a;
  a.nk = pr->nk; a.nM = pr->nM; a.nMp = L.nMp; a.F = pr->models_per_sample; a.B = (int)L.B; a.nslices = L.nslices;
  a.nparts = L.nparts; a.MC = L.MC;
  bool sep = false, aligned = (pr->nM % 2 == 0);
  for (int p = 0; p < HM_MAXP; ++p) {
    const bool on = p < L.P;
    const hm_tracer& t = pr->tracers[p];
    a.gridU[p] = on ? t.grid_dev : nullptr;
    a.gridW[p] = on ? (t.grid_mode == HM_GRID_SEPARATE ? t.wk_dev : t.grid_dev) : nullptr;
    if (on && t.grid_mode == HM_GRID_SEPARATE) sep = true;
    if (on && (((uintptr_t)a.gridU[p] | (uintptr_t)a.gridW[p]) & 15)) aligned = false;
  }
  a.weights = (const double*)ws;
  a.partial = (double*)ws + L.partial_offset;
  // sweeps under default flags: the row-per-lane kernel finishes the spectra itself (no partial sums, no epilogue launch)
  static const bool allow_fused = getenv("HALOB200_NO_FUSED_EPILOGUE") == nullptr;      // (A/B switch for timing runs)
  const bool fused = allow_fused && L.rows && !pr->simple_twohalo && !(pr->k_trunc > 0.) && !pr->beta_dev &&
                     !hm_needs_cpart(pr) && !pr->output_is_peer;
  a.scal = (const double*)ws + L.scal_offset; a.Pk_lin = pr->Pk_lin_dev; a.out = out_dev;
  a.item_counter = (int*)((double*)ws + L.counter_offset);
  hm_epi_args e;
  e.nk = pr->nk; e.B = (int)L.B; e.F = pr->models_per_sample; e.P = L.P; e.NW = L.NW; e.S = L.S;
  e.nparts = L.nparts; e.nchunks = L.nchunks + 1; e.rows_layout = L.rows; e.use_cpart = hm_needs_cpart(pr);
  e.simple_twohalo = pr->simple_twohalo; e.subtract_shotnoise = pr->subtract_shotnoise;
  e.correct_discrete = pr->correct_discrete; e.k_trunc = pr->k_trunc; e.k = pr->k_dev; e.Pk_lin = pr->Pk_lin_dev;
  e.fence_out = pr->output_is_peer;
  e.partial = a.partial; e.scal = (const double*)ws + L.scal_offset; e.cpart = (const double*)ws + L.cpart_offset;
  e.inl = pr->beta_dev ? (const double*)ws + L.inl_offset : nullptr;
  hm_beta_args bt;
  bt.nk = pr->nk; bt.nM = pr->nM; bt.nMp = L.nMp; bt.F = pr->models_per_sample;
  bt.do_I11 = pr->beta_do_I11; bt.do_I12_I21 = pr->beta_do_I12_I21; bt.M = pr->M_dev; bt.beta = pr->beta_dev;
  bt.vec2 = (pr->nM % 2 == 0) && !((uintptr_t)pr->beta_dev & 15);
  for (int p = 0; p < HM_MAXP; ++p) { bt.gridW[p] = a.gridW[p]; bt.mass[p] = (p < L.P) ? pr->tracers[p].mass_tracer : 0; }
  bt.weights = a.weights; bt.scal = e.scal; bt.inl = (double*)ws + L.inl_offset;
  bt.split = L.bsplit; bt.part = (double*)ws + L.bpart_offset; bt.count = (int*)((double*)ws + L.bcount_offset);
  for (int p = 0; p < HM_MAXP; ++p) e.discrete[p] = (p < L.P) ? pr->tracers[p].discrete_tracer : 0;
  e.out = out_dev;
  cudaStream_t st = (cudaStream_t)stream;
#define HM_CASE(PP, RR)                                                            \
  case PP:                                                                         \
    if (stages & 1) {                                                              \
      rc = L.stream ? hm_launch_stream<PP, RR>(a, L.rows, fused, st) : hm_launch_fallback<PP, RR>(a, aligned, sep, st); \
      if (rc) return rc;                                                           \
      if (pr->beta_dev && (rc = hm_launch_beta<PP>(bt, (int)L.B, st))) return rc;  \
    }                                                                              \
    return ((stages & 2) && !fused) ? hm_launch_epilogue<PP>(e, st) : HM_OK;
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

extern "C" int hm_mass_quadrature(const hm_problem* pr, double* out_dev, const void* ws, size_t ws_bytes,
                                  void* stream) {
  return hm_mass_quadrature_stages(pr, out_dev, ws, ws_bytes, stream, 3);
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
