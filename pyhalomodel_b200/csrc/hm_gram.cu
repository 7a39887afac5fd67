// hm_gram.cu -- many-tracer halo-model spectra: the one-halo cross terms as a per-wavenumber Gram contraction on
// the FP64 tensor cores (BASELINE config 4: 64 HOD tracers, 2080 unique pairs per k).
//
// For P tracers the cross one-halo integrals of model.power_spectrum (pyhalomodel.py:479, :526-533) are
//     P1h_uv(k) = rhom * sum_m w1[m] * W_u(k,m) * W_v(k,m),      W_p(k,m) = c_p[m] * grid_p[k,m]
// i.e. per k the upper triangle of (w1 o W) W^T, a P x P x nM product.  With P(P+1)/2 pairs the arithmetic
// intensity (~6 flop/B at P=64) is beyond what the streaming-quadrature kernel's per-pair register accumulators
// can hold, so this path tiles it for mma.sync.m8n8k4.f64 (SASS DMMA; there is no tcgen05 FP64 MMA):
//
//   hm_gram_weights_kernel  w1, w2 and r2 = w2/sqrt(w1) per mass; A; per (tracer, mass) s_p = sqrt(w1) c_p (the MMA operand
//                           is Y = s_p*grid, so that the contraction needs no weight: sum_m w1 W_u W_v = sum_m Y_u Y_v
//                           with W = c_p*grid) and d_p (auto one-halo weight)
//   hm_gram_scalars_kernel  shot noise and simple-two-halo integrals per (tracer, model); non-default flags only
//   hm_gram_kernel<NBT>     persistent warp-specialised CTA per SM over equal contiguous ranges of (group of 4 wavenumbers,
//                           32-mass chunk) units: producer warps stream the raw tracer rows with cp.async into a three-
//                           stage ring (mbarrier completion); each pair of MMA warps owns one wavenumber's whole upper
//                           triangle (18 8x8 DMMA accumulator blocks per warp), scales its fragments in registers and
//                           keeps the diagonal work (two-halo sums, auto one-halo sums) on them.
//   hm_gram_epilogue_kernel adds the partial results of the CTAs that shared a group and finishes P_2h, P_1h, P_hm for
//                           every unique pair (:456-501)
//
// The diagonal (auto) terms are NOT Gram entries: they use Uk, amplitude, variance (pyhalomodel.py:465-477).
#include "hm_common.cuh"

#define HM_GMAXP HM_GRAM_MAX_TRACERS
#ifndef HM_GMCH
#define HM_GMCH 32                 // masses per chunk = one stage of the ring (256-byte row segments)
#endif
#define HM_GRK 4                   // wavenumbers per group; two MMA warps share a wavenumber
#if defined(HM_GRAM_SIXTEEN_WARPS) || defined(HM_GRAM_THREE_WAY)
// (layout experiments) SIXTEEN_WARPS: 8 DMMA-only warps + 4 diagonal-work warps + 4 producer warps; THREE_WAY: three MMA
// warps per wavenumber, a third of its triangle (12 accumulator blocks) and of its diagonal work each, + 4 producer warps
#define HM_GMMA_WARPS (3 * HM_GRK)
#else
#define HM_GMMA_WARPS (2 * HM_GRK)
#endif
#ifndef HM_GNS
#define HM_GNS 3                   // stages (64 KB each at 64 tracers)
#endif
#define HM_GPROD_WARPS 4
#define HM_GTHREADS (32 * (HM_GMMA_WARPS + HM_GPROD_WARPS))
#ifdef HM_GRAM_PRESCALE            // the producer warps also scale the staged tile in place and keep the diagonal work
#define HM_GPROD_REGS 112
#define HM_GMMA_REGS 192
#else
#define HM_GPROD_REGS 40
#define HM_GMMA_REGS 232           // (split-triangle layout)
#endif
#define HM_GMMA_ALL_REGS 248       // whole-triangle MMA warps: 144 accumulator registers
#define HM_GDIAG_REGS 152          // diagonal-work warps
#define HM_G16_MMA_REGS 160        // sixteen-warp layout: launch at 128, DMMA warps 160, diagonal-work warps 112
#define HM_G16_DIAG_REGS 112
#ifndef HM_G3_MMA_REGS
#define HM_G3_MMA_REGS 152         // three-way layout: launch at 128 (512 threads), producers 40, MMA warps 152
#endif
#define HM_GNSCAL 4
#define HM_GNTS 3                  // per (model, tracer) scalars: shot noise, simple two-halo integral, c_p[M0]
#define HM_GREC(PTK) (2 * (2 * (PTK) + 1))     // doubles per mass-pair record of the factor table: s [PTK], d [PTK], r2, x 2 masses

struct hm_gram_layout {
  int P, NB, PT, PTK, S, NW, nMp, nchunks, ngroups, CH, ncta, J;
  long long total;
  size_t B, gw_stride, scal_offset, tscal_offset, partial_offset, total_bytes;
};

static hm_gram_layout hm_gram_make_layout(const hm_gram_problem* pr) {
  hm_gram_layout L;
  L.P = pr->n_tracers;
  L.NB = (L.P + 7) / 8;
  L.PT = 8 * L.NB;
  L.PTK = (L.NB <= 2) ? 16 : (L.NB <= 4) ? 32 : 64;      // tracer rows of the kernel instance that runs (hm_gram_kernel<NBT>)
  L.S = L.P * (L.P + 1) / 2;
  L.NW = L.P + L.S;                       // [2h sums P][auto 1h P][cross 1h P(P-1)/2]
  L.nMp = (pr->nM + 1) & ~1;
  L.nchunks = (L.nMp + 255) / 256;
  L.B = (size_t)pr->n_samples * (size_t)pr->models_per_sample;
  // units of the Gram kernel (see hm_gram_args): equal contiguous ranges per persistent CTA; a group of HM_GRK
  // wavenumbers is then shared by at most J CTAs, each with its own partial-result slot
  L.ngroups = (pr->nk + HM_GRK - 1) / HM_GRK;
  L.CH = (pr->nM + HM_GMCH - 1) / HM_GMCH;
  L.total = (long long)L.B * L.ngroups * L.CH;
  L.ncta = (int)(L.total < hm_num_sms() ? L.total : hm_num_sms());
  const long long minu = L.total / L.ncta;               // >= 1
  L.J = (int)((L.CH + minu - 1) / minu) + 1;
  auto align = [](size_t x) { return (x + 31) & ~(size_t)31; };
  // per model: the vectors w1, w2, then one record per PAIR of masses with the factors the Gram kernel's lanes read:
  // s_p [PTK], d_p [PTK], r2 as double2 over the two masses (HM_GREC doubles).  A lane's factors of a mass pair then sit
  // at compile-time offsets from ONE address (tracer 8 I + fr is 8 I records on), instead of eight row pointers that the
  // compiler rebuilt from the thread index for every 8-mass block (18 % of an MMA warp's stall samples).
  L.gw_stride = (size_t)2 * L.nMp + (size_t)(L.nMp / 2) * HM_GREC(L.PTK);
  L.scal_offset = align(L.B * L.gw_stride);
  L.tscal_offset = align(L.scal_offset + L.B * HM_GNSCAL);
  L.partial_offset = align(L.tscal_offset + L.B * (size_t)L.P * HM_GNTS);
  L.total_bytes = (L.partial_offset + L.B * (size_t)L.ngroups * L.J * L.NW * HM_GRK) * sizeof(double);
  return L;
}

// ---------------------------------------------------------------------------------------------------------------
struct hm_gram_wargs {
  int nM, nMp, P, PTK, F;
  int correct_discrete;
  int need_integrals;                 // the shot-noise / simple two-halo integrals enter the spectra (non-default flags)
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

// grid (mass chunk, 1 + PTK, model): CTA row 0 writes w1, w2, r2 and integrates A; CTA row 1 + p writes tracer p's factors
// s_p = sqrt(w1) c_p (W = c_p * grid, Y = s_p * grid) and d_p (auto one-halo weight) into the mass-pair records (zeros for
// the padding rows p >= P).  Every row recomputes w1 for its masses (one g(nu) per thread, identical code => identical
// bits), which keeps the whole preparation in ONE launch.
__global__ void __launch_bounds__(256) hm_gram_weights_kernel(const hm_gram_wargs a) {
  __shared__ double scratch[32];
  const int b = blockIdx.z, chunk = blockIdx.x, tid = threadIdx.x;
  const int role = blockIdx.y;                       // 0: mass-function weights; 1 + p: tracer p
  const int s = b / a.F;
  const hm_model_desc d = a.models ? a.models[b] : a.single;
  const double* sig = a.sigma + (size_t)s * a.nM;
  if (role == 0 && chunk == 0) {   // A = 1 - int g b dnu (pyhalomodel.py:413-414)
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
  double* rec = gw + (size_t)2 * a.nMp + (size_t)(m >> 1) * HM_GREC(a.PTK) + (m & 1);      // my mass in its pair's record
  if (role > a.P) {                                  // padding tracer row of the kernel instance: zero factors
    rec[2 * (role - 1)] = 0.;
    rec[2 * (a.PTK + role - 1)] = 0.;
    return;
  }
  double w1 = 0., w2 = 0.;
  if (m < a.nM) {
    const double nu = d.dc / sig[m];
    const double nu_hi = (m + 1 < a.nM) ? d.dc / sig[m + 1] : nu;
    const double nu_lo = (m > 0) ? d.dc / sig[m - 1] : nu;
    const double tw = 0.5 * (nu_hi - nu_lo);
    const double g = hm_g_nu(d, nu);
    const double Mm = a.M[m];
    w1 = tw * (g / Mm);             // one-halo weight (:530)
    if (role == 0) w2 = tw * ((hm_b_nu(d, nu) * g) / Mm);    // two-halo weight (:539)
  }
  if (role == 0) {
    gw[m] = w1;
    gw[(size_t)a.nMp + m] = w2;
    rec[2 * (2 * a.PTK)] = (w1 > 0.) ? w2 / sqrt(w1) : 0.;                  // r2: sum_m w2 W_p = sum_m r2 Y_p  (:539)
    return;
  }
  const int p = role - 1;
  double c = 0., dw = 0.;
  if (m < a.nM) {
    const size_t aoff = a.amp_stride_is_grid[p] ? (size_t)s * a.grid_sample_stride : (size_t)s * a.nM;
    const double amp = a.amplitude[p][aoff + m];
    const double norm = a.norm[p][b];
    double fac = (a.correct_discrete && a.discrete[p]) ? __dmul_rn(amp, __dadd_rn(amp, -1.)) : __dmul_rn(amp, amp);
    if (a.variance[p]) fac = __dadd_rn(fac, a.variance[p][(size_t)s * a.nM + m]);   // :466-473
    double uS;
    if (a.mode[p] == HM_GRID_UK) { c = amp / norm; uS = 1. / norm; }
    else { c = 1.; uS = 1. / (amp * norm); }                                        // HM_GRID_WK
    dw = __dmul_rn(__dmul_rn(w1, fac), uS * uS);                                    // d_p = w1 fac uS^2 (:475-477)
    if (m == 0 && !a.need_integrals) {            // default flags: only c_p[M0] is read from the per-tracer scalars
      double* t = a.tscal + ((size_t)b * a.P + p) * HM_GNTS;
      t[0] = t[1] = 0.;
      t[2] = c;                                   // the low-mass term of the epilogue
    }
  }
  rec[2 * p] = (m < a.nM) ? sqrt(w1) * c : 0.;
  rec[2 * (a.PTK + p)] = dw;
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
    double* t = a.tscal + ((size_t)b * a.P + p) * HM_GNTS;
    t[0] = a.discrete[p] ? sn * rhom : 0.;
    t[1] = si * rhom;
    t[2] = (a.mode[p] == HM_GRID_UK) ? a.amplitude[p][aoff] / norm : 1.;      // c_p[M0]: low-mass term of the epilogue
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Work decomposition of the Gram kernel.  A "group" is HM_GRK consecutive wavenumbers of one model, a "unit" one
// 16-mass chunk of a group; the B * ngroups * CH units are cut into ncta equal contiguous ranges (one per persistent
// CTA, stream-K style), so every SM gets the same number of chunks whatever the shape.  A CTA's range touches a few
// groups; it writes one partial result per group it touches into slot j = cta - (first CTA of the group), and the
// epilogue adds a group's slots in CTA order (fixed => bitwise reproducible).
// ---------------------------------------------------------------------------------------------------------------
struct hm_gram_args {
  int nk, nM, nMp, P, NW, F, ngroups, CH, ncta, J;
  long long total;          // units
  size_t gw_stride;
  const double* grid[HM_GMAXP];
  const double* gw;
  double* partial;          // [B * ngroups][J][NW][HM_GRK]
};

__host__ __device__ __forceinline__ long long hm_gram_ubeg(long long c, long long total, int ncta) { return c * total / ncta; }
__host__ __device__ __forceinline__ int hm_gram_cta_of(long long unit, long long total, int ncta) {
  int c = (int)(unit * ncta / total);
  while (c + 1 < ncta && hm_gram_ubeg(c + 1, total, ncta) <= unit) ++c;
  while (c > 0 && hm_gram_ubeg(c, total, ncta) > unit) --c;
  return c;
}

__device__ __forceinline__ void hm_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Persistent, warp-specialised CTA (one per SM, 12 warps):
//   producer warps 8..11  stream the raw tracer rows of the next chunks -- HM_GRK wavenumbers x PT tracers x 32 masses,
//                         64 KB at 64 tracers -- with cp.async (LDGSTS, 256-byte row segments) into a ring of three
//                         stages; completion is signalled on the stage's mbarrier by the copies themselves
//                         (cp.async.mbarrier.arrive.noinc), so the role never waits for data.  They give most of their
//                         registers back first (setmaxnreg), which lets the MMA warps run at 232.
//   MMA warps 0..7        warps 2r and 2r + 1 own wavenumber r of the group and one half each of the upper triangle of
//                         its Gram matrix: block rows are paired (I, NBT - 1 - I) -- NBT + 1 blocks per pair -- and the
//                         pairs dealt alternately, 18 8x8 accumulator blocks per warp at 64 tracers.  The m8n8k4 A
//                         fragment of block row I and the B fragment of block column I are the same register (lane
//                         (fr, fk) holds Y[8 I + fr][mass fk]), so per four masses eight fragments feed 18 DMMAs, and
//                         there is no second pass over shared memory: the staged tile is the raw grid, and the lane
//                         scales its own fragments, Y = s_p * grid with s_p = sqrt(w1) c_p from L1 (the factors are
//                         shared by the eight warps), and keeps the diagonal work of its block rows (two-halo sums
//                         sum_m r2 Y_p, auto one-halo sums sum_m d_p grid^2) on them as well.
// Within an 8-mass block the k4-step "x" takes masses {0,2,4,6} and "y" {1,3,5,7} (16-byte fragment loads; any
// assignment works as long as A and B agree: the sum over masses is order-free).  Shared-memory rows are 256 bytes (32
// masses); the 16-byte slot index is XORed with 4 * (tracer & 1), which makes the fragment loads of a quarter warp
// (two tracer rows x four slots) conflict-free without padding.
template <int NBT, int H>
struct hm_gram_own {       // what MMA-group warp role H does: H = 0, 1: half of the triangle and its diagonal work (block rows
                           // paired (I, NBT - 1 - I), pairs dealt alternately); H = 2: the whole triangle, no diagonal work;
                           // H = 3: the diagonal work of every block row, no DMMA
  __host__ __device__ static constexpr bool half(int I) { return (((I < NBT - 1 - I) ? I : NBT - 1 - I) & 1) == (H & 1); }
  // H = 6, 7, 8: a third of the triangle and its diagonal work -- block rows {0, 4}, {1, 3}, {2, 5, 6, 7} of eight (12 blocks
  // each), {0}, {1}, {2, 3} of four, {0}, {1}, {} of two
  __host__ __device__ static constexpr bool third(int I) {
    return NBT == 8 ? (H == 6 ? (I == 0 || I == 4) : H == 7 ? (I == 1 || I == 3) : (I == 2 || I >= 5))
                    : (H == 6 ? I == 0 : H == 7 ? I == 1 : I >= 2);
  }
  __host__ __device__ static constexpr bool mma(int I) {                                                       // H = 4, 5: half
    return H >= 6 ? third(I) : (H == 2 || ((H < 2 || H >= 4) && half(I)));
  }
  __host__ __device__ static constexpr bool diag(int I) { return H >= 6 ? third(I) : (H == 3 || (H < 2 && half(I))); }   // triangle, no diagonal work
  __host__ __device__ static constexpr int nblocks() {
    int n = 0;
    for (int I = 0; I < NBT; ++I)
      if (mma(I)) n += NBT - I;
    return n;
  }
  __host__ __device__ static constexpr int nrows() {
    int n = 0;
    for (int I = 0; I < NBT; ++I)
      if (diag(I)) ++n;
    return n;
  }
};

template <int NBT, int H>
__device__ __forceinline__ void hm_gram_mma_role(const hm_gram_args& a, const unsigned char* gsm, uint32_t bar_full,
                                                 uint32_t bar_empty, long long u0, long long u1, int krow, int lane) {
  using Own = hm_gram_own<NBT, H>;
  constexpr int PT = 8 * NBT;
  constexpr int STAGE_BYTES = HM_GRK * PT * HM_GMCH * 8;
  constexpr int NB = Own::nblocks() > 0 ? Own::nblocks() : 1, NR = Own::nrows() > 0 ? Own::nrows() : 1;
  const int P = a.P;
  const int fr = lane >> 2, fk = lane & 3;
  const uint32_t swz = (uint32_t)(fr & 1) << 2;
  double acc[NB][2], a2h[NR], a1h[NR];
#pragma unroll
  for (int t = 0; t < NB; ++t) acc[t][0] = acc[t][1] = 0.;
#pragma unroll
  for (int i = 0; i < NR; ++i) a2h[i] = a1h[i] = 0.;
  int s = 0;
  uint32_t parity = 0;
  // The per-(tracer, mass) factors come from L1/L2 and do not depend on the ring: they are loaded ONE 8-mass block
  // ahead (also across chunk boundaries), so their latency hides behind the DMMAs of the current block.  The loads
  // are unconditional -- a branch per tracer row would serialise their latencies: out-of-range masses read a clamped
  // (finite) record and padding tracers zero entries, and since the staged grid values are zero-filled there, every
  // product below is an exact zero.
  // Index arithmetic is kept off the critical path: one record address per block with the tracer rows at immediate
  // offsets, and (group, chunk) cursors advanced by increments instead of 64-bit divisions per unit.
  double2 Sn[NBT], Dn[NR], r2n;
  const int recbase = 2 * a.nMp + 2 * fr;               // (doubles) first record of a model's factor table, my tracer row
  long long G = u0 / a.CH;                              // group of the current unit, its chunk and its model's factors
  int ch = (int)(u0 - G * a.CH);
  int gin = (int)(G % a.ngroups);
  const double* __restrict__ gwb = a.gw + (size_t)(G / a.ngroups) * a.gw_stride;
  const int nunits = (int)(u1 - u0);
  auto load_factors = [&](const double* __restrict__ gw, int chn, int blk) {
    const int mm = chn * HM_GMCH + 8 * blk + 2 * fk;
    const int mc = (mm < a.nMp - 2) ? mm : a.nMp - 2;
    // the record of my mass pair: s of tracer 8 I + fr at entry 8 I, d at PT + 8 I, r2 at 2 PT (from my row: 2 PT - fr)
    const double2* rec = reinterpret_cast<const double2*>(gw + (recbase + (mc >> 1) * HM_GREC(PT)));
    if (Own::nrows() > 0) r2n = rec[2 * PT - fr];
    int i = 0;
#pragma unroll
    for (int I = 0; I < NBT; ++I) {
      Sn[I] = rec[8 * I];
      if (Own::diag(I)) Dn[i++] = rec[PT + 8 * I];
    }
  };
#ifdef HM_GRAM_PRESCALE            // the staged tile already holds Y = s_p * grid, and the producers keep the diagonal sums
  constexpr bool PRE = true;
#else
  constexpr bool PRE = false;
#endif
#ifdef HM_GRAM_THREE_WAY           // no room for a block of factors in flight at 152 registers: load them with the fragments
  constexpr bool PREFETCH = false;
#else
  constexpr bool PREFETCH = !PRE;
#endif
  if (PREFETCH) load_factors(gwb, ch, 0);
  for (int n = 0; n < nunits; ++n) {
    // the unit after this one (for the factor prefetch and the end-of-group test)
    int chN = ch + 1, ginN = gin;
    const double* __restrict__ gwN = gwb;
    if (chN == a.CH) {
      chN = 0;
      if (++ginN == a.ngroups) { ginN = 0; gwN += a.gw_stride; }
    }
    const bool last_unit = n == nunits - 1;
    hm_mbar_wait(bar_full + 8 * s, parity);             // the chunk has landed
    const unsigned char* tile = gsm + (size_t)s * STAGE_BYTES + (size_t)krow * PT * (HM_GMCH * 8) + (size_t)fr * (HM_GMCH * 8);
#ifdef HM_GRAM_SWPIPE
    // (experiment) software pipeline inside a chunk: the fragments of block b + 1 are loaded before the DMMAs of block b and
    // scaled BETWEEN them (one fragment every 36 / 8 DMMAs), so that the scaling and diagonal instructions queue behind
    // this warp's own DMMAs -- which it would have to wait for anyway -- instead of forming a phase of their own
    {
      constexpr int NBLK = HM_GMCH / 8, NDM = 2 * NB;
      double2 Y[NBT];
      auto load_x = [&](int blk, double2 (&X)[NBT]) {
#pragma unroll
        for (int I = 0; I < NBT; ++I)
          X[I] = *reinterpret_cast<const double2*>(tile + (size_t)I * 8 * (HM_GMCH * 8) + ((((uint32_t)(4 * blk + fk)) ^ swz) << 4));
      };
      auto scale_frag = [&](int I, double2 (&X)[NBT]) {      // X[I] -> Y = s_p * grid in place, after the diagonal work on it
        int i = 0;
#pragma unroll
        for (int J = 0; J < NBT; ++J)
          if (J < I && Own::diag(J)) ++i;
        const double2 x = X[I];
        const double2 y = make_double2(Sn[I].x * x.x, Sn[I].y * x.y);
        if (Own::diag(I)) {
          a2h[i] = fma(r2n.x, y.x, fma(r2n.y, y.y, a2h[i]));
          a1h[i] = fma(Dn[i].x * x.x, x.x, fma(Dn[i].y * x.y, x.y, a1h[i]));
        }
        X[I] = y;
      };
      {
        double2 X0[NBT];
        load_x(0, X0);
#pragma unroll
        for (int I = 0; I < NBT; ++I) { scale_frag(I, X0); Y[I] = X0[I]; }
      }
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        const bool more = blk < NBLK - 1;
        double2 Xn[NBT];
        if (more) { load_factors(gwb, ch, blk + 1); load_x(blk + 1, Xn); }
        else if (!last_unit) load_factors(gwN, chN, 0);
        int cnt = 0, kf = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          int t = 0;
#pragma unroll
          for (int I = 0; I < NBT; ++I)
            if (Own::mma(I)) {
#pragma unroll
              for (int Jb = I; Jb < NBT; ++Jb) {
                hm_dmma884(acc[t][0], acc[t][1], half ? Y[I].y : Y[I].x, half ? Y[Jb].y : Y[Jb].x);
                ++t; ++cnt;
                if (more && kf < NBT && cnt == ((kf + 1) * NDM) / NBT) { scale_frag(kf, Xn); ++kf; }
              }
            }
        }
        if (more) {
#pragma unroll
          for (int I = 0; I < NBT; ++I) Y[I] = Xn[I];
        }
      }
    }
#else
#ifndef HM_GRAM_BLK_UNROLL
#define HM_GRAM_BLK_UNROLL 1
#endif
    constexpr int BLKU = HM_GRAM_BLK_UNROLL;
#pragma unroll BLKU
    for (int blk = 0; blk < HM_GMCH / 8; ++blk) {
      double2 X[NBT], Y[NBT];
      if (!PREFETCH && !PRE) load_factors(gwb, ch, blk);
#pragma unroll
      for (int I = 0; I < NBT; ++I)
        X[I] = *reinterpret_cast<const double2*>(tile + (size_t)I * 8 * (HM_GMCH * 8) + ((((uint32_t)(4 * blk + fk)) ^ swz) << 4));
      {
        int i = 0;
#pragma unroll
        for (int I = 0; I < NBT; ++I) {
#ifdef HM_GRAM_DEBUG_NO_SCALE      // (timing experiment, wrong results: the fragments go to the DMMAs unscaled)
          Y[I] = X[I];
#else
          Y[I] = PRE ? X[I] : make_double2(Sn[I].x * X[I].x, Sn[I].y * X[I].y);                  // Y = s_p * grid
#endif
#ifdef HM_GRAM_DEBUG_NO_DIAG       // (timing experiment, wrong results: no two-halo / auto one-halo sums)
          if (false) {
#else
          if (!PRE && Own::diag(I)) {
#endif
            a2h[i] = fma(r2n.x, Y[I].x, fma(r2n.y, Y[I].y, a2h[i]));                             // sum_m r2 Y_p = sum_m w2 W_p  (:539)
            a1h[i] = fma(Dn[i].x * X[I].x, X[I].x, fma(Dn[i].y * X[I].y, X[I].y, a1h[i]));       // sum_m d_p U^2  (:475-477)
            ++i;
          }
        }
      }
      // the factors have been used: fetch those of the next block (or of the first block of the next chunk) behind the
      // DMMAs below
      if (PREFETCH) {
        if (blk < HM_GMCH / 8 - 1) load_factors(gwb, ch, blk + 1);
        else if (!last_unit) load_factors(gwN, chN, 0);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        int t = 0;
#pragma unroll
        for (int I = 0; I < NBT; ++I)
          if (Own::mma(I)) {
#pragma unroll
            for (int Jb = I; Jb < NBT; ++Jb) {
              hm_dmma884(acc[t][0], acc[t][1], half ? Y[I].y : Y[I].x, half ? Y[Jb].y : Y[Jb].x);
              ++t;
            }
          }
      }
    }
#endif
    __syncwarp();
    if (lane == 0) hm_mbar_arrive(bar_empty + 8 * s);   // my reads of the stage have been consumed
    if (++s == HM_GNS) { s = 0; parity ^= 1u; }

    if (ch == a.CH - 1 || last_unit) {                  // the group ends here for this CTA: write my part of its partial result
      const int j = (int)blockIdx.x - hm_gram_cta_of(G * a.CH, a.total, a.ncta);
      double* dst = a.partial + ((size_t)G * a.J + j) * (size_t)a.NW * HM_GRK + krow;
      // C fragment layout (m8n8k4.f64): c0/c1 = C[8 I + lane/4][8 J + 2 (lane%4) + {0,1}]
      int t = 0, i = 0;
#pragma unroll
      for (int I = 0; I < NBT; ++I) {
        if (Own::mma(I)) {
#pragma unroll
          for (int Jb = I; Jb < NBT; ++Jb) {
            const int uu = 8 * I + fr;
#pragma unroll
            for (int ee = 0; ee < 2; ++ee) {
              const int vv = 8 * Jb + 2 * fk + ee;
              if (uu < vv && vv < P) {                  // strict upper triangle, row-major index of (u, v)
                const size_t idx = (size_t)2 * P + (size_t)uu * P - (size_t)uu * (uu + 1) / 2 + (vv - uu - 1);
                dst[idx * HM_GRK] = acc[t][ee];
              }
              acc[t][ee] = 0.;
            }
            ++t;
          }
        }
        if (!PRE && Own::diag(I)) {
          // the four lanes of a tracer row hold its masses 2 fk, 2 fk + 1 (mod 8)
          double v2 = a2h[i], v1 = a1h[i];
          v2 += __shfl_xor_sync(0xffffffffu, v2, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
          v2 += __shfl_xor_sync(0xffffffffu, v2, 2); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
          const int p = 8 * I + fr;
          if (fk == 0 && p < P) {
            dst[(size_t)p * HM_GRK] = v2;
            dst[(size_t)(P + p) * HM_GRK] = v1;
          }
          a2h[i] = a1h[i] = 0.;
          ++i;
        }
      }
    }
    if (chN == 0) ++G;
    ch = chN; gin = ginN; gwb = gwN;
  }
}

template <int NBT>
__global__ void __launch_bounds__(HM_GTHREADS, 1) hm_gram_kernel(const hm_gram_args a) {
  extern __shared__ __align__(128) unsigned char gsm[];
  constexpr int PT = 8 * NBT;
  constexpr int STAGE_BYTES = HM_GRK * PT * HM_GMCH * 8;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + (size_t)HM_GNS * STAGE_BYTES);
  const uint32_t ring = hm_smem_addr(gsm), bar_full = hm_smem_addr(bars), bar_empty = hm_smem_addr(bars + HM_GNS);
#ifdef HM_GRAM_PRESCALE
  const uint32_t bar_raw = hm_smem_addr(bars + 2 * HM_GNS);   // the raw tile has landed (cp.async); bar_full: it has been scaled
#else
  const uint32_t bar_raw = bar_full;
#endif
  const int P = a.P;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform: role dispatch does not diverge
  if (tid == 0) {
    for (int s = 0; s < HM_GNS; ++s) {
      hm_mbar_init(bar_full + 8 * s, 32 * HM_GPROD_WARPS);    // one (cp.async-driven) arrival per producer thread
      hm_mbar_init(bar_empty + 8 * s, HM_GMMA_WARPS);
#ifdef HM_GRAM_PRESCALE
      hm_mbar_init(bar_raw + 8 * s, 32 * HM_GPROD_WARPS);
#endif
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long u0 = hm_gram_ubeg(blockIdx.x, a.total, a.ncta), u1 = hm_gram_ubeg(blockIdx.x + 1, a.total, a.ncta);

  if (warp >= HM_GMMA_WARPS) {
    // =============================== producer warps ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HM_GPROD_REGS));
    const int pt = tid - 32 * HM_GMMA_WARPS;
    constexpr int SLOTS = HM_GMCH / 2, TPG = 32 * HM_GPROD_WARPS / SLOTS;     // 16-byte slots of a row segment; tracers per pass
    const int c = pt % SLOTS, pq = pt / SLOTS;          // my slot of a row segment; my tracer within a group of TPG
    int s = 0;
    uint32_t parity = 0;
    // (group, chunk) cursor advanced by increments: no 64-bit divisions per unit (they made the role spend half its time
    // on index arithmetic, and the MMA warps wait for its copies at the start of a kernel and after every group change)
    long long G = u0 / a.CH;
    int ch = (int)(u0 - G * a.CH);
    int g = (int)(G % a.ngroups), b = (int)(G / a.ngroups);
    const int nunits = (int)(u1 - u0);
    // my 16-byte slots of a stage: (wavenumber r, tracer pg * 8 + pq), slot c XOR-swizzled by the tracer's parity
    const uint32_t dst_thread = (uint32_t)((pq * SLOTS + (c ^ ((pq & 1) << 2))) * 16);
#ifdef HM_GRAM_PRESCALE
    // ---- second duty of the role: once a raw tile has landed, scale it IN PLACE (Y = s_p * grid, what the DMMAs contract) and
    // keep the diagonal work -- the two-halo sums sum_m r2 Y_p and the auto one-halo sums sum_m d_p grid^2 -- so that the MMA
    // warps issue nothing on the FP64 pipe but DMMAs (scaling and diagonal work on them cost 33 of 117 us).  A thread owns
    // SPT 16-byte slots of one tracer row for the four wavenumbers (slot order rotated by the tracer: no bank conflicts);
    // it reads and writes only its own slots, so the pass needs no synchronisation beyond the two barriers.
    constexpr int TPT = 32 * HM_GPROD_WARPS / PT, SPT = SLOTS / TPT;   // threads per tracer row; slots per thread
    const int sp = pt / TPT, part = pt - sp * TPT;
    const uint32_t swz_s = (uint32_t)(sp & 1) << 2;
    double a2h[HM_GRK], a1h[HM_GRK];
#pragma unroll
    for (int r = 0; r < HM_GRK; ++r) a2h[r] = a1h[r] = 0.;
    long long Gq = G;                                   // cursor of the unit being scaled (one behind the copies)
    int chq = ch;
    const double* __restrict__ gwq = a.gw + (size_t)b * a.gw_stride;
    int ginq = g, sq = 0;
    uint32_t parq = 0;
    auto scale_unit = [&](bool last) {
      constexpr int JB = SPT < 4 ? SPT : 4;             // slots per pass: the factors of a pass stay in registers (36 at JB = 4)
      unsigned char* st = gsm + (size_t)sq * STAGE_BYTES;
#pragma unroll
      for (int j0 = 0; j0 < SPT; j0 += JB) {
        double2 fS[JB], fD[JB], fR[JB];
        uint32_t off[JB];
#pragma unroll
        for (int j = 0; j < JB; ++j) {                  // (first pass: the factors fly while the tile lands)
          const int cs = part * SPT + ((j0 + j + sp) & (SPT - 1));
          const int mm = chq * HM_GMCH + 2 * cs;
          const int mc = (mm < a.nMp - 2) ? mm : a.nMp - 2;
          const double2* rec = reinterpret_cast<const double2*>(gwq + (2 * a.nMp + (mc >> 1) * HM_GREC(PT)));
          fS[j] = rec[sp]; fD[j] = rec[PT + sp]; fR[j] = rec[2 * PT];
          off[j] = (uint32_t)((sp * SLOTS + (cs ^ swz_s)) * 16);
        }
        if (j0 == 0) hm_mbar_wait(bar_raw + 8 * sq, parq);
#pragma unroll
        for (int r = 0; r < HM_GRK; ++r)
#pragma unroll
          for (int j = 0; j < JB; ++j) {
            double2* q = reinterpret_cast<double2*>(st + (size_t)r * PT * SLOTS * 16 + off[j]);
            const double2 X = *q;
            const double2 Y = make_double2(fS[j].x * X.x, fS[j].y * X.y);
            a2h[r] = fma(fR[j].x, Y.x, fma(fR[j].y, Y.y, a2h[r]));                             // sum_m w2 W_p  (:539)
            a1h[r] = fma(fD[j].x * X.x, X.x, fma(fD[j].y * X.y, X.y, a1h[r]));                 // sum_m d_p U^2  (:475-477)
            *q = Y;
          }
      }
      hm_mbar_arrive(bar_full + 8 * sq);                // (release: the MMA warps' wait sees my stores)
      if (++sq == HM_GNS) { sq = 0; parq ^= 1u; }
      if (chq == a.CH - 1 || last) {                    // the group ends here for this CTA: my rows' diagonal sums
        const int jslot = (int)blockIdx.x - hm_gram_cta_of(Gq * a.CH, a.total, a.ncta);
        double* dstp = a.partial + ((size_t)Gq * a.J + jslot) * (size_t)a.NW * HM_GRK;
#pragma unroll
        for (int r = 0; r < HM_GRK; ++r) {
          double v2 = a2h[r], v1 = a1h[r];
#pragma unroll
          for (int o = TPT / 2; o >= 1; o >>= 1) {
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          }
          if (part == 0 && sp < P) {
            dstp[(size_t)sp * HM_GRK + r] = v2;
            dstp[(size_t)(P + sp) * HM_GRK + r] = v1;
          }
          a2h[r] = a1h[r] = 0.;
        }
      }
      if (++chq == a.CH) {
        chq = 0; ++Gq;
        if (++ginq == a.ngroups) { ginq = 0; gwq += a.gw_stride; }
      }
    };
#endif
    for (int n = 0; n < nunits; ++n) {
      const int smp = b / a.F, row0 = g * HM_GRK;
      const int m = ch * HM_GMCH + 2 * c;
      const bool in = m < a.nM;                         // nM is even: the two masses of a slot are in or out together
      hm_mbar_wait(bar_empty + 8 * s, parity ^ 1u);     // the MMA warps are done with the stage
      const uint32_t dst0 = ring + (uint32_t)s * STAGE_BYTES + dst_thread;
#pragma unroll
      for (int r = 0; r < HM_GRK; ++r) {
        const int row = (row0 + r < a.nk) ? row0 + r : a.nk - 1;           // clamp: duplicates are never written
        const size_t off = ((size_t)smp * a.nk + row) * (size_t)a.nM + (in ? m : 0);
#pragma unroll
        for (int pg = 0; pg < PT / TPG; ++pg) {
          const int p = pg * TPG + pq;
          const bool on = in && p < P;
          const double* src = a.grid[on ? p : 0] + off;
          const uint32_t dst = dst0 + (uint32_t)((r * PT + pg * TPG) * SLOTS * 16);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(on ? 16 : 0) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_raw + 8 * s) : "memory");
      if (++s == HM_GNS) { s = 0; parity ^= 1u; }
      if (++ch == a.CH) {
        ch = 0;
        if (++g == a.ngroups) { g = 0; ++b; }
      }
#ifdef HM_GRAM_PRESCALE
      if (n >= 1) scale_unit(false);                    // the unit before this one: its copies were issued an iteration ago
#endif
    }
#ifdef HM_GRAM_PRESCALE
    if (nunits >= 1) scale_unit(true);
#endif
    return;
  }

  // ================================ MMA-group warps ================================
#if defined(HM_GRAM_SIXTEEN_WARPS)
  // (experiment) warps 2r, 2r + 1 (r = 0..3): the two halves of wavenumber r's triangle, DMMAs and fragment scaling only;
  // warp 8 + r: the diagonal work of wavenumber r.  Two DMMA-only warps per scheduler.
  if (warp < 2 * HM_GRK) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HM_G16_MMA_REGS));
    if (warp & 1) hm_gram_mma_role<NBT, 5>(a, gsm, bar_full, bar_empty, u0, u1, warp >> 1, lane);
    else hm_gram_mma_role<NBT, 4>(a, gsm, bar_full, bar_empty, u0, u1, warp >> 1, lane);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HM_G16_DIAG_REGS));
    hm_gram_mma_role<NBT, 3>(a, gsm, bar_full, bar_empty, u0, u1, warp - 2 * HM_GRK, lane);
  }
#elif defined(HM_GRAM_THREE_WAY)
  // (experiment) warps 3r, 3r + 1, 3r + 2 share wavenumber r: three DMMA-issuing warps per scheduler instead of two
  {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HM_G3_MMA_REGS));
    const int r = warp / 3, h = warp - 3 * r;
    if (h == 0) hm_gram_mma_role<NBT, 6>(a, gsm, bar_full, bar_empty, u0, u1, r, lane);
    else if (h == 1) hm_gram_mma_role<NBT, 7>(a, gsm, bar_full, bar_empty, u0, u1, r, lane);
    else hm_gram_mma_role<NBT, 8>(a, gsm, bar_full, bar_empty, u0, u1, r, lane);
  }
#elif !defined(HM_GRAM_WHOLE_TRIANGLE)
  // warps 2r and 2r + 1 share wavenumber r: half of the triangle and of the diagonal work each
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HM_GMMA_REGS));
  if (warp & 1) hm_gram_mma_role<NBT, 1>(a, gsm, bar_full, bar_empty, u0, u1, warp >> 1, lane);
  else hm_gram_mma_role<NBT, 0>(a, gsm, bar_full, bar_empty, u0, u1, warp >> 1, lane);
#else
  // (experiment, 127 against 122 us on config 4) warp r (0..3) runs the DMMAs of wavenumber r's whole triangle and nothing
  // else on the FP64 pipe but the scaling of its fragments; warp 4 + r, on the same scheduler, does the diagonal work of
  // that wavenumber from the same staged tile.  The MMA warp then issues DMMAs 63 % of its time instead of 35 %, but it is
  // alone on its scheduler's tensor pipe, and unrolling its block loop to overlap the next block's loads spills (248
  // registers: 144 of them accumulators).
  if (warp < HM_GRK) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HM_GMMA_ALL_REGS));
    hm_gram_mma_role<NBT, 2>(a, gsm, bar_full, bar_empty, u0, u1, warp, lane);
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HM_GDIAG_REGS));
    hm_gram_mma_role<NBT, 3>(a, gsm, bar_full, bar_empty, u0, u1, warp - HM_GRK, lane);
  }
#endif
}

// ---------------------------------------------------------------------------------------------------------------
struct hm_gram_eargs {
  int nk, nM, nMp, P, S, NW, F, B, ngroups, CH, ncta, J;
  long long total;
  int simple_twohalo, subtract_shotnoise, correct_discrete, fence_out;
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

// CTA = 32 consecutive unique pairs x 8 consecutive groups of HM_GRK wavenumbers of one model; thread (pair, group).
// A thread finishes its pair for the four wavenumbers of its group with 32-byte reads of the partial results
// ([slot][sum][4 wavenumbers]): consecutive lanes read consecutive cross sums, the same two-halo sum of tracer u and
// consecutive ones of tracer v.  The 32 pairs x 3 terms x 32 wavenumbers of the CTA go through shared memory so that
// every store instruction writes 256 contiguous bytes of one output row.
#define HM_GE_PAIRS 32
#define HM_GE_GROUPS 8
__global__ void __launch_bounds__(HM_GE_PAIRS * HM_GE_GROUPS) hm_gram_epilogue_kernel(const hm_gram_eargs a) {
  static_assert(HM_GRK == 4, "a thread handles the four wavenumbers of a group as two double2");
  constexpr int ROWS = HM_GE_GROUPS * HM_GRK;                         // wavenumbers per CTA
  __shared__ double so[HM_GE_PAIRS * 3][ROWS + 1];
  const int P = a.P;
  const int sl = threadIdx.x % HM_GE_PAIRS, gl = threadIdx.x / HM_GE_PAIRS;
  const int gtiles = (a.ngroups + HM_GE_GROUPS - 1) / HM_GE_GROUPS;
  const int b = blockIdx.y / gtiles, g0 = (blockIdx.y - b * gtiles) * HM_GE_GROUPS;
  const int s0 = blockIdx.x * HM_GE_PAIRS, s = s0 + sl, g = g0 + gl;
  const int smp = b / a.F;
  if (s < a.S && g < a.ngroups) {
    const long long G = (long long)b * a.ngroups + g;
    const int row0 = g * HM_GRK;
    // pair s -> (u, v), u <= v, row-major: rows start at off(u) = u P - u(u-1)/2; invert with a float guess + fix-up
    int u = (int)(((2. * P + 1.) - sqrt((2. * P + 1.) * (2. * P + 1.) - 8. * (double)s)) * 0.5);
    u = u < 0 ? 0 : (u > P - 1 ? P - 1 : u);
    while (u > 0 && u * P - u * (u - 1) / 2 > s) --u;
    while ((u + 1) * P - (u + 1) * u / 2 <= s) ++u;
    const int v = u + (s - (u * P - u * (u - 1) / 2));
    const double A = a.scal[(size_t)b * HM_GNSCAL], rhom = a.scal[(size_t)b * HM_GNSCAL + 1];
    const size_t i1h = (u == v) ? (size_t)(P + u)
                                : (size_t)2 * P + (size_t)u * P - (size_t)u * (u + 1) / 2 + (v - u - 1);
    double fu[HM_GRK], fv[HM_GRK], f1[HM_GRK];
#pragma unroll
    for (int r = 0; r < HM_GRK; ++r) fu[r] = fv[r] = f1[r] = 0.;
    // the partial results of the group, one per CTA that worked on it, added in CTA order
    const int c_lo = hm_gram_cta_of(G * a.CH, a.total, a.ncta), c_hi = hm_gram_cta_of((G + 1) * a.CH - 1, a.total, a.ncta);
    for (int j = 0; j <= c_hi - c_lo; ++j) {
      const double2* pp = reinterpret_cast<const double2*>(a.partial + ((size_t)G * a.J + j) * (size_t)a.NW * HM_GRK);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double2 xu = pp[(size_t)u * 2 + h], xv = pp[(size_t)v * 2 + h], x1 = pp[i1h * 2 + h];
        fu[2 * h] += xu.x; fu[2 * h + 1] += xu.y;
        fv[2 * h] += xv.x; fv[2 * h + 1] += xv.y;
        f1[2 * h] += x1.x; f1[2 * h + 1] += x1.y;
      }
    }
    const double Aterm = A / a.M[0];
    const double cu0 = a.tscal[((size_t)b * P + u) * HM_GNTS + 2], cv0 = a.tscal[((size_t)b * P + v) * HM_GNTS + 2];
    const double Iu_simple = a.tscal[((size_t)b * P + u) * HM_GNTS + 1], Iv_simple = a.tscal[((size_t)b * P + v) * HM_GNTS + 1];
    const double sn = a.tscal[((size_t)b * P + u) * HM_GNTS];
#pragma unroll
    for (int r = 0; r < HM_GRK; ++r) {
      const int row = (row0 + r < a.nk) ? row0 + r : a.nk - 1;
      const size_t goff = ((size_t)smp * a.nk + row) * (size_t)a.nM;     // W_p(k, M[0]) = c_p[0] * grid_p[k, 0]  (:541-542)
      double xu = fu[r], xv = fv[r];
      if (a.mass[u]) xu += Aterm * (cu0 * a.grid[u][goff]);
      if (a.mass[v]) xv += Aterm * (cv0 * a.grid[v][goff]);
      const double Iu = a.simple_twohalo ? Iu_simple : xu * rhom;
      const double Iv = a.simple_twohalo ? Iv_simple : xv * rhom;
      const double p2h = __dmul_rn(__dadd_rn(__dmul_rn(Iu, Iv), 0.), a.Pk_lin[(size_t)smp * a.nk + row]);   // :524
      double p1h = f1[r] * rhom;                                                                           // :532
      if (u == v && a.discrete[u]) {                                                                       // :484-491
        if (a.correct_discrete && !a.subtract_shotnoise) p1h += sn;
        else if (!a.correct_discrete && a.subtract_shotnoise) p1h -= sn;
      }
      if (a.k_trunc > 0.) {                                                                                // :494-495
        const double q = a.k[row] / a.k_trunc;
        p1h *= 1. - exp(-(q * q));
      }
      so[sl * 3 + 0][gl * HM_GRK + r] = p2h;
      so[sl * 3 + 1][gl * HM_GRK + r] = p1h;
      so[sl * 3 + 2][gl * HM_GRK + r] = __dadd_rn(p2h, p1h);                                               // :500-501
    }
  }
  __syncthreads();
  // coalesced stores: one (pair, term) output row segment of ROWS wavenumbers per warp instruction
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = g0 * HM_GRK + lane;
    for (int q = warp; q < HM_GE_PAIRS * 3; q += HM_GE_GROUPS) {
      const int sp = s0 + q / 3;
      if (sp < a.S && row < a.nk) a.out[(((size_t)b * a.S + sp) * 3 + q % 3) * (size_t)a.nk + row] = so[q][lane];
    }
  }
  if (a.fence_out) {           // `out` is a peer GPU's buffer (fused gather): one cumulative system-scope fence
    __syncthreads();           // per CTA, see hm_epilogue_kernel
    if (threadIdx.x == 0) __threadfence_system();
  }
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
  return hm_gram_power_spectrum_stages(pr, out_dev, lowmass_dev, ws, ws_bytes, stream, 7);
}

// stages: bit 0 = the three weight kernels, bit 1 = the Gram (DMMA) kernel, bit 2 = the epilogue
extern "C" int hm_gram_power_spectrum_stages(const hm_gram_problem* pr, double* out_dev, double* lowmass_dev, void* ws,
                                             size_t ws_bytes, void* stream, int stages) {
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
  w.nM = pr->nM; w.nMp = L.nMp; w.P = L.P; w.PTK = L.PTK; w.F = pr->models_per_sample; w.correct_discrete = pr->correct_discrete;
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
  w.need_integrals = pr->simple_twohalo || ((pr->correct_discrete != 0) != (pr->subtract_shotnoise != 0));
  w.gw = wsd; w.scal = wsd + L.scal_offset; w.tscal = wsd + L.tscal_offset; w.lowmass = lowmass_dev;
  if (stages & 1) {
    hm_gram_weights_kernel<<<dim3(L.nchunks, 1 + L.PTK, (unsigned)L.B), 256, 0, st>>>(w);
    HM_LAUNCH_CHECK();
    if (w.need_integrals) {      // shot noise / simple two-halo integrals: only non-default flags read them
      hm_gram_scalars_kernel<<<dim3(L.P, (unsigned)L.B), 256, 0, st>>>(w);
      HM_LAUNCH_CHECK();
    }
  }

  hm_gram_args g;
  g.nk = pr->nk; g.nM = pr->nM; g.nMp = L.nMp; g.P = L.P; g.NW = L.NW; g.F = pr->models_per_sample;
  g.ngroups = L.ngroups; g.CH = L.CH; g.ncta = L.ncta; g.J = L.J; g.total = L.total; g.gw_stride = L.gw_stride;
  for (int p = 0; p < HM_GMAXP; ++p) g.grid[p] = (p < L.P) ? pr->tracers[p].grid_dev : nullptr;
  g.gw = wsd; g.partial = wsd + L.partial_offset;
  // the kernel is instantiated for block grids of 2, 4 and 8 (16, 32, 64 tracer rows); pad up to the next one
  const int NBT = L.PTK / 8;
  const size_t smem = (size_t)HM_GNS * HM_GRK * (8 * NBT) * HM_GMCH * sizeof(double) + 128;
  static hm_once_per_device once;
  if (once.needed()) {
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gram_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gram_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gram_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  }
  if (stages & 2) {
    if (NBT == 2) hm_gram_kernel<2><<<(unsigned)L.ncta, HM_GTHREADS, smem, st>>>(g);
    else if (NBT == 4) hm_gram_kernel<4><<<(unsigned)L.ncta, HM_GTHREADS, smem, st>>>(g);
    else hm_gram_kernel<8><<<(unsigned)L.ncta, HM_GTHREADS, smem, st>>>(g);
    HM_LAUNCH_CHECK();
  }
  if (!(stages & 4)) return HM_OK;

  hm_gram_eargs e;
  e.nk = pr->nk; e.nM = pr->nM; e.nMp = L.nMp; e.P = L.P; e.S = L.S; e.NW = L.NW; e.F = pr->models_per_sample;
  e.B = (int)L.B; e.ngroups = L.ngroups; e.CH = L.CH; e.ncta = L.ncta; e.J = L.J; e.total = L.total;
  e.simple_twohalo = pr->simple_twohalo; e.subtract_shotnoise = pr->subtract_shotnoise;
  e.correct_discrete = pr->correct_discrete; e.k_trunc = pr->k_trunc; e.gw_stride = L.gw_stride;
  e.fence_out = pr->output_is_peer;
  e.k = pr->k_dev; e.M = pr->M_dev; e.Pk_lin = pr->Pk_lin_dev; e.partial = g.partial; e.gw = wsd;
  e.scal = w.scal; e.tscal = w.tscal;
  for (int p = 0; p < HM_GMAXP; ++p) {
    e.grid[p] = g.grid[p];
    e.mass[p] = w.mass[p];
    e.discrete[p] = w.discrete[p];
  }
  e.out = out_dev;
  const int gtiles = (L.ngroups + HM_GE_GROUPS - 1) / HM_GE_GROUPS;
  hm_gram_epilogue_kernel<<<dim3((L.S + HM_GE_PAIRS - 1) / HM_GE_PAIRS, (unsigned)(L.B * gtiles)), HM_GE_PAIRS * HM_GE_GROUPS, 0, st>>>(e);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
