// hm_gram.cu -- many-tracer halo-model spectra: the one-halo cross terms as a per-wavenumber Gram contraction on
// the FP64 tensor cores (BASELINE config 4: 64 HOD tracers, 2080 unique pairs per k).
//
// For P tracers the cross one-halo integrals of model.power_spectrum (pyhalomodel.py:479, :526-533) are
//     P1h_uv(k) = rhom * sum_m w1[m] * W_u(k,m) * W_v(k,m),      W_p(k,m) = c_p[m] * grid_p[k,m]
// i.e. per k the upper triangle of (w1 o W) W^T, a P x P x nM product.  With P(P+1)/2 pairs the arithmetic
// intensity (~6 flop/B at P=64) is beyond what the streaming-quadrature kernel's per-pair register accumulators
// can hold, so this path tiles it for mma.sync.m8n8k4.f64 (SASS DMMA; there is no tcgen05 FP64 MMA):
//
//   hm_gram_weights_kernel  w1, w2 and r2 = w2/sqrt(w1) per mass; A
//   hm_gram_tracer_kernel   s_p = sqrt(w1) c_p (the staged tile is Y = s_p*grid, so that the contraction needs no weight:
//                           sum_m w1 W_u W_v = sum_m Y_u Y_v with W = c_p*grid) and d_p (auto one-halo weight)
//   hm_gram_scalars_kernel  shot noise and simple-two-halo integrals per (tracer, model)
//   hm_gram_kernel<RK,NBT>  persistent warp-specialised CTA per SM over (model, RK wavenumbers, mass part) items:
//                           prep warps stream the tracer rows of the next 64-mass chunk with cp.async into one of
//                           three W-tile buffers, scale them in place to W and keep the diagonal work (two-halo
//                           sums sum_m w2 W_p, auto one-halo sums sum_m d_p grid_p^2) in registers; MMA warps run
//                           DMMA over statically owned 8x8 upper-triangle fragments; named barriers hand tiles over.
//   hm_gram_epilogue_kernel adds the mass parts and finishes P_2h, P_1h, P_hm for every unique pair (:456-501)
//
// The diagonal (auto) terms are NOT Gram entries: they use Uk, amplitude, variance (pyhalomodel.py:465-477).
#include "hm_common.cuh"

#define HM_GMAXP HM_GRAM_MAX_TRACERS
#define HM_GMCH 32                 // masses per chunk
#define HM_GLD (HM_GMCH + 8)       // padded row stride of the tile: 320 B = 64 mod 128, so the 16-byte DMMA
                                   // fragment loads of a quarter warp (2 rows x 4 lanes) hit all 32 banks once
#define HM_GNBUF 5                 // tile buffers (40 KB each at 64 tracer rows x 2 wavenumbers) ...
#define HM_GDEPTH 3                // ... of which this many are in flight from HBM ahead of the one being scaled
#define HM_GNSCAL 4
#define HM_GNTS 3                  // per (model, tracer) scalars: shot noise, simple two-halo integral, c_p[M0]
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
  const int maxparts = (pr->nM + 255) / 256;                             // at least 256 masses per part
  while (nparts < maxparts && rows * nparts < 6LL * 148) nparts *= 2;
  if (nparts > maxparts) nparts = maxparts;
  L.MP = (((pr->nM + nparts - 1) / nparts) + HM_GMCH - 1) / HM_GMCH * HM_GMCH;
  L.nparts = (pr->nM + L.MP - 1) / L.MP;
  auto align = [](size_t x) { return (x + 31) & ~(size_t)31; };
  L.gw_stride = (size_t)(3 + 2 * L.P) * L.nMp;          // w1, w2, s_p [P], d_p [P], r2
  L.scal_offset = align(L.B * L.gw_stride);
  L.tscal_offset = align(L.scal_offset + L.B * HM_GNSCAL);
  L.partial_offset = align(L.tscal_offset + L.B * (size_t)L.P * HM_GNTS);
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
    gw[m] = 0.;
    gw[(size_t)a.nMp + m] = 0.;
    gw[(size_t)(2 + 2 * a.P) * a.nMp + m] = 0.;
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
  gw[(size_t)(2 + 2 * a.P) * a.nMp + m] = (w1 > 0.) ? w2 / sqrt(w1) : 0.;   // sum_m w2 W_p = sum_m r2 Y_p  (:539)
}

// per-(tracer, mass) factors: s_p = sqrt(w1) c_p (W = c_p * grid, Y = s_p * grid) and d_p (auto one-halo weight);
// grid = (mass chunk, tracer, model)
__global__ void __launch_bounds__(256) hm_gram_tracer_kernel(const hm_gram_wargs a) {
  const int p = blockIdx.y, b = blockIdx.z, s = b / a.F;
  const int m = blockIdx.x * 256 + threadIdx.x;
  if (m >= a.nMp) return;
  double* gw = a.gw + (size_t)b * a.gw_stride;
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
    dw = __dmul_rn(__dmul_rn(gw[m], fac), uS * uS);                                 // d_p = w1 fac uS^2 (:475-477)
  }
  gw[(size_t)(2 + p) * a.nMp + m] = (m < a.nM) ? sqrt(gw[m]) * c : 0.;
  gw[(size_t)(2 + a.P + p) * a.nMp + m] = dw;
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
struct hm_gram_args {
  int nk, nM, nMp, P, NB, PT, NW, F, ntiles, nparts, MP;
  long long nitems;
  size_t gw_stride;
  const double* grid[HM_GMAXP];
  const double* gw;
  double* partial;          // [B][nparts][NW][nk]
};

__device__ __forceinline__ void hm_dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}


struct hm_gram_cursor {      // (item, chunk) iterator over the contiguous item range of a persistent CTA
  int item, ch, nch_item;
  int b, smp, row0, mbeg, mend;
};

template <int RK>
__device__ __forceinline__ void hm_gram_decode(hm_gram_cursor& c, const hm_gram_args& a) {
  int it = c.item;
  const int part = it % a.nparts; it /= a.nparts;
  const int tile = it % a.ntiles;
  c.b = it / a.ntiles;
  c.smp = c.b / a.F;
  c.row0 = tile * RK;
  c.mbeg = part * a.MP;
  c.mend = (c.mbeg + a.MP < a.nM) ? c.mbeg + a.MP : a.nM;
  c.nch_item = (c.mend - c.mbeg + HM_GMCH - 1) / HM_GMCH;
}

// ---- DMMA over the fragments one warp owns -------------------------------------------------------------------
// Fragment ownership (8x8 blocks of the upper triangle of the NBT x NBT block grid): warp pair WQ = warp/2 owns
// block rows I1 = WQ and I2 = NBT-1-WQ (together NBT+1 fragments) and, within the pair, warp E = warp%2 takes
// the block columns j = 2jj+E.  Everything is a template parameter, so DMMAs are unconditional, accumulators
// are statically indexed registers and the balance is 4-5 fragments per warp at NBT = 8.
template <int NBT, int WQ, int E>
struct hm_gram_frag {
  static constexpr int I1 = WQ, I2 = NBT - 1 - WQ;
  static constexpr bool ROW1 = I1 <= I2, ROW2 = I1 < I2;
  __host__ __device__ static constexpr bool v1(int jj) { return ROW1 && (2 * jj + E) >= I1 && (2 * jj + E) < NBT; }
  __host__ __device__ static constexpr bool v2(int jj) { return ROW2 && (2 * jj + E) >= I2 && (2 * jj + E) < NBT; }
};

template <int RK, int NBT, int WQ, int E, int JJ, int HALF>
__device__ __forceinline__ void hm_gram_mma_jj(double (&acc)[RK][2][4][2], int r, const double2& a1, const double2& a2,
                                               const double2& bb) {
  using F = hm_gram_frag<NBT, WQ, E>;
  if constexpr (F::v1(JJ)) hm_dmma884(acc[r][0][JJ][0], acc[r][0][JJ][1], HALF ? a1.y : a1.x, HALF ? bb.y : bb.x);
  if constexpr (F::v2(JJ)) hm_dmma884(acc[r][1][JJ][0], acc[r][1][JJ][1], HALF ? a2.y : a2.x, HALF ? bb.y : bb.x);
}

// C[u,v] += sum_m Y_u Y_v over one staged 64-mass tile (Y = sqrt(w1) W: the weight is folded into both operands).  Two k4-steps per 16-byte fragment load: within an
// 8-mass block step 0 takes masses {0,2,4,6} and step 1 {1,3,5,7} (any assignment works as long as A, B and w1
// agree: the sum over masses is order-free).
template <int RK, int NBT, int WQ, int E>
__device__ __forceinline__ void hm_gram_mma(double (&acc)[RK][2][4][2], const double* __restrict__ Yt, int fr, int fk) {
  using F = hm_gram_frag<NBT, WQ, E>;
  constexpr int PT = 8 * NBT;
  if constexpr (!F::ROW1) return;
  const double* Yb = Yt + 2 * fk + (size_t)fr * HM_GLD;
#pragma unroll 2
  for (int blk = 0; blk < HM_GMCH / 8; ++blk) {
#pragma unroll
    for (int r = 0; r < RK; ++r) {
      const double* Yr = Yb + (size_t)r * PT * HM_GLD + 8 * blk;
      const double2 a1 = *reinterpret_cast<const double2*>(Yr + (size_t)(8 * F::I1) * HM_GLD);
      double2 a2 = make_double2(0., 0.);
      if constexpr (F::ROW2) a2 = *reinterpret_cast<const double2*>(Yr + (size_t)(8 * F::I2) * HM_GLD);
      double2 b0 = make_double2(0., 0.), b1 = b0, b2 = b0, b3 = b0;
      if constexpr (F::v1(0) || F::v2(0)) b0 = *reinterpret_cast<const double2*>(Yr + (size_t)(8 * (0 + E)) * HM_GLD);
      if constexpr (F::v1(1) || F::v2(1)) b1 = *reinterpret_cast<const double2*>(Yr + (size_t)(8 * (2 + E)) * HM_GLD);
      if constexpr (F::v1(2) || F::v2(2)) b2 = *reinterpret_cast<const double2*>(Yr + (size_t)(8 * (4 + E)) * HM_GLD);
      if constexpr (F::v1(3) || F::v2(3)) b3 = *reinterpret_cast<const double2*>(Yr + (size_t)(8 * (6 + E)) * HM_GLD);
      hm_gram_mma_jj<RK, NBT, WQ, E, 0, 0>(acc, r, a1, a2, b0);
      hm_gram_mma_jj<RK, NBT, WQ, E, 1, 0>(acc, r, a1, a2, b1);
      hm_gram_mma_jj<RK, NBT, WQ, E, 2, 0>(acc, r, a1, a2, b2);
      hm_gram_mma_jj<RK, NBT, WQ, E, 3, 0>(acc, r, a1, a2, b3);
      hm_gram_mma_jj<RK, NBT, WQ, E, 0, 1>(acc, r, a1, a2, b0);
      hm_gram_mma_jj<RK, NBT, WQ, E, 1, 1>(acc, r, a1, a2, b1);
      hm_gram_mma_jj<RK, NBT, WQ, E, 2, 1>(acc, r, a1, a2, b2);
      hm_gram_mma_jj<RK, NBT, WQ, E, 3, 1>(acc, r, a1, a2, b3);
    }
  }
}

// Write one warp's fragments of a finished item and reset them.
// C fragment layout (m8n8k4.f64): c0/c1 = C[8i + lane/4][8j + 2*(lane%4) + {0,1}]
template <int RK, int NBT, int WQ, int E>
__device__ __forceinline__ void hm_gram_flush(double (&acc)[RK][2][4][2], double* __restrict__ dst, int P, int nk,
                                              int row0, int fr, int fk) {
  using F = hm_gram_frag<NBT, WQ, E>;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const bool valid = (h == 0) ? F::v1(jj) : F::v2(jj);
      if (valid) {
        const int u = 8 * (h == 0 ? F::I1 : F::I2) + fr;
#pragma unroll
        for (int ee = 0; ee < 2; ++ee) {
          const int v = 8 * (2 * jj + E) + 2 * fk + ee;
          if (u < v && v < P) {                        // strict upper triangle, row-major index of (u, v)
            const size_t idx = (size_t)2 * P + (size_t)u * P - (size_t)u * (u + 1) / 2 + (v - u - 1);
#pragma unroll
            for (int r = 0; r < RK; ++r)
              if (row0 + r < nk) dst[idx * nk + row0 + r] = acc[r][h][jj][ee];
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RK; ++r) acc[r][h][jj][0] = acc[r][h][jj][1] = 0.;
    }
}

#define HM_GRAM_WARP_SWITCH(FN, ...)                           \
  switch (warp) {                                              \
    case 0: FN<RK, NBT, 0, 0>(__VA_ARGS__); break;             \
    case 1: FN<RK, NBT, 0, 1>(__VA_ARGS__); break;             \
    case 2: FN<RK, NBT, 1, 0>(__VA_ARGS__); break;             \
    case 3: FN<RK, NBT, 1, 1>(__VA_ARGS__); break;             \
    case 4: FN<RK, NBT, 2, 0>(__VA_ARGS__); break;             \
    case 5: FN<RK, NBT, 2, 1>(__VA_ARGS__); break;             \
    case 6: FN<RK, NBT, 3, 0>(__VA_ARGS__); break;             \
    default: FN<RK, NBT, 3, 1>(__VA_ARGS__); break;            \
  }

// ---- named barriers between the two warp roles (ids 1..10; 0 is __syncthreads) --------------------------------
#define HM_GBAR_FULL 1                   // + buf: prep warps arrive, MMA warps wait
#define HM_GBAR_EMPTY (1 + HM_GNBUF)     // + buf: MMA warps arrive, prep warps wait

// v[0..16) per lane -> the lane whose low four bits are j holds the sum of v[j] over the 16 lanes of its half warp
__device__ __forceinline__ double hm_half_transpose_sum16(double (&v)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
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
#define HM_GTHREADS (64 * HM_GWARPS)
__device__ __forceinline__ void hm_bar_sync(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(HM_GTHREADS) : "memory");
}
__device__ __forceinline__ void hm_bar_arrive(int id) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(HM_GTHREADS) : "memory");
}

__device__ __forceinline__ bool hm_gram_advance(hm_gram_cursor& c, long long it1) {   // false when the range is done
  if (++c.ch < c.nch_item) return true;
  c.ch = 0;
  return ++c.item < (int)it1;
}

// Persistent, warp-specialised CTA (one per SM, 16 warps):
//   prep warps 8..15  stream the tracer rows of 32-mass chunks from HBM with cp.async (LDGSTS, no registers held by
//                     loads in flight) into a ring of HM_GNBUF tile buffers, HM_GDEPTH chunks (120 KB per SM) ahead of
//                     the chunk being worked on -- with a single chunk in flight the role ran at 2.8 TB/s and bounded
//                     the kernel (measured with the DMMAs compiled out: 102 of 147 us); scale the landed tile in place
//                     to Y = s_p * grid (s_p = sqrt(w1) c_p, so the contraction needs no weight) and accumulate the
//                     diagonal work -- two-halo sums sum_m w2 W_p and auto one-halo sums sum_m d_p grid_p^2 -- in
//                     registers, reduced once per item;
//   MMA warps 0..7    run DMMA over their fragments of the staged tile.
// The roles hand tiles over with named barriers (bar.arrive / bar.sync), so loads, staging and tensor work of
// different chunks overlap and there is no CTA-wide synchronisation.
template <int RK, int NBT>
__global__ void __launch_bounds__(HM_GTHREADS, 1) hm_gram_kernel(const hm_gram_args a) {
  static_assert(RK == 2 && HM_GMCH == 32, "a prep lane owns one mass pair of the chunk for both wavenumber rows");
  static_assert(2 * NBT <= 16, "the diagonal sums of a prep lane fit one 16-value reduction");
  extern __shared__ __align__(16) double gsm[];
  constexpr int PT = 8 * NBT;
  constexpr size_t tile_doubles = (size_t)RK * PT * HM_GLD;
  double* Ys = gsm;                                     // [HM_GNBUF][RK][PT][HM_GLD]  tiles (raw grid rows until scaled)
  const int P = a.P;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform: role dispatch does not diverge

  const long long per = (a.nitems + gridDim.x - 1) / gridDim.x;
  const long long it0 = (long long)blockIdx.x * per;
  const long long it1 = (it0 + per < a.nitems) ? it0 + per : a.nitems;
  if (it0 >= it1) return;

  hm_gram_cursor cur;
  cur.item = (int)it0; cur.ch = 0;
  hm_gram_decode<RK>(cur, a);

  if (warp >= HM_GWARPS) {
    // =============================== prep warps ===============================
    // Warp pw owns tracers pw, pw + 8, ...; inside it a lane owns one mass pair of the 32-mass chunk (mp) for BOTH
    // wavenumber rows and every second one of the warp's tracers (th), so the per-(tracer, mass) factors s_p, d_p are
    // loaded once for the two rows.  They come from L2 (~1000 cycles under load): the factors of chunk n + 1 are
    // loaded while chunk n is scaled, a full iteration before they are used.
    static_assert(NBT % 2 == 0, "the two lane halves split the warp's tracers");
    constexpr int TL = NBT / 2;                         // tracers per lane
    const int pw = warp - HM_GWARPS;
    const int mp = lane & 15, th = lane >> 4;
    const double2 zero = make_double2(0., 0.);
    double vals[16];                                    // diagonal sums of the current item: [tt][row][2h, auto 1h]
#pragma unroll
    for (int i = 0; i < 16; ++i) vals[i] = 0.;

    auto issue = [&](const hm_gram_cursor& c, int buf) {
      const int m = c.mbeg + c.ch * HM_GMCH + 2 * mp;
      const bool in = m < c.mend;
      const int mc = in ? m : 0;                        // keep the (unread) source address inside the row
      const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(Ys + (size_t)buf * tile_doubles + 2 * mp);
#pragma unroll
      for (int r = 0; r < RK; ++r) {
        const int row = (c.row0 + r < a.nk) ? c.row0 + r : a.nk - 1;         // clamp: duplicates are never written
        const size_t off = ((size_t)c.smp * a.nk + row) * (size_t)a.nM + mc;
#pragma unroll
        for (int tt = 0; tt < TL; ++tt) {
          const int p = pw + 8 * (2 * tt + th);
          const bool on = in && p < P;
          const double* src = a.grid[on ? p : 0] + off;
          const uint32_t dst = dst0 + (uint32_t)(((size_t)r * PT + p) * HM_GLD * sizeof(double));
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(on ? 16 : 0) : "memory");
        }
      }
    };
    double2 cc[TL], dd[TL], w2v;                        // s_p, d_p of my tracers and r2 at my mass pair, for chunk `cur`
    auto load_factors = [&](const hm_gram_cursor& c) {
      const int m = c.mbeg + c.ch * HM_GMCH + 2 * mp;
      const bool in = m < c.mend;                       // nM and MP are even: a lane's two masses are in or out together
      const double* __restrict__ gw = a.gw + (size_t)c.b * a.gw_stride;
#pragma unroll
      for (int tt = 0; tt < TL; ++tt) {
        const int p = pw + 8 * (2 * tt + th);
        const bool on = in && p < P;
        cc[tt] = on ? *reinterpret_cast<const double2*>(gw + (size_t)(2 + p) * a.nMp + m) : zero;
        dd[tt] = on ? *reinterpret_cast<const double2*>(gw + (size_t)(2 + P + p) * a.nMp + m) : zero;
      }
      w2v = in ? *reinterpret_cast<const double2*>(gw + (size_t)(2 + 2 * P) * a.nMp + m) : zero;    // r2
    };

    hm_gram_cursor fill = cur;                          // `fill` runs HM_GDEPTH chunks ahead of `cur`
    bool more = true;
#pragma unroll
    for (int d = 0; d < HM_GDEPTH; ++d) {               // one cp.async group per chunk, empty when there is none
      if (more) {
        issue(fill, d);
        more = hm_gram_advance(fill, it1);
        if (more && fill.ch == 0) hm_gram_decode<RK>(fill, a);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    load_factors(cur);
    int n = 0;
    while (true) {
      const int buf = n % HM_GNBUF;
      if (more) {                                       // chunk n + DEPTH goes into the buffer chunk n + DEPTH - NBUF used,
        const int nb = (n + HM_GDEPTH) % HM_GNBUF;      // once the MMA warps are done with that one
        if (n + HM_GDEPTH >= HM_GNBUF) hm_bar_sync(HM_GBAR_EMPTY + nb);
        issue(fill, nb);
        more = hm_gram_advance(fill, it1);
        if (more && fill.ch == 0) hm_gram_decode<RK>(fill, a);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(HM_GDEPTH) : "memory");      // my copies of chunk n have landed
      double* Y = Ys + (size_t)buf * tile_doubles + 2 * mp;
      double2 xv[TL][RK];
#ifdef HM_GRAM_DEBUG_NO_SCALE
#pragma unroll
      for (int tt = 0; tt < TL; ++tt)
#pragma unroll
        for (int r = 0; r < RK; ++r) xv[tt][r] = zero;
      if (false)
#endif
#pragma unroll
      for (int tt = 0; tt < TL; ++tt)
#pragma unroll
        for (int r = 0; r < RK; ++r) {
          const int p = pw + 8 * (2 * tt + th);
          double2* cell = reinterpret_cast<double2*>(Y + ((size_t)r * PT + p) * HM_GLD);
          xv[tt][r] = *cell;                            // raw grid values (zero-filled where off)
          const double2 y = make_double2(cc[tt].x * xv[tt][r].x, cc[tt].y * xv[tt][r].y);
          *cell = y;                                    // Y = s_p * grid, in place
#ifndef HM_GRAM_DEBUG_NO_DIAG
          double& v2h = vals[(tt * RK + r) * 2];
          v2h = fma(w2v.x, y.x, fma(w2v.y, y.y, v2h));                                             // sum_m r2 Y_p = sum_m w2 W_p  (:539)
#endif
        }
      // No __threadfence_block() here: bar.arrive orders this thread's shared-memory stores before the barrier completes
      // for the threads that bar.sync on it, and a CTA-scope fence would also wait for the cp.async copies of the NEXT
      // chunks still in flight -- it turned the ring into one DRAM round trip per chunk (prep role alone: 100 us).
      hm_bar_arrive(HM_GBAR_FULL + buf);                // the tile is staged: the rest of the iteration is off the MMA warps' path
#ifndef HM_GRAM_DEBUG_NO_DIAG
#pragma unroll
      for (int tt = 0; tt < TL; ++tt)
#pragma unroll
        for (int r = 0; r < RK; ++r) {
          double& v1h = vals[(tt * RK + r) * 2 + 1];
          v1h = fma(dd[tt].x * xv[tt][r].x, xv[tt][r].x, fma(dd[tt].y * xv[tt][r].y, xv[tt][r].y, v1h));   // sum_m d_p U^2  (:475-477)
        }
#endif
      const bool item_done = cur.ch + 1 == cur.nch_item;
      const hm_gram_cursor done = cur;
      ++n;
      const bool go_on = hm_gram_advance(cur, it1);
      if (go_on) {
        if (cur.item != done.item) hm_gram_decode<RK>(cur, a);
        load_factors(cur);                              // for the next iteration: a whole iteration of latency hiding
      }
      if (item_done) {                                  // item finished: reduce and write my tracers' diagonal sums
        const int part = done.item % a.nparts;
        double* dst = a.partial + ((size_t)done.b * a.nparts + part) * (size_t)a.NW * a.nk;
        const double tot = hm_half_transpose_sum16(vals, lane);          // lane (th, j) now holds value j of its half
        const int tt = mp >> 2, r = (mp >> 1) & 1, kind = mp & 1, p = pw + 8 * (2 * tt + th);
        if (tt < TL && p < P && done.row0 + r < a.nk) dst[(size_t)(kind * P + p) * a.nk + done.row0 + r] = tot;
#pragma unroll
        for (int i = 0; i < 16; ++i) vals[i] = 0.;
      }
      if (!go_on) break;
    }
    return;
  }

  // ================================ MMA warps ================================
  double acc[RK][2][4][2];
#pragma unroll
  for (int r = 0; r < RK; ++r)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) acc[r][h][jj][0] = acc[r][h][jj][1] = 0.;
  const int fr = lane >> 2, fk = lane & 3;
  int n = 0;
  while (true) {
    const int buf = n % HM_GNBUF;
    hm_bar_sync(HM_GBAR_FULL + buf);                    // the W tile of this chunk is staged
    {
#ifndef HM_GRAM_DEBUG_NO_MMA
      const double* Yt = Ys + (size_t)buf * tile_doubles;
      HM_GRAM_WARP_SWITCH(hm_gram_mma, acc, Yt, fr, fk)
#endif
    }
    hm_bar_arrive(HM_GBAR_EMPTY + buf);                 // all my reads of the tile have been consumed by DMMAs
    if (cur.ch + 1 == cur.nch_item) {                   // item finished: write and reset my fragments
      const int part = cur.item % a.nparts;
      double* dst = a.partial + ((size_t)cur.b * a.nparts + part) * (size_t)a.NW * a.nk;
      HM_GRAM_WARP_SWITCH(hm_gram_flush, acc, dst, P, a.nk, cur.row0, fr, fk)
    }
    ++n;
    const int prev_item = cur.item;
    if (!hm_gram_advance(cur, it1)) break;
    if (cur.item != prev_item) hm_gram_decode<RK>(cur, a);
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct hm_gram_eargs {
  int nk, nM, nMp, P, S, NW, F, nparts, B;
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

// thread per (model, unique pair, wavenumber), wavenumber fastest => coalesced reads of the partial sums and writes
__global__ void __launch_bounds__(256) hm_gram_epilogue_kernel(const hm_gram_eargs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_model = (long long)a.S * a.nk;
  if (gid < per_model * a.B) {
  const int b = (int)(gid / per_model);
  const long long rem = gid - (long long)b * per_model;
  const int s = (int)(rem / a.nk), row = (int)(rem - (long long)s * a.nk);
  const int smp = b / a.F, P = a.P;
  // pair s -> (u, v), u <= v, row-major: rows start at off(u) = u P - u(u-1)/2; invert with a float guess + fix-up
  int u = (int)(((2. * P + 1.) - sqrt((2. * P + 1.) * (2. * P + 1.) - 8. * (double)s)) * 0.5);
  u = u < 0 ? 0 : (u > P - 1 ? P - 1 : u);
  while (u > 0 && u * P - u * (u - 1) / 2 > s) --u;
  while ((u + 1) * P - (u + 1) * u / 2 <= s) ++u;
  const int v = u + (s - (u * P - u * (u - 1) / 2));
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
  if (a.mass[u]) fu += Aterm * (a.tscal[((size_t)b * P + u) * HM_GNTS + 2] * a.grid[u][goff]);
  if (a.mass[v]) fv += Aterm * (a.tscal[((size_t)b * P + v) * HM_GNTS + 2] * a.grid[v][goff]);
  const double Iu = a.simple_twohalo ? a.tscal[((size_t)b * P + u) * HM_GNTS + 1] : fu * rhom;
  const double Iv = a.simple_twohalo ? a.tscal[((size_t)b * P + v) * HM_GNTS + 1] : fv * rhom;
  const double p2h = __dmul_rn(__dadd_rn(__dmul_rn(Iu, Iv), 0.), a.Pk_lin[(size_t)smp * a.nk + row]);   // :524
  double p1h = f1 * rhom;                                                                              // :532
  if (u == v && a.discrete[u]) {                                                                       // :484-491
    const double sn = a.tscal[((size_t)b * P + u) * HM_GNTS];
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
  if (stages & 1) {
    hm_gram_weights_kernel<<<dim3(L.nchunks, (unsigned)L.B), 256, 0, st>>>(w);
    HM_LAUNCH_CHECK();
    hm_gram_tracer_kernel<<<dim3(L.nchunks, L.P, (unsigned)L.B), 256, 0, st>>>(w);
    HM_LAUNCH_CHECK();
    hm_gram_scalars_kernel<<<dim3(L.P, (unsigned)L.B), 256, 0, st>>>(w);
    HM_LAUNCH_CHECK();
  }

  hm_gram_args g;
  g.nk = pr->nk; g.nM = pr->nM; g.nMp = L.nMp; g.P = L.P; g.NB = L.NB; g.PT = L.PT; g.NW = L.NW;
  g.F = pr->models_per_sample; g.ntiles = L.ntiles; g.nparts = L.nparts; g.MP = L.MP; g.gw_stride = L.gw_stride;
  for (int p = 0; p < HM_GMAXP; ++p) g.grid[p] = (p < L.P) ? pr->tracers[p].grid_dev : nullptr;
  g.gw = wsd; g.partial = wsd + L.partial_offset;
  // the kernel is instantiated for block grids of 2, 4 and 8 (16, 32, 64 tracer rows); pad up to the next one
  const int NBT = (L.NB <= 2) ? 2 : (L.NB <= 4) ? 4 : 8;
  const size_t smem = (size_t)HM_GNBUF * L.RK * (8 * NBT) * HM_GLD * sizeof(double);
  static hm_once_per_device once;
  if (once.needed()) {
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gram_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gram_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    HM_CUDA_TRY(cudaFuncSetAttribute(hm_gram_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  }
  g.nitems = (long long)L.B * L.ntiles * L.nparts;
  HM_REQUIRE(g.nitems < (1LL << 31), "too many work items for one launch");
  const long long grid = g.nitems < hm_num_sms() ? g.nitems : hm_num_sms();
  if (stages & 2) {
    if (NBT == 2) hm_gram_kernel<2, 2><<<(unsigned)grid, HM_GTHREADS, smem, st>>>(g);
    else if (NBT == 4) hm_gram_kernel<2, 4><<<(unsigned)grid, HM_GTHREADS, smem, st>>>(g);
    else hm_gram_kernel<2, 8><<<(unsigned)grid, HM_GTHREADS, smem, st>>>(g);
    HM_LAUNCH_CHECK();
  }
  if (!(stages & 4)) return HM_OK;

  hm_gram_eargs e;
  e.nk = pr->nk; e.nM = pr->nM; e.nMp = L.nMp; e.P = L.P; e.S = L.S; e.NW = L.NW; e.F = pr->models_per_sample;
  e.nparts = L.nparts; e.B = (int)L.B;
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
  const long long nthreads = (long long)L.B * L.S * pr->nk;
  hm_gram_epilogue_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(e);
  HM_LAUNCH_CHECK();
  return HM_OK;
}
