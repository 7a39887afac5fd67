// hm_common.cuh -- shared device/host helpers of libhalob200 (sm_100a, fp64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/halob200.h"
#include "hm_tables.h"

// ---- error plumbing (never throw across the C ABI) -------------------------------------------------------------
void hm_set_error(const char* fmt, ...);
int hm_check_device();   // HM_OK or HM_ERR_NODEVICE (cached)

#define HM_CUDA_TRY(expr)                                                                            \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      hm_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e));     \
      return HM_ERR_CUDA;                                                                            \
    }                                                                                                \
  } while (0)

#define HM_REQUIRE(cond, msg)                                  \
  do {                                                         \
    if (!(cond)) {                                             \
      hm_set_error("%s: %s", __func__, msg);                   \
      return HM_ERR_INVALID;                                   \
    }                                                          \
  } while (0)

#define HM_LAUNCH_CHECK()                                                                    \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      hm_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return HM_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

static inline int hm_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remember on which devices of this
// process a kernel has been configured (a repeated call from a racing thread is harmless)
struct hm_once_per_device {
  unsigned long long mask = 0ull;
  bool needed() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;
    if ((mask >> dev) & 1ull) return false;
    mask |= 1ull << dev;
    return true;
  }
};

// ---- streaming loads: (k,M) grids are read exactly once, keep them out of L1 ---------------------------------
__device__ __forceinline__ double2 hm_ldg_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double hm_ldg_stream1(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// ---- mass function g(nu) and linear bias b(nu) ---------------------------------------------------------------
// Operation order follows pyhalomodel.py:217-284.  Powers x^y are evaluated as exp(y ln x) with one shared ln(nu)
// per evaluation instead of CUDA's pow (a ~200-instruction dependent chain each, and these kernels are
// latency-bound): |y ln x| <= ~8 here, so the result is within ~1e-15 relative of the correctly rounded power --
// the g(nu), b(nu) tests hold 2e-14 against the reference's libm values.
__device__ __forceinline__ double hm_powl(double lnx, double y) { return exp(y * lnx); }

__device__ __forceinline__ double hm_g_nu(const hm_model_desc& d, double nu) {
  const double* P = d.par;
  switch (d.family) {
    case HM_PS74:
      return sqrt(2. / M_PI) * exp(-(nu * nu) / 2.);                              // :223
    case HM_ST99:
    case HM_SMT01:
    case HM_DESPALI16: {
      const double A = P[0], q = P[1], p = P[2];
      const double nu2 = nu * nu;
      return A * (1. + hm_powl(log(q * nu2), -p)) * exp(-q * nu2 / 2.);            // :228
    }
    case HM_TINKER10:
    case HM_TINKER10_PBS: {
      const double f1 = 1. + hm_powl(log(P[1] * nu), -2. * P[3]);                  // :235
      const double f2 = hm_powl(log(nu), 2. * P[4]);                               // :236
      const double f3 = exp(-P[2] * (nu * nu) / 2.);                               // :237
      return P[0] * f1 * f2 * f3;
    }
  }
  return nan("");
}

__device__ __forceinline__ double hm_b_nu(const hm_model_desc& d, double nu) {
  const double* P = d.par;
  const double dc = d.dc;
  switch (d.family) {
    case HM_PS74:
      return 1. + (nu * nu - 1.) / dc;                                            // :248
    case HM_ST99:
    case HM_DESPALI16: {
      const double q = P[1], p = P[2];
      const double qn2 = q * (nu * nu);
      return 1. + (qn2 - 1. + 2. * p / (1. + hm_powl(log(qn2), p))) / dc;          // :252
    }
    case HM_SMT01: {
      const double a = P[3], b = P[4], c = P[5];
      const double sa = sqrt(a);
      const double anu2 = a * (nu * nu);
      const double la = log(anu2);
      const double f1 = sa * anu2;
      const double f2 = sa * b * hm_powl(la, 1. - c);
      const double f3 = hm_powl(la, c);
      const double f4 = f3 + b * (1. - c) * (1. - c / 2.);
      return 1. + (f1 + f2 - f3 / f4) / (dc * sa);                                 // :254-262
    }
    case HM_TINKER10: {
      const double ln = log(nu);
      const double nua = hm_powl(ln, P[6]);
      const double fA = P[5] * nua / (nua + hm_powl(log(dc), P[6]));               // :279
      const double fB = P[7] * hm_powl(ln, P[8]);
      const double fC = P[9] * hm_powl(ln, P[10]);
      return 1. - fA + fB + fC;                                                    // :282
    }
    case HM_TINKER10_PBS: {
      const double f1 = (P[2] * (nu * nu) - (1. + 2. * P[4])) / dc;                // :269
      const double f2 = (2. * P[3] / dc) / (1. + hm_powl(log(P[1] * nu), 2. * P[3]));   // :270
      return 1. + f1 + f2;
    }
  }
  return nan("");
}

// ---- g(nu) and b(nu) g(nu) together, for kernels that evaluate them on every mass of a sweep -------------------
// Same formulas as hm_g_nu / hm_b_nu with the transcendental work shared: ONE ln(nu) per mass serves every family
// (ln(q nu^2) = ln q + 2 ln nu, ln(beta nu) = ln beta + ln nu; the parameter logarithms are per-model constants,
// hm_family_consts), reciprocal powers come from a division instead of a second exp ((q nu^2)^p = 1/(q nu^2)^-p) and
// products of powers of e share one exp.  Each regrouping moves the result by a few 1e-16 relative (|exponent| * eps
// for the merged exponentials, i.e. up to ~1e-14 where g ~ e^-100 contributes nothing); 13 exp + 1 log per mass for the
// five families instead of 18 exp + 9 log.  Spectra computed with these weights hold rtol 1e-10 against the reference
// with four orders of magnitude to spare (tests/test_gpu_parity.py).
struct hm_family_consts {
  double ln0;      // ST/SMT/Despali: ln q          Tinker: ln beta
  double ln1;      // SMT: ln a
  double pw;       // Tinker: dc^a
  double inv_dc;   // 1 / dc
};

__device__ __forceinline__ hm_family_consts hm_make_family_consts(const hm_model_desc& d) {
  hm_family_consts c;
  const double* P = d.par;
  c.ln0 = c.ln1 = c.pw = 0.;
  c.inv_dc = 1. / d.dc;
  switch (d.family) {
    case HM_ST99:
    case HM_DESPALI16: c.ln0 = log(P[1]); break;
    case HM_SMT01: c.ln0 = log(P[1]); c.ln1 = log(P[3]); break;
    case HM_TINKER10: c.ln0 = log(P[1]); c.pw = exp(P[6] * log(d.dc)); break;
    case HM_TINKER10_PBS: c.ln0 = log(P[1]); break;
  }
  return c;
}

// g = g(nu), bg = b(nu) g(nu);  lnnu = ln(nu)
__device__ __forceinline__ void hm_gb_nu_shared(const hm_model_desc& d, const hm_family_consts& c, double nu, double lnnu,
                                                double& g, double& bg) {
  const double* P = d.par;
  const double nu2 = nu * nu;
  switch (d.family) {
    case HM_PS74: {
      g = sqrt(2. / M_PI) * exp(-nu2 / 2.);
      bg = (1. + (nu2 - 1.) * c.inv_dc) * g;
      return;
    }
    case HM_ST99:
    case HM_DESPALI16: {
      const double A = P[0], q = P[1], p = P[2];
      const double e1 = exp(-p * (c.ln0 + 2. * lnnu));                 // (q nu^2)^-p
      g = A * (1. + e1) * exp(-q * nu2 / 2.);
      bg = (1. + (q * nu2 - 1. + 2. * p * e1 / (e1 + 1.)) * c.inv_dc) * g;   // 2p/(1+(q nu^2)^p)
      return;
    }
    case HM_SMT01: {
      const double A = P[0], q = P[1], p = P[2], a = P[3], b = P[4], cc = P[5];
      const double e1 = exp(-p * (c.ln0 + 2. * lnnu));
      g = A * (1. + e1) * exp(-q * nu2 / 2.);
      const double sa = sqrt(a), anu2 = a * nu2;
      const double f3 = exp(cc * (c.ln1 + 2. * lnnu));                 // (a nu^2)^c
      const double f2 = sa * b * (anu2 / f3);                          // sqrt(a) b (a nu^2)^(1-c)
      const double f4 = f3 + b * (1. - cc) * (1. - cc / 2.);
      bg = (1. + (sa * anu2 + f2 - f3 / f4) * (c.inv_dc / sa)) * g;
      return;
    }
    case HM_TINKER10:
    case HM_TINKER10_PBS: {
      const double e1 = exp(-2. * P[3] * (c.ln0 + lnnu));              // (beta nu)^(-2 phi)
      g = P[0] * (1. + e1) * exp(2. * P[4] * lnnu - P[2] * nu2 / 2.);  // nu^(2 eta) exp(-gamma nu^2/2)
      if (d.family == HM_TINKER10) {
        const double nua = exp(P[6] * lnnu);
        bg = (1. - P[5] * nua / (nua + c.pw) + P[7] * exp(P[8] * lnnu) + P[9] * exp(P[10] * lnnu)) * g;
      } else {
        const double f1 = (P[2] * nu2 - (1. + 2. * P[4])) * c.inv_dc;
        const double f2 = (2. * P[3] * c.inv_dc) * e1 / (e1 + 1.);      // (2 phi/dc)/(1+(beta nu)^(2 phi))
        bg = (1. + f1 + f2) * g;
      }
      return;
    }
  }
  g = bg = nan("");
}

// ---- warp / block reductions (fixed order => bitwise reproducible run to run) --------------------------------
__device__ __forceinline__ double hm_warp_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Sum `v` over the whole CTA; every thread gets the result.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double hm_block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = hm_warp_sum(v);
  __syncthreads();                      // protect scratch against the previous use
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = (lane < nwarps) ? scratch[lane] : 0.;
  return hm_warp_sum(t);
}

// v[0..32) per lane  ->  lane l holds the warp-wide sum of v[l]  (31 shuffles instead of 160).
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


static __constant__ double c_gl_x[16] = HM_GL16_X_INIT;
static __constant__ double c_gl_w[16] = HM_GL16_W_INIT;

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


// The same node with the shared-transcendental evaluation of g and b g (hm_gb_nu_shared: one ln(nu) serves every power,
// about half the instructions of hm_g_nu * hm_b_nu); for the kernels that integrate A for every model of a sweep.
__device__ __forceinline__ double hm_lowmass_node_shared(const hm_model_desc& d, const hm_family_consts& c, double nu0, int t) {
  if (t >= HM_GL_NODES) return 0.;
  const int panel = t >> 4, node = t & 15;
  const double unit = 40. / (double)((1 << HM_GL_PANELS) - 1);
  const double a = nu0 + unit * (double)((1 << panel) - 1);
  const double b = nu0 + unit * (double)((1 << (panel + 1)) - 1);
  const double half = 0.5 * (b - a);
  const double x = 0.5 * (a + b) + half * c_gl_x[node];
  double g, bg;
  hm_gb_nu_shared(d, c, x, log(x), g, bg);
  return half * c_gl_w[node] * bg;
}


// ---- shared-memory addresses and mbarriers (inline PTX) ---------------------------------------------------------
__device__ __forceinline__ uint32_t hm_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void hm_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void hm_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void hm_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// non-blocking probe: 1 when the phase with this parity has completed
__device__ __forceinline__ uint32_t hm_mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
__device__ __forceinline__ void hm_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
