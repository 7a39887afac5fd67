// Stand-alone consumer of libhalob200.so: plain C++ host code, cudaMalloc'd buffers, no Python or torch in the loop.
// Reads one problem from a binary file written by tests/test_gpu_cabi.py, runs hm_power_spectrum through the C ABI
// (include/halob200.h) and writes the spectra back; the Python test compares them with the oracle.
//
// File layout (little endian): int32 nk, nM, P, family; double dc, rhom, par[12], k_trunc;
//   int32 flags[3] (simple_twohalo, subtract_shotnoise, correct_discrete); then doubles k[nk], M[nM], sigma[nM],
//   Pk_lin[nk]; per tracer: int32 mode, mass, discrete, has_variance; double norm; grid[nk*nM]; amplitude[nM];
//   variance[nM] if has_variance.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "halob200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

template <class T> static bool rd(FILE* f, T* p, size_t n) { return fread(p, sizeof(T), n, f) == n; }

static double* upload(const std::vector<double>& v) {
  double* d = nullptr;
  if (cudaMalloc(&d, v.size() * sizeof(double)) != cudaSuccess) return nullptr;
  cudaMemcpy(d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
  return d;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: cabi_host in.bin out.bin\n"); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 1; }
  int32_t hdr[4], flags[3];
  double dc, rhom, par[12], k_trunc;
  if (!rd(f, hdr, 4) || !rd(f, &dc, 1) || !rd(f, &rhom, 1) || !rd(f, par, 12) || !rd(f, &k_trunc, 1) || !rd(f, flags, 3)) return 1;
  const int nk = hdr[0], nM = hdr[1], P = hdr[2];
  std::vector<double> k(nk), M(nM), sig(nM), Pk(nk);
  if (!rd(f, k.data(), nk) || !rd(f, M.data(), nM) || !rd(f, sig.data(), nM) || !rd(f, Pk.data(), nk)) return 1;

  if (hm_abi_version() != HM_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 3; }
  hm_problem pr = {};
  pr.nk = nk; pr.nM = nM; pr.n_tracers = P; pr.n_samples = 1; pr.models_per_sample = 1;
  pr.simple_twohalo = flags[0]; pr.subtract_shotnoise = flags[1]; pr.correct_discrete = flags[2];
  pr.k_trunc = k_trunc;
  pr.k_dev = upload(k); pr.M_dev = upload(M); pr.sigma_dev = upload(sig); pr.Pk_lin_dev = upload(Pk);
  pr.model_host.family = hdr[3]; pr.model_host.dc = dc; pr.model_host.rhom = rhom;
  for (int i = 0; i < 12; ++i) pr.model_host.par[i] = par[i];
  for (int p = 0; p < P; ++p) {
    int32_t t[4]; double norm;
    if (!rd(f, t, 4) || !rd(f, &norm, 1)) return 1;
    std::vector<double> grid((size_t)nk * nM), amp(nM), var(nM);
    if (!rd(f, grid.data(), grid.size()) || !rd(f, amp.data(), nM)) return 1;
    if (t[3] && !rd(f, var.data(), nM)) return 1;
    hm_tracer& tr = pr.tracers[p];
    tr.grid_dev = upload(grid); tr.amplitude_dev = upload(amp);
    tr.variance_dev = t[3] ? upload(var) : nullptr;
    tr.normalisation_dev = upload(std::vector<double>(1, norm));
    tr.grid_mode = t[0]; tr.mass_tracer = t[1]; tr.discrete_tracer = t[2];
  }
  fclose(f);

  const int S = P * (P + 1) / 2;
  const size_t ws_bytes = hm_power_spectrum_workspace(&pr);
  void* ws = nullptr; double *out = nullptr, *low = nullptr;
  CK(cudaMalloc(&ws, ws_bytes)); CK(cudaMalloc(&out, (size_t)S * 3 * nk * sizeof(double))); CK(cudaMalloc(&low, sizeof(double)));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  const int rc = hm_power_spectrum(&pr, out, low, ws, ws_bytes, st);
  if (rc != HM_OK) { fprintf(stderr, "hm_power_spectrum: %d %s\n", rc, hm_last_error()); return 4; }
  CK(cudaStreamSynchronize(st));
  std::vector<double> res((size_t)S * 3 * nk + 1);
  CK(cudaMemcpy(res.data(), out, (size_t)S * 3 * nk * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&res[(size_t)S * 3 * nk], low, sizeof(double), cudaMemcpyDeviceToHost));
  // error behaviour across the ABI: a bad problem returns a status and a message, never aborts
  hm_problem bad = pr; bad.n_tracers = 0;
  if (hm_power_spectrum(&bad, out, low, ws, ws_bytes, st) != HM_ERR_INVALID || !hm_last_error()[0]) return 5;
  if (hm_power_spectrum(&pr, out, low, ws, 16, st) != HM_ERR_WORKSPACE) return 6;
  FILE* g = fopen(argv[2], "wb");
  fwrite(res.data(), sizeof(double), res.size(), g);
  fclose(g);
  printf("cabi_host ok: %d spectra of %d wavenumbers, A = %.15g\n", S, nk, res.back());
  return 0;
}
