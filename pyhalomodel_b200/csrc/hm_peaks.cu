// hm_peaks.cu -- FP64 roofline denominators measured on the device the library runs on: DFMA issue throughput of
// the FP64 pipe and DMMA (mma.sync m8n8k4 f64, the tensor-core path of hm_gram.cu) throughput.  MEASURED_PEAKS.json
// holds HBM and bf16 figures only, so bench.py calls this in the same run that reports FP64-bound kernels.
#include "hm_common.cuh"

__global__ void hm_peak_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x, a1 = 1.1, a2 = 2.2, a3 = 3.3, a4 = 4.4, a5 = 5.5, a6 = 6.6, a7 = 7.7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void hm_peak_dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.;
  const double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0.;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Synchronises `stream`.  Best of the occupancies 8, 16 and 32 warps per SM, one CTA per SM; flops counted as 2 per
// FMA (DMMA m8n8k4: 256 FMAs per warp instruction).
extern "C" int hm_measure_fp64_peaks(double* dfma_tflops_host, double* dmma_tflops_host, void* stream) {
  int rc = hm_check_device();
  if (rc) return rc;
  HM_REQUIRE(dfma_tflops_host && dmma_tflops_host, "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = hm_num_sms(), iters = 20000;
  double* out = nullptr;
  HM_CUDA_TRY(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(double)));
  cudaEvent_t e0, e1;
  HM_CUDA_TRY(cudaEventCreate(&e0));
  HM_CUDA_TRY(cudaEventCreate(&e1));
  double best_f = 0., best_m = 0.;
  for (int warps = 8; warps <= 32; warps *= 2) {
    float ms = 0.f;
    hm_peak_dfma_kernel<<<sms, warps * 32, 0, st>>>(out, 200);
    cudaEventRecord(e0, st);
    hm_peak_dfma_kernel<<<sms, warps * 32, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = (double)sms * warps * 32 * iters * 8 * 2 / ms / 1e9;
    if (tf > best_f) best_f = tf;
    hm_peak_dmma_kernel<<<sms, warps * 32, 0, st>>>(out, 200);
    cudaEventRecord(e0, st);
    hm_peak_dmma_kernel<<<sms, warps * 32, 0, st>>>(out, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double tm = (double)sms * warps * iters * 8 * 512.0 / ms / 1e9;
    if (tm > best_m) best_m = tm;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  HM_LAUNCH_CHECK();
  *dfma_tflops_host = best_f;
  *dmma_tflops_host = best_m;
  return HM_OK;
}
