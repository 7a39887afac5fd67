// Microbenchmark: the consumer loop of hm_quad_rows_kernel<2,5> (BASELINE config 3) on a static shared-memory
// stage -- no producer, no barriers -- to separate the loop's own ceiling from the hand-off and memory effects.
// A lane owns two rows; per pair of mass columns it issues 4 16-byte loads of its grid values, 25 broadcast 16-byte
// loads of weights, 12 DMUL and 100 DFMA.  Variants: 0 = the full loop, 1 = without the broadcast weight loads
// (weights from registers), 2 = FP64 only.  Reported as the HBM rate the loop could absorb: 2 KB per warp iteration.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/rows_consumer tools/micro/rows_consumer.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int P = 2, FM = 5, NW = 5, NV = FM * NW, RT = 64, CS = 32, ROWB = CS * 8 + 16;
constexpr int DATA_BYTES = RT * P * ROWB, STAGE_BYTES = DATA_BYTES + NV * CS * 8;

template <int VARIANT>
__global__ void __launch_bounds__(384, 1) loop_k(double* out, int iters, int nwarps) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* sd = reinterpret_cast<double*>(smem);
  for (int i = threadIdx.x; i < STAGE_BYTES / 8; i += blockDim.x) sd[i] = 1e-3 * (1 + i % 89);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= nwarps) return;
  const int cbeg = (warp % 8) * (CS / 8);
  const unsigned char* stage = smem;
  const unsigned char* wsm = stage + DATA_BYTES;
  const unsigned my_row = (unsigned)lane * ROWB;
  double acc[2][NV];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[r][j] = 0.;
  double2 wreg = make_double2(1e-3 * lane, 2e-3);
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int c = cbeg; c < cbeg + CS / 8; c += 2) {
      double2 x[2][P];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (VARIANT == 2) x[r][p] = make_double2(wreg.x + r + it, wreg.y + p + c);
          else x[r][p] = *reinterpret_cast<const double2*>(stage + my_row + (unsigned)(p * RT + r * 32) * ROWB + c * 8);
        }
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double2 v[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          if (i < P) v[r] = x[r][i];
          else if (i < 2 * P) v[r] = make_double2(x[r][i - P].x * x[r][i - P].x, x[r][i - P].y * x[r][i - P].y);
          else v[r] = make_double2(x[r][0].x * x[r][1].x, x[r][0].y * x[r][1].y);
        }
#pragma unroll
        for (int f = 0; f < FM; ++f) {
          double2 w2;
          if (VARIANT == 0) w2 = *reinterpret_cast<const double2*>(wsm + ((f * NW + i) * CS + c) * 8);
          else w2 = make_double2(wreg.x + f, wreg.y + i);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            acc[r][f * NW + i] = fma(w2.x, v[r].x, acc[r][f * NW + i]);
            acc[r][f * NW + i] = fma(w2.y, v[r].y, acc[r][f * NW + i]);
          }
        }
      }
    }
  }
  double t = 0.;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < NV; ++j) t += acc[r][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int VARIANT>
static void run(const char* name, int nwarps, double* out) {
  const int iters = 4000, sms = 148;
  cudaFuncSetAttribute(loop_k<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES);
  loop_k<VARIANT><<<sms, 384, STAGE_BYTES>>>(out, 10, nwarps);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  loop_k<VARIANT><<<sms, 384, STAGE_BYTES>>>(out, iters, nwarps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double pairs = (double)sms * nwarps * iters * (CS / 8 / 2);      // warp iterations over a pair of columns
  const double bytes = pairs * RT * 2 * P * 8;                            // grid bytes a pair consumes
  const double flops = pairs * 32 * (100 + 12) * 2;
  printf("%-28s %2d warps/SM: %8.3f ms  %6.0f clk per pair-iteration per warp  -> absorbs %5.2f TB/s, %5.1f TFLOP/s\n",
         name, nwarps, ms, ms * 1e-3 * 1.965e9 / (iters * (CS / 8 / 2)), bytes / (ms * 1e-3) / 1e12, flops / (ms * 1e-3) / 1e12);
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 384 * sizeof(double));
  for (int nw : {4, 8, 12}) {
    run<0>("full loop", nw, out);
    run<1>("weights in registers", nw, out);
    run<2>("FP64 only", nw, out);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
