// Microbenchmark: what the memory system delivers to the TMA feed of hm_quad_rows_kernel, as a function of the box
// shape.  148 persistent CTAs walk (sample, row tile) items of a [rows][1024] float64 tensor pair along the mass axis
// exactly like the sweep kernel's producer -- a ring of stages, one thread issuing cp.async.bulk.tensor.2d, completion
// on mbarriers -- but the consumer is a single warp that releases a stage as soon as it has landed (no arithmetic).
// A stage is always 2 tensors x 2 boxes = 32 KB; the box is R rows x C columns with R * C = 1024 doubles:
//   64 x 16 (the kernel's shape: 128-byte row segments, 256 contiguous bytes per row and stage), 32 x 32, 16 x 64, 8 x 128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tma_boxes tools/micro/tma_boxes.cu -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t par) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(par) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

constexpr int STAGE = 32768;

__global__ void __launch_bounds__(64, 1) feed(const __grid_constant__ CUtensorMap tA, const __grid_constant__ CUtensorMap tB,
                                              int R, int C, int NS, int nitems, int nM, double* sink) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)NS * STAGE);
  const uint32_t ring = saddr(sm), full = saddr(bars), empty = saddr(bars + NS);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nst = nM / (2 * C);
  int s = 0; uint32_t par = 0;
  if (threadIdx.x == 32) {          // producer
    for (int it = blockIdx.x; it < nitems; it += gridDim.x)
      for (int st = 0; st < nst; ++st) {
        mbar_wait(empty + 8 * s, par ^ 1u);
        mbar_expect(full + 8 * s, STAGE);
        const uint32_t dst = ring + s * STAGE;
        for (int b = 0; b < 2; ++b) {
          tma2d(dst + b * 8192, &tA, st * 2 * C + b * C, it * R, full + 8 * s);
          tma2d(dst + 16384 + b * 8192, &tB, st * 2 * C + b * C, it * R, full + 8 * s);
        }
        if (++s == NS) { s = 0; par ^= 1u; }
      }
  } else if (threadIdx.x < 32) {    // consumer: touch one word per lane, release
    double acc = 0.;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x)
      for (int st = 0; st < nst; ++st) {
        mbar_wait(full + 8 * s, par);
        acc += reinterpret_cast<const double*>(sm + (size_t)s * STAGE)[threadIdx.x * 128];
        __syncwarp();
        if (threadIdx.x == 0) mbar_arrive(empty + 8 * s);
        if (++s == NS) { s = 0; par ^= 1u; }
      }
    if (acc == 12345.678) sink[0] = acc;
  }
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int nM = 1024, rows = 512 * 256;          // 512 samples x 256 wavenumbers: 1.07 GB per tensor
  double *A, *B, *sink;
  CK(cudaMalloc(&A, (size_t)rows * nM * 8)); CK(cudaMalloc(&B, (size_t)rows * nM * 8)); CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(A, 0, (size_t)rows * nM * 8)); CK(cudaMemset(B, 0, (size_t)rows * nM * 8));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  enc_fn enc = (enc_fn)p;
  CK(cudaFuncSetAttribute(feed, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int shapes[][2] = {{64, 16}, {32, 32}, {16, 64}, {8, 128}};
  for (int l2 = 0; l2 < 2; ++l2)
    for (auto& sh : shapes)
      for (int NS = 3; NS <= 6; NS += 1) {
        const int R = sh[0], C = sh[1];
        CUtensorMap tA, tB;
        const cuuint64_t dims[2] = {(cuuint64_t)nM, (cuuint64_t)rows};
        const cuuint64_t strides[1] = {(cuuint64_t)nM * 8};
        const cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)R};
        const cuuint32_t estr[2] = {1, 1};
        const CUtensorMapL2promotion prom = l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
        if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
            enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, B, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
          printf("encode failed for %d x %d\n", R, C);
          continue;
        }
        const int nitems = rows / R;
        const size_t smem = (size_t)NS * STAGE + 256;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        feed<<<148, 64, smem>>>(tA, tB, R, C, NS, nitems, nM, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 5; ++i) feed<<<148, 64, smem>>>(tA, tB, R, C, NS, nitems, nM, sink);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("box %3d rows x %3d cols (%4d B segments), %d stages of 32 KB, L2 promotion %s: %.1f us per pass, %.0f GB/s\n", R, C,
               C * 8, NS, l2 ? "256B" : "none", ms / 5 * 1e3, 2.0 * rows * nM * 8 / (ms / 5 * 1e-3) / 1e9);
      }
  return 0;
}
