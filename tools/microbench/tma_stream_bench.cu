// Microbenchmark: HBM bandwidth a GEMM-style TMA producer reaches when it streams the A operand of a row-major
// [M][K] bf16 matrix as {64 columns (128 B) x 128 rows} SWIZZLE_128B boxes, k-blocks innermost - as a function of the
// row length K (the stride between the 128-byte pieces of one box), the number of 16 KB stages in flight and the
// tensor map's L2 promotion.  Same total bytes in every case.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hse_facerec_tf_b200/csrc tools/microbench/tma_stream_bench.cu -o tools/microbench/tma_stream_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include "ptx.cuh"
using namespace hfr;

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmA, int num_mb, int num_kb,
                                                       int stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sBar = base + stages * 16384;
  const int warp = uniform_warp_idx();
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(sBar + 8 * s, 1);
      mbar_init(sBar + 128 + 8 * s, 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  int stage = 0;
  uint32_t phase = 0;
  if (warp == 0) {
    for (int mb = blockIdx.x; mb < num_mb; mb += gridDim.x)
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(sBar + 128 + 8 * stage, phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(sBar + 8 * stage, 16384);
          tma_load_2d(base + stage * 16384, &tmA, sBar + 8 * stage, kb * 64, mb * 128);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
  } else {
    for (int mb = blockIdx.x; mb < num_mb; mb += gridDim.x)
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(sBar + 8 * stage, phase);
        if (elect_one()) mbar_arrive(sBar + 128 + 8 * stage);
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  const size_t total_elems = (size_t)50176 * 1024 * 4;  // 411 MB: larger than L2
  void* d;
  cudaMalloc(&d, total_elems * 2);
  cudaMemset(d, 0, total_elems * 2);
  void* flush;
  cudaMalloc(&flush, 256u << 20);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const CUtensorMapL2promotion promos[3] = {CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  const char* pname[3] = {"none", "128B", "256B"};
  for (int pi = 0; pi < 3; ++pi)
    for (int K : {64, 256, 512, 1024, 2048, 4096})
      for (int stages : {3, 5, 10}) {
        const uint64_t M = total_elems / K;
        CUtensorMap m;
        const cuuint64_t dims[2] = {(cuuint64_t)K, M};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, promos[pi], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
          cudaMemsetAsync(flush, it, 256u << 20);
          cudaEventRecord(e0);
          stream_kernel<<<148, 64, stages * 16384 + 2048>>>(m, (int)(M / 128), K / 64, stages);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        printf("promo=%s K=%4d (row %5d B) stages=%2d : %7.1f us  %7.1f GB/s\n", pname[pi], K, K * 2, stages, best * 1e3,
               total_elems * 2 / (best * 1e-3) / 1e9);
      }
  return 0;
}
