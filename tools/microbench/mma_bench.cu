// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16 bf16, M=128, K=16) for different N and different
// shared-memory operand layouts, issued back-to-back by one thread into one TMEM accumulator (the dependent chain a
// GEMM k-loop produces).  Informs tile-shape decisions in gemm_tc.cuh / conv_window.cuh.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hse_facerec_tf_b200/csrc tools/microbench/mma_bench.cu -o tools/microbench/mma_bench
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
#include "conv_window.cuh"
using namespace hfr;

// layout modes
//  0: SWIZZLE_128B K-major, 4 k-steps advance inside a 128-byte row (what gemm_tc uses)
//  1: no-swizzle, 16-byte rows: SBO = 128 (8-row groups contiguous), LBO = 2048 (second k-chunk in another plane), aligned
//  2: no-swizzle, SBO = 176 (window with ww = 11), start address offset by 16*tap -> misaligned groups
//  3: SWIZZLE_32B (32-byte rows), one k-step per tile
template <int N>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int mode, int reps, long long* out_cycles, int n_distinct) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
  if (threadIdx.x == 32) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  // fill operands with small finite values
  for (int i = threadIdx.x; i < 130 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3C003C00u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc(1, 128, N);
    const uint32_t sA = base, sB = base + 64 * 1024;
    uint32_t phase = 0;
    long long best = 1ll << 60;
    // descriptors are built outside the timed loop: a single issuing thread runs ~1 instruction / 4-6 cycles, so
    // anything but the MMA itself in the loop body hides the tensor-core time
    uint64_t ad[8], bd[8];
#pragma unroll
    for (int j0 = 0; j0 < 8; ++j0) {
      const int j = j0 % n_distinct;
      if (mode == 0) {
        ad[j0] = umma_desc_sw128(sA + (j >> 2) * 16384) + 2u * (j & 3);
        bd[j0] = umma_desc_sw128(sB + (j >> 2) * (N * 128)) + 2u * (j & 3);
      } else if (mode == 1) {
        ad[j0] = umma_desc_noswz(sA + j * 128, 2048 + 4096, 128);
        bd[j0] = umma_desc_noswz(sB + j * (N * 32), N * 16, 128);
      } else if (mode == 2) {
        ad[j0] = umma_desc_noswz(sA + (j / 4) * 176 + (j % 4) * 16, 3344, 176);
        bd[j0] = umma_desc_noswz(sB + j * (N * 32), N * 16, 128);
      } else {
        ad[j0] = umma_desc_sw32(sA + j * 4096);
        bd[j0] = umma_desc_sw32(sB + j * (N * 32));
      }
    }
    for (int trial = 0; trial < 5; ++trial) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int j0 = 0; j0 < 8; ++j0) umma<false>(tmem, ad[j0], bd[j0], idesc, (r | j0) != 0);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out_cycles[blockIdx.x] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N>
static void run(int mode, int reps, int n_distinct, int grid) {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  cudaFuncSetAttribute(mma_bench_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mma_bench_kernel<N><<<grid, 128, 200 * 1024>>>(mode, reps, d, n_distinct);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d mode=%d distinct=%2d grid=%3d reps=%d : %8.1f cycles/MMA (ideal %d)  %s\n", N, mode, n_distinct, grid, reps,
         (double)mx / reps, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int reps = 512;
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int nd : {1, 8}) {
        run<64>(mode, reps, nd, grid);
        run<128>(mode, reps, nd, grid);
        run<256>(mode, reps, nd, grid);
      }
    }
  }
  return 0;
}
