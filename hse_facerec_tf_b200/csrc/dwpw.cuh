// Fused MobileNet block for sm_100a:  depthwise 3x3 (+bias, ReLU6)  ->  pointwise 1x1 (+bias, ReLU6), stride 1, bf16.
//
// The depthwise output never exists in HBM (or L2): CUDA-core warps compute it from a TMA-loaded halo window straight
// into the shared-memory A operand of the pointwise tcgen05 GEMM, in the SWIZZLE_128B K-major layout the tensor core
// reads.  Replaces the graph nodes conv_dw_N/depthwise .. conv_pw_N_relu of the reference's frozen MobileNet
// (facerec_test.py:120 sess.run) - two kernels and one full activation round trip per block in the unfused path.
//
//   unit          = (128-pixel tile, BLOCK_N output channels); tile = tile_n images x tile_h rows x 8 columns
//   warp 0        TMA producer: per 64-channel K-block the input window [tile_n][WH][WW][64ch] and the pw weight tile
//   warp 1        tcgen05.mma issuer (M=128 x N=BLOCK_N, fp32 accumulators in TMEM, 2 stages: unit parity)
//   warp 2        TMEM allocator
//   warps 4..11   workers: depthwise producer for unit i, then epilogue (TMEM -> bias/ReLU6 -> bf16 -> TMA store) of
//                 unit i-1, so the tensor core runs unit i while unit i-1 drains
#pragma once
#include "conv_window.cuh"
#include "ptx.cuh"

namespace hfr {

struct DwPwParams {
  int Cin, Cout;
  int tiles_x, tiles_y, img_groups;   // tiles per image row/column, batch / tile_n
  int tile_h, tile_n;                 // tile_h * tile_n == 16
  int ww, wh;                         // window size in pixels per image (8 + 2, tile_h + 2)
  int num_kb, n_blocks, num_units;
  const float* dw_w;                  // [9][Cin]
  const float* dw_b;                  // [Cin]
  const float* pw_b;                  // [Cout] or null
  int dw_act, pw_act;
};

constexpr int kDwVecStride = 84;  // floats per 8-channel vector in smem: 9 taps + bias = 80, padded -> conflict-free

template <int BLOCK_N>
struct DwPwSmem {
  static constexpr int kWinBytes = 25 * 1024;            // >= 2 * 10 * 10 * 128 (= 25600) and >= 18 * 10 * 128
  static constexpr int kABytes = 16384;
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kFixed = 1024 + 2 * kWinBytes + 2 * kABytes + 2 * kBBytes + 2 * 16384 + 256;
  static int total(int Cin) { return kFixed + (Cin / 8) * kDwVecStride * 4; }
};

template <int BLOCK_N>
__global__ void __launch_bounds__(384, 1)
dwpw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const DwPwParams p) {
  using SM = DwPwSmem<BLOCK_N>;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sWin = smem_base;
  const uint32_t sA = sWin + 2 * SM::kWinBytes;
  const uint32_t sB = sA + 2 * SM::kABytes;
  const uint32_t sEpi = sB + 2 * SM::kBBytes;
  const uint32_t sBar = sEpi + 2 * 16384;
  float* sDw = reinterpret_cast<float*>(smem_gen + (sBar - smem_base) + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (sBar - smem_base) + 192);
  auto wfull = [&](int s) { return sBar + 8u * s; };
  auto wempty = [&](int s) { return sBar + 16u + 8u * s; };
  auto afull = [&](int s) { return sBar + 32u + 8u * s; };
  auto aempty = [&](int s) { return sBar + 48u + 8u * s; };
  auto bfull = [&](int s) { return sBar + 64u + 8u * s; };
  auto bempty = [&](int s) { return sBar + 80u + 8u * s; };
  auto tfull = [&](int s) { return sBar + 96u + 8u * s; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < 2; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 8);   // one arrive per worker warp
      mbar_init(afull(s), 8);
      mbar_init(aempty(s), 1);
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
      mbar_init(tfull(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  pdl_launch_dependents();
  // depthwise weights + bias of every input channel (constants: may be read before the predecessor kernel is done): [Cin/8][10][8] fp32 at a conflict-free pitch
  for (int i = threadIdx.x; i < (p.Cin / 8) * 80; i += blockDim.x) {
    const int vec = i / 80, r = i - vec * 80, tap = r >> 3, e = r & 7;
    const int c = vec * 8 + e;
    sDw[vec * kDwVecStride + r] = tap < 9 ? p.dw_w[(size_t)tap * p.Cin + c] : p.dw_b[c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int tiles_per_group = p.tiles_x * p.tiles_y;
  auto decode = [&](int u, int& nb, int& grp, int& ty, int& tx) {
    nb = u % p.n_blocks;
    int mt = u / p.n_blocks;
    grp = mt / tiles_per_group;
    int r = mt - grp * tiles_per_group;
    ty = r / p.tiles_x;
    tx = r - ty * p.tiles_x;
  };
  const uint32_t win_bytes = (uint32_t)p.tile_n * p.wh * p.ww * 128;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int nb, grp, ty, tx;
        decode(u, nb, grp, ty, tx);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(wempty(s), ph ^ 1);
          mbar_expect_tx(wfull(s), win_bytes);
          tma_load_4d(sWin + s * SM::kWinBytes, &tmX, wfull(s), kb * 64, tx * 8 - 1, ty * p.tile_h - 1, grp * p.tile_n);
          mbar_wait(bempty(s), ph ^ 1);
          mbar_expect_tx(bfull(s), SM::kBBytes);
          tma_load_2d(sB + s * SM::kBBytes, &tmB, bfull(s), kb * 64, nb * BLOCK_N);
          if (++s == 2) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(1, 128, BLOCK_N);
      int s = 0;
      uint32_t ph = 0, i = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++i) {
        const uint32_t d_tmem = tmem_base + (i & 1) * BLOCK_N;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(afull(s), ph);
          mbar_wait(bfull(s), ph);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(sA + s * SM::kABytes);
          const uint64_t bdesc = umma_desc_sw128(sB + s * SM::kBBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<false>(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0);
          umma_commit(aempty(s));
          umma_commit(bempty(s));
          if (++s == 2) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(tfull(i & 1));
      }
    }
  } else if (warp >= 4) {
    const int t = threadIdx.x - 128;         // 0..255
    const int v = t & 7;                     // 16-byte channel vector within the 64-channel K-block
    const int x = (t >> 3) & 7;              // tile column
    const int rg = t >> 6;                   // row group: tile rows rg*4 .. rg*4+3
    const int img = (rg * 4) / p.tile_h;     // image within the tile
    const int y0 = (rg * 4) % p.tile_h;      // first output row within that image
    const int g = (warp - 4) >> 2;           // epilogue warpgroup
    const int ew = warp & 3;
    const int erow = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const bool leader = (ew == 0 && lane == 0);

    auto drain = [&](int u, uint32_t i) {
      int nb, grp, ty, tx;
      decode(u, nb, grp, ty, tx);
      mbar_wait(tfull(i & 1), (i >> 1) & 1);
      tc_fence_after();
      constexpr int NCHUNK = BLOCK_N / 64;
      for (int c = g; c < NCHUNK; c += 2) {
        const int n0 = nb * BLOCK_N + c * 64;
        if (leader) tma_store_wait_read<0>();
        named_bar_sync(1 + g, 128);
        const uint32_t st_row = sEpi + g * 16384 + erow * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_addr + (i & 1) * BLOCK_N + c * 64 + h * 32, r);
          tmem_ld_wait();
          float o[32];
          if (p.pw_b != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.pw_b + n0 + h * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = __ldg(b4 + q);
              o[4 * q] = __uint_as_float(r[4 * q]) + bb.x;
              o[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bb.y;
              o[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bb.z;
              o[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = o[8 * q + 2 * e], b = o[8 * q + 2 * e + 1];
              if (p.pw_act == 1) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
              if (p.pw_act == 2) { a = fminf(fmaxf(a, 0.f), 6.f); b = fminf(fmaxf(b, 0.f), 6.f); }
              __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
              w[e] = *reinterpret_cast<uint32_t*>(&b2);
            }
            const uint32_t a = st_row + (((uint32_t)(h * 4 + q) ^ (erow & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + g, 128);
        if (leader) {
          tma_store_4d(&tmD, sEpi + g * 16384, n0, tx * 8, ty * p.tile_h, grp * p.tile_n);
          tma_store_commit();
        }
      }
      tc_fence_before();  // TMEM reads of this accumulator are complete before the next A tiles are published
    };

    int s = 0;
    uint32_t ph = 0, i = 0;
    int prev_u = -1;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++i) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(wfull(s), ph);
        mbar_wait(aempty(s), ph ^ 1);
        const uint32_t win = sWin + s * SM::kWinBytes + (uint32_t)img * p.wh * p.ww * 128;
        const float* wv = sDw + (kb * 8 + v) * kDwVecStride;
        float acc[4][8];
        {
          const float4 b0 = *reinterpret_cast<const float4*>(wv + 72), b1 = *reinterpret_cast<const float4*>(wv + 76);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
            acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
          }
        }
#pragma unroll
        for (int sx = 0; sx < 3; ++sx) {
          float wk[3][8];
#pragma unroll
          for (int kr = 0; kr < 3; ++kr) {
            const float4 a = *reinterpret_cast<const float4*>(wv + (kr * 3 + sx) * 8);
            const float4 b = *reinterpret_cast<const float4*>(wv + (kr * 3 + sx) * 8 + 4);
            wk[kr][0] = a.x; wk[kr][1] = a.y; wk[kr][2] = a.z; wk[kr][3] = a.w;
            wk[kr][4] = b.x; wk[kr][5] = b.y; wk[kr][6] = b.z; wk[kr][7] = b.w;
          }
#pragma unroll
          for (int iy = 0; iy < 6; ++iy) {
            uint4 raw;
            const uint32_t a = win + (uint32_t)((y0 + iy) * p.ww + x + sx) * 128 + v * 16;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(a));
            const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
            float xv[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              xv[2 * e] = __uint_as_float(w4[e] << 16);
              xv[2 * e + 1] = __uint_as_float(w4[e] & 0xFFFF0000u);
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const int kr = iy - o;
              if (kr >= 0 && kr < 3) {
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[o][e] = fmaf(xv[e], wk[kr][e], acc[o][e]);
              }
            }
          }
        }
        // depthwise output -> A operand (row m = tile pixel, 16-byte chunk v, SWIZZLE_128B)
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int m = (rg * 4 + o) * 8 + x;
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a = acc[o][2 * e], b = acc[o][2 * e + 1];
            if (p.dw_act == 1) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            if (p.dw_act == 2) { a = fminf(fmaxf(a, 0.f), 6.f); b = fminf(fmaxf(b, 0.f), 6.f); }
            __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
            w[e] = *reinterpret_cast<uint32_t*>(&b2);
          }
          const uint32_t a = sA + s * SM::kABytes + m * 128 + (((uint32_t)v ^ (m & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
        }
        fence_proxy_async_smem();   // generic-proxy writes of A visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(afull(s));
          mbar_arrive(wempty(s));
        }
        if (++s == 2) {
          s = 0;
          ph ^= 1;
        }
      }
      if (prev_u >= 0) drain(prev_u, i - 1);
      prev_u = u;
    }
    if (prev_u >= 0) drain(prev_u, i - 1);
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace hfr
