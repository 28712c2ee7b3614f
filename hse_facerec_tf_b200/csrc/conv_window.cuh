// "Window" implicit-GEMM convolution for sm_100a: KHxKW stride-1 convolution over an NHWC bf16 tensor with Cout <= 64,
// where every activation byte crosses L2->SM once per tile and the weights stay resident in shared memory.
//
//   * M tile = 16 (rows) x 8 (columns) output pixels of one image.  Cin/8 TMA loads (one per 8-channel plane) bring the
//     (16+KH-1) x (8+KW-1) input window into smem as channel planes [Cin/8][wh][ww][8 ch = 16 B]; out-of-image pixels
//     are zero-filled by the TMA unit = the convolution's zero padding.
//   * The tensor core reads that window IN PLACE: with the no-swizzle K-major canonical layout a UMMA operand is
//     "8-row groups of 16-byte rows", group stride SBO, k-chunk stride LBO.  8 horizontally adjacent pixels of a plane
//     are exactly such a group (16-B pitch), the 16 tile rows are 16 groups at SBO = ww*16 B, and the second 8 channels
//     of a K=16 step are the next plane at LBO = wh*ww*16 B.  Tap (r,s) is just a different start address.  No im2col
//     copy exists anywhere - neither in HBM nor in smem.
//   * Weights [taps][Cin/8][64 rows][16 B] are loaded once per (persistent) CTA.
//   * Epilogue: two warpgroups ping-pong on tiles; bias + ReLU/ReLU6 -> bf16 -> swizzled staging -> 4-D TMA store
//     (partial tiles are clipped by the TMA unit).
//
// Used for the stem (after space-to-depth, see stem_s2d_kernel: 7x7/2 -> 4x4/1 over 16 channels) and for 3x3 convolutions
// with 64 output channels (ResNet stage 2), which the generic im2col-TMA path leaves L2-bandwidth-bound.
#pragma once
#include "gemm_tc.cuh"  // ACT_* codes
#include "ptx.cuh"

namespace hfr {

struct WinParams {
  int tiles_x, tiles_y, num_tiles;  // per image: ceil(Wo/8) x ceil(Ho/16); num_tiles = B * tiles_x * tiles_y
  int taps_h, taps_w;
  int planes;                       // Cin / 8
  int ww, wh;                       // window size in pixels
  int pad_l, pad_t;
  int N;                            // real output channels (<= 64)
  const float* bias;
  int act;
  int w_chunks;                     // taps_h * taps_w * planes  (16-byte k-chunks of the weight matrix)
  int plane_major;                  // input is [B][planes][H][W][8]: the whole window is ONE TMA box with long rows
  int plane_pitch;                  // bytes between consecutive channel planes of a window in smem
  int shifted;                      // plane-major only: taps_w copies of the window, copy s shifted by s pixels, each
                                    // exactly 8 pixels wide -> every 8-row operand group is one aligned 128-byte line
  int copy_pitch;                   // bytes between consecutive shifted copies
  int round_tf32;                   // fp32 output only: round to tf32 (the consumer is a tf32 tensor-core layer)
  int stages;                       // window ring depth (kWinStagesMin .. kWinStagesMax)
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// K-major, no swizzle: 8-row groups of 16-byte rows; lbo = byte distance of the second 16-byte k-chunk, sbo = byte
// distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// Window ring depth: a slot is recycled only after the MMAs that read it have completed, so the depth must cover the
// load latency (ncu r2: with 3 slots the stem sat at 0.94 us per tile - (load latency + MMA time) / 3 - against 0.41 us
// of MMA work).  The launcher takes as many slots as fit next to the resident weights, between these bounds.
constexpr int kWinStagesMin = 3, kWinStagesMax = 6;

// dynamic smem: [1 KB align][weights w_bytes][p.stages x window win_bytes (1 KB-rounded)][4 x 16 KB staging][barriers]
// TH x TW taps with KS K=16 steps each known at compile time unroll the MMA issue completely (TH = 0: runtime loops).
// TOut = float (experimental, the tf32-mode stem): fp32 output in two 32-column chunks per tile.  PASSES = 2: the weight
// matrix holds a second copy of every tap with the bf16 residual of the weights (w = hi + lo, 16 mantissa bits), issued
// as a second sweep of MMAs over the same window into the same accumulator.
template <int TH, int TW, int KS, typename TOut = __nv_bfloat16, int PASSES = 1>
__global__ void __launch_bounds__(384, 1)
conv_window_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ CUtensorMap tmD, const WinParams p, const int w_bytes, const int win_stride) {
  constexpr int BLOCK_N = 64;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sW = smem_base;
  const uint32_t sWin = sW + w_bytes;                      // w_bytes is a multiple of 1024
  const int kWinStages = p.stages;
  const uint32_t sEpi = sWin + kWinStages * win_stride;    // win_stride is a multiple of 1024
  const uint32_t sBar = sEpi + 4 * 16384;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (sBar - smem_base) + 192);
  auto full_bar = [&](int s) { return sBar + 8u * s; };            // up to kWinStagesMax = 6 slots: 48 bytes each kind
  auto empty_bar = [&](int s) { return sBar + 48u + 8u * s; };
  auto tfull_bar = [&](int s) { return sBar + 96u + 8u * s; };
  auto tempty_bar = [&](int s) { return sBar + 112u + 8u * s; };
  const uint32_t wfull_bar = sBar + 128u;

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < kWinStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    mbar_init(wfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const uint32_t win_bytes = p.shifted ? (uint32_t)p.taps_w * p.planes * p.wh * 128   // bytes the window loads deliver
                                       : (uint32_t)p.planes * p.wh * p.ww * 16;
  const uint32_t plane_bytes = (uint32_t)p.plane_pitch;

  // Producer and MMA issuer: the whole warp runs the loops (uniform control flow), one elected lane issues - see
  // elect_one() in ptx.cuh.
  if (warp == 0) {
    if (elect_one()) {
      // weights: resident for the lifetime of the CTA (3-D map {8, 64 rows, chunks}, one K=16 step = 2 chunks per box)
      mbar_expect_tx(wfull_bar, (uint32_t)p.w_chunks * BLOCK_N * 16);
      for (int c0 = 0; c0 < p.w_chunks; c0 += 2) tma_load_3d(sW + c0 * BLOCK_N * 16, &tmW, wfull_bar, 0, 0, c0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const int b = t / tiles_per_img;
      const int r = t - b * tiles_per_img;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      mbar_wait(empty_bar(stage), phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full_bar(stage), win_bytes);
        if (p.plane_major && p.shifted) {
          for (int sh = 0; sh < p.taps_w; ++sh)   // box {8 px * 8 ch, wh, planes, 1} starting sh pixels to the right
            tma_load_4d(sWin + stage * win_stride + sh * p.copy_pitch, &tmX, full_bar(stage),
                        (tx * 8 - p.pad_l + sh) * 8, ty * 16 - p.pad_t, 0, b);
        } else if (p.plane_major) {
          // [B][planes][H][W*8]: one box {ww*8, wh, planes, 1} - rows of ww*16 contiguous bytes
          tma_load_4d(sWin + stage * win_stride, &tmX, full_bar(stage), (tx * 8 - p.pad_l) * 8, ty * 16 - p.pad_t, 0, b);
        } else {
          for (int pl = 0; pl < p.planes; ++pl)  // NHWC: one 8-channel plane per load: smem [plane][wh][ww][16 B]
            tma_load_4d(sWin + stage * win_stride + pl * plane_bytes, &tmX, full_bar(stage), pl * 8, tx * 8 - p.pad_l,
                        ty * 16 - p.pad_t, b);
        }
      }
      __syncwarp();
      if (++stage == kWinStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(1, 128, BLOCK_N);
    const uint32_t sbo_a = p.shifted ? 128u : (uint32_t)p.ww * 16;
    const int ksteps = p.planes >> 1;  // K=16 (two 8-channel planes) per MMA
    // Descriptors advance by adds on the address field (16-byte units): an N=64 MMA occupies the tensor core for only
    // ~48 cycles (tools/microbench/mma_bench.cu), so the issue loop must be a handful of uniform-datapath instructions.
    const uint64_t a_desc0 = umma_desc_noswz(sWin, plane_bytes, sbo_a);
    const uint64_t b_desc0 = umma_desc_noswz(sW, BLOCK_N * 16, 128);
    const uint32_t a_row_step = p.shifted ? 8u : (uint32_t)p.ww;                 // one window row
    const uint32_t a_col_step = p.shifted ? (uint32_t)p.copy_pitch >> 4 : 1u;    // one window column
    const uint32_t a_k_step = (2u * plane_bytes) >> 4, b_k_step = (2u * BLOCK_N * 16) >> 4;  // two 8-channel planes
    mbar_wait(wfull_bar, 0);
    int stage = 0;
    uint32_t phase = 0, tile = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tile) {
      const uint32_t as = tile & 1, aphase = (tile >> 1) & 1;
      mbar_wait(tempty_bar(as), aphase ^ 1);
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      if (elect_one()) {
        uint64_t a_row = a_desc0 + (uint64_t)((uint32_t)(stage * win_stride) >> 4);
        uint64_t bd = b_desc0;
        uint32_t first = 0;
        if constexpr (TH > 0) {
          // independent descriptor offsets, no loop-carried chain through the (slow) uniform datapath; 32-bit arithmetic
          // on the descriptors' low words only (the offsets never reach the high word: start address + LBO fields)
          const uint32_t a_lo = (uint32_t)a_row, a_hi = (uint32_t)(a_row >> 32);
          const uint32_t b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
#pragma unroll
          for (int ps = 0; ps < PASSES; ++ps)
#pragma unroll
            for (int r = 0; r < TH; ++r)
#pragma unroll
              for (int s = 0; s < TW; ++s)
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                  const int i = ((ps * TH + r) * TW + s) * KS + j;
                  umma_bf16_lohi(d_tmem, a_lo + (r * a_row_step + s * a_col_step + j * a_k_step), a_hi,
                                 b_lo + (uint32_t)(i * b_k_step), b_hi, idesc, i != 0);
                }
        } else
        for (int ps = 0; ps < PASSES; ++ps, a_row -= (uint64_t)p.taps_h * a_row_step)
        for (int r = 0; r < p.taps_h; ++r, a_row += a_row_step) {
          uint64_t a_tap = a_row;
          for (int s = 0; s < p.taps_w; ++s, a_tap += a_col_step) {
            uint64_t ad = a_tap;
            for (int j = 0; j < ksteps; ++j, ad += a_k_step, bd += b_k_step) {
              umma<false>(d_tmem, ad, bd, idesc, first);
              first = 1;
            }
          }
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (++stage == kWinStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t bar_id = 1 + g;
    const bool leader = (ew == 0 && lane == 0);
    const uint32_t as = g;
    uint32_t tile = 0, my_tiles = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tile) {
      if ((tile & 1) != (uint32_t)g) continue;
      const int b = t / tiles_per_img;
      const int rr = t - b * tiles_per_img;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const uint32_t aphase = my_tiles & 1;
      const uint32_t buf = my_tiles & 1;
      ++my_tiles;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      // bf16: a tile is one 16 KB staging buffer, double-buffered across this warpgroup's tiles.  fp32: a tile is two
      // 32-column chunks = both buffers, so every earlier store must have finished reading them.
      if (leader) {
        if constexpr (sizeof(TOut) == 4) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
      }
      named_bar_sync(bar_id, 128);
      const uint32_t st_row = sEpi + (g * 2 + (sizeof(TOut) == 4 ? 0 : buf)) * 16384 + row * 128;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        const bool has_bias = (p.bias != nullptr && h * 32 < p.N);
        float4 bb[8];     // bias loads ahead of the TMEM read: the two latencies overlap
        if (has_bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + h * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) bb[q] = __ldg(b4 + q);
        }
        tmem_ld_32x32(tmem_base + lane_addr + as * BLOCK_N + h * 32, r);
        tmem_ld_wait();
        float v[32];
        if (has_bias) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[4 * q] = __uint_as_float(r[4 * q]) + bb[q].x;
            v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bb[q].y;
            v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bb[q].z;
            v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bb[q].w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        if constexpr (sizeof(TOut) == 4) {
          // chunk h = output channels 32h .. 32h+31 as 128-byte fp32 rows in staging buffer h of this warpgroup
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = v[j];
            if (p.act == ACT_RELU) x = fmaxf(x, 0.f);
            if (p.act == ACT_RELU6) x = fminf(fmaxf(x, 0.f), 6.f);
            v[j] = p.round_tf32 ? round_tf32(x) : x;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t a = st_row + h * 16384 + (((uint32_t)q ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                         "f"(v[4 * q + 2]), "f"(v[4 * q + 3]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2_act(v[8 * q + 2 * e], v[8 * q + 2 * e + 1], p.act);
            const uint32_t a = st_row + (((uint32_t)(h * 4 + q) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 128);
      if (leader) {
        if constexpr (sizeof(TOut) == 4) {
          tma_store_4d(&tmD, sEpi + (g * 2 + 0) * 16384, 0, tx * 8, ty * 16, b);
          if (p.N > 32) tma_store_4d(&tmD, sEpi + (g * 2 + 1) * 16384, 32, tx * 8, ty * 16, b);
        } else {
          tma_store_4d(&tmD, sEpi + (g * 2 + buf) * 16384, 0, tx * 8, ty * 16, b);
        }
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace hfr
