// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   acc[M,N] = A[M,K] * B[N,K]^T   (both operands K-major)
//
//   warp 0           TMA producer: A/B k-blocks (128-byte rows, SWIZZLE_128B) into a STAGES-deep smem ring
//   warp 1           MMA issuer:   tcgen05.mma, M=128 x N=BLOCK_N per CTA, fp32 accumulators in TMEM (2 stages)
//   warp 2           TMEM allocator
//   warps 4..11      epilogue:     two warpgroups ping-pong on tiles (accumulator stage g = tile & 1);
//                                  tcgen05.ld -> registers -> fused epilogue
//   The producer and issuer warps run their loops whole-warp in uniform control flow; one elected lane issues
//   (elect_one() in ptx.cuh explains why that matters for the issue rate).
//
// CTAS = 2 (CTA pair, cta_group::2): the tile is 256 rows x BLOCK_N over two SMs of a cluster.  Each CTA stages its own
// 128 rows of A and BLOCK_N/2 rows of B; both CTAs' TMA loads complete on the LEADER's (rank 0) full barriers, only the
// leader issues MMAs, its commits are multicast to both CTAs' empty / accumulator-full barriers, each CTA's epilogue
// drains its own TMEM half and arrives on the leader's accumulator-empty barrier.  Work units are strided over clusters.
//
// Epilogues
//   EPI_STORE  out = act(acc + bias[n] (+ residual[m,n]))  -> T, staged in swizzled smem, written with TMA stores.
//              Used by the 1x1 pointwise convolutions (A = NHWC activations viewed as [B*H*W, Cin]) and, with the
//              im2col producer, by the 3x3 / 7x7 convolutions.
//   EPI_KNN    score = gnorm[n] - 2*acc; running per-row top-2 over the unit's gallery range, written once per unit.
//   EPI_KNN4   the same with a running top-4 (candidates for k-NN, k <= 4).
//
// Reference semantics being replaced: TF Conv2D(1x1)+Add+ReLU6 nodes `conv_pw_N*` of the frozen graph
// (facerec_test.py:120 sess.run) and sklearn's ArgKmin euclidean reduction (facerec_test.py:272,284-285).
#pragma once
#include "ptx.cuh"

namespace hfr {

enum { EPI_STORE = 0, EPI_KNN = 1, EPI_KNN4 = 2 };  // EPI_KNN4: running top-4 per row (k-NN with k <= 4)
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_RELU6 = 2 };
enum { AMODE_2D = 0, AMODE_IM2COL = 1, AMODE_STEM16 = 2 };

struct GemmParams {
  int M, N, K;            // K in elements
  int num_m_blocks;       // ceil(M/128); CTA pairs: ceil(M/256)
  int num_n_blocks;       // ceil(N/BLOCK_N)
  int n_blocks_per_unit;  // STORE: 1.  KNN: gallery n-blocks handled by one work unit
  int num_units;          // num_m_blocks * ceil(num_n_blocks / n_blocks_per_unit)
  // EPI_STORE
  const float* bias;      // [N] or nullptr
  const void* residual;   // [M,N] of T or nullptr
  int act;
  int round_tf32;         // round stored fp32 values to tf32 (they feed the next tf32 GEMM)
  // EPI_KNN
  const float* gnorm;     // [N] squared norms of the gallery rows
  float* part_score;      // [M][splits][2 warpgroups][2]   (EPI_KNN4: [...][4])
  int* part_idx;          // [M][splits][2 warpgroups][2]
  int splits;
  // AMODE_IM2COL (implicit GEMM over an NHWC tensor): k-block kb -> (tap, channel block)
  int conv_kw, conv_cblocks;   // taps along W, channel blocks (of BK elements) per tap
  int conv_wo, conv_ho;        // output width/height: m -> (n, ho, wo)
  int conv_stride, conv_pad_w, conv_pad_h, conv_dil;
  // AMODE_STEM16 (stem convolution after space-to-depth: 16 channels = 32 bytes per pixel, one MMA K-step per tap):
  // k-block = 4 taps; taps are (a, b) over a conv_kh x conv_kw window, stride 1, no padding.  B = [taps][N][16].
  int conv_kh, conv_taps;
  // K-concatenated A operand (AMODE_2D / AMODE_IM2COL): the first kb_split k-blocks come from a second, plain 2-D matrix
  // A0[M, kb_split * BK] (tensor map tmA0), the rest from the main operand; B = [N, K0 + K] holds both weight matrices side
  // by side.  ResNet's first block of a stage:  ReLU(increase(x_mid) + projection(x_in))  is ONE accumulation, the
  // 'increase' output is never written or re-read.
  int kb_split;
};

template <typename T>
struct GemmTraits;
template <>
struct GemmTraits<__nv_bfloat16> {
  static constexpr bool kTF32 = false;
  static constexpr uint32_t kFmt = 1;
  static constexpr int BK = 64;  // elements per 128-byte row
};
template <>
struct GemmTraits<float> {
  static constexpr bool kTF32 = true;
  static constexpr uint32_t kFmt = 2;
  static constexpr int BK = 32;
};

// CTAS = 2: CTA pair (cta_group::2) - the tile is 256 rows x BLOCK_N, each CTA stages its own 128 rows of A and
// BLOCK_N / 2 rows of B per k-block, so a k-block costs each SM 2/3 (BLOCK_N = 256) of the L2->SM bytes of the 1-CTA tile.
template <int BLOCK_N, int EPI, int CTAS = 1>
struct GemmSmem {
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BLOCK_N / CTAS * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = (EPI == EPI_STORE) ? 4 * 128 * 128 : 2 * BLOCK_N * 4;  // 2 warpgroups x 2 buffers
  static constexpr int kBudget = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - kEpiBytes;
  static constexpr int kStagesMax = kBudget / kStageBytes;
  static constexpr int kStages = kStagesMax > 6 ? 6 : kStagesMax;
  static constexpr int kTotal = 1024 + kStages * kStageBytes + kEpiBytes + 256;
};

template <typename T, int BLOCK_N, int EPI, int AMODE, int CTAS = 1>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmA0, const GemmParams p) {
  using TR = GemmTraits<T>;
  using SM = GemmSmem<BLOCK_N, EPI, CTAS>;
  constexpr int STAGES = SM::kStages;
  constexpr int BK = TR::BK;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // 128 / 256 / 512: power of two >= 32
  static_assert(BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * SM::kABytes;
  const uint32_t sEpi = smem_base + STAGES * SM::kStageBytes;
  uint8_t* sEpi_gen = smem_gen + STAGES * SM::kStageBytes;
  const uint32_t sBar = sEpi + SM::kEpiBytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sEpi_gen + SM::kEpiBytes + 192);
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 64u + 8u * s; };
  auto tfull_bar = [&](int s) { return sBar + 128u + 8u * s; };
  auto tempty_bar = [&](int s) { return sBar + 144u + 8u * s; };
  auto rfull_bar = [&](int s) { return sBar + 160u + 8u * s; };  // residual tile landed in staging buffer s (0..3)

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 is the leader (issues the MMAs, owns the full / tempty barriers both CTAs signal).  Work units are
  // strided over clusters; the CTA of rank r owns m block  CTAS * (unit's m block) + r.
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const int unit0 = (int)blockIdx.x / CTAS, unit_stride = (int)gridDim.x / CTAS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI == EPI_STORE) tma_prefetch_desc(&tmD);
    if (EPI == EPI_STORE && p.residual != nullptr) tma_prefetch_desc(&tmR);
    if (p.kb_split > 0) tma_prefetch_desc(&tmA0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * CTAS);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int s = 0; s < 4; ++s) mbar_init(rfull_bar(s), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS, CTAS>(smem_u32(tmem_slot));
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // the next kernel may run its prologue now ...
  pdl_wait();               // ... and this one touches global memory only after its predecessor has completed

  const int num_kb = (AMODE == AMODE_IM2COL)  ? p.conv_kw * p.conv_kw * p.conv_cblocks + p.kb_split
                     : (AMODE == AMODE_STEM16) ? (p.conv_taps + 3) / 4
                                               : (p.K + BK - 1) / BK;   // AMODE_2D: p.K is the total (K0 + K)
  const int units_n = (p.num_n_blocks + p.n_blocks_per_unit - 1) / p.n_blocks_per_unit;

  // unit -> (m block, first n block, n block count).  STORE: n fastest (neighbouring CTAs share the A tile in L2).
  // KNN: m fastest (all CTAs sweep the same gallery range while it is L2-resident).
  auto decode = [&](int u, int& mb, int& nb0, int& nbn) {
    if (EPI == EPI_STORE) {
      mb = (u / units_n) * CTAS + (int)cta_rank;
      nb0 = u % units_n;
      nbn = 1;
    } else {
      mb = (u % p.num_m_blocks) * CTAS + (int)cta_rank;
      int sp = u / p.num_m_blocks;
      nb0 = sp * p.n_blocks_per_unit;
      nbn = min(p.n_blocks_per_unit, p.num_n_blocks - nb0);
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp loops, one lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      // pair: completion bytes of both CTAs' loads land on the leader's barrier
      const uint32_t full_bar0 = (CTAS == 2) ? mapa_shared(full_bar(0), 0) : full_bar(0);
      for (int u = unit0; u < p.num_units; u += unit_stride) {
        int mb, nb0, nbn;
        decode(u, mb, nb0, nbn);
        int im_n = 0, im_h = 0, im_w = 0;
        if (AMODE != AMODE_2D) {
          int m0 = mb * 128;
          im_w = m0 % p.conv_wo;
          int t = m0 / p.conv_wo;
          im_h = t % p.conv_ho;
          im_n = t / p.conv_ho;
        }
        for (int nb = nb0; nb < nb0 + nbn; ++nb) {
          int cb = 0, tap_r = 0, tap_s = 0;  // im2col: channel block and filter tap of this k-block
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              const uint32_t fb = full_bar0 + 8u * stage;
              if (cta_rank == 0) mbar_expect_tx(full_bar(stage), SM::kStageBytes * CTAS);
              const bool prefix = (AMODE != AMODE_STEM16) && kb < p.kb_split;   // k-block of the concatenated A0 operand
              [[maybe_unused]] const int kb2 = kb - p.kb_split;
              if constexpr (CTAS == 2) {
                static_assert(CTAS == 1 || AMODE != AMODE_STEM16, "pair mode: 2-D and im2col A operands only");
                if (prefix)
                  tma_load_2d_pair(sA + stage * SM::kABytes, &tmA0, fb, kb * BK, mb * 128);
                else if (AMODE == AMODE_IM2COL)
                  tma_load_im2col_4d_pair(sA + stage * SM::kABytes, &tmA, fb, cb * BK,
                                          im_w * p.conv_stride - p.conv_pad_w, im_h * p.conv_stride - p.conv_pad_h, im_n,
                                          (uint16_t)(tap_s * p.conv_dil), (uint16_t)(tap_r * p.conv_dil));
                else
                  tma_load_2d_pair(sA + stage * SM::kABytes, &tmA, fb, kb2 * BK, mb * 128);
                // this CTA's half of the tile's B rows
                tma_load_2d_pair(sB + stage * SM::kBBytes, &tmB, fb, kb * BK,
                                 nb * BLOCK_N + (int)cta_rank * (BLOCK_N / 2));
              } else if (prefix) {
                tma_load_2d(sA + stage * SM::kABytes, &tmA0, full_bar(stage), kb * BK, mb * 128);
              } else if (AMODE == AMODE_IM2COL) {
                tma_load_im2col_4d(sA + stage * SM::kABytes, &tmA, full_bar(stage), cb * BK,
                                   im_w * p.conv_stride - p.conv_pad_w, im_h * p.conv_stride - p.conv_pad_h, im_n,
                                   (uint16_t)(tap_s * p.conv_dil), (uint16_t)(tap_r * p.conv_dil));
              } else if (AMODE == AMODE_STEM16) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  int tap = kb * 4 + t;
                  if (tap >= p.conv_taps) tap = 0;  // padding tap: its weights are zero (OOB in B's tap dimension)
                  const int a = tap / p.conv_kw, b = tap - a * p.conv_kw;
                  tma_load_im2col_4d(sA + stage * SM::kABytes + t * 4096, &tmA, full_bar(stage), 0, im_w, im_h, im_n,
                                     (uint16_t)b, (uint16_t)a);
                }
              } else {
                tma_load_2d(sA + stage * SM::kABytes, &tmA, full_bar(stage), kb2 * BK, mb * 128);
              }
              if (CTAS == 1 && AMODE == AMODE_STEM16)
                tma_load_3d(sB + stage * SM::kBBytes, &tmB, full_bar(stage), 0, nb * BLOCK_N, kb * 4);
              else if (CTAS == 1)
                tma_load_2d(sB + stage * SM::kBBytes, &tmB, full_bar(stage), kb * BK, nb * BLOCK_N);
            }
            __syncwarp();
            if (AMODE == AMODE_IM2COL && kb >= p.kb_split) {
              if (++cb == p.conv_cblocks) {
                cb = 0;
                if (++tap_s == p.conv_kw) {
                  tap_s = 0;
                  ++tap_r;
                }
              }
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      if constexpr (CTAS == 2) {
        // tail: every slot-free arrival the leader's commits multicast into this CTA has landed before it may exit
        for (int i = 0; i < STAGES; ++i) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    if (cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc(TR::kFmt, 128 * CTAS, BLOCK_N);
      const uint64_t a_desc0 = umma_desc_sw128(sA), b_desc0 = umma_desc_sw128(sB);
      const uint64_t a_desc0_sw32 = umma_desc_sw32(sA), b_desc0_sw32 = umma_desc_sw32(sB);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile = 0;
      for (int u = unit0; u < p.num_units; u += unit_stride) {
        int mb, nb0, nbn;
        decode(u, mb, nb0, nbn);
        for (int nb = nb0; nb < nb0 + nbn; ++nb, ++tile) {
          const uint32_t as = tile & 1, aphase = (tile >> 1) & 1;
          mbar_wait(tempty_bar(as), aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            // the descriptor's address field (16-byte units) is advanced by adds
            const uint32_t a_off = (uint32_t)(stage * SM::kABytes) >> 4, b_off = (uint32_t)(stage * SM::kBBytes) >> 4;
            if (elect_one()) {
              if constexpr (AMODE == AMODE_STEM16) {
#pragma unroll
                for (int k = 0; k < 4; ++k)  // one tap (16 channels = 32 bytes = one K-step) per MMA
                  umma<TR::kTF32>(d_tmem, a_desc0_sw32 + a_off + (uint32_t)(k * 4096 >> 4),
                                  b_desc0_sw32 + b_off + (uint32_t)(k * BLOCK_N * 32 >> 4), idesc, (kb | k) != 0);
              } else if constexpr (CTAS == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_pair<TR::kTF32>(d_tmem, a_desc0 + a_off + 2u * k, b_desc0 + b_off + 2u * k, idesc, (kb | k) != 0);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 4 x 32-byte K slices per 128-byte swizzle row
                  umma<TR::kTF32>(d_tmem, a_desc0 + a_off + 2u * k, b_desc0 + b_off + 2u * k, idesc, (kb | k) != 0);
              }
              if constexpr (CTAS == 2) {
                umma_commit_pair(empty_bar(stage));  // frees the slot in both CTAs
                if (kb == num_kb - 1) umma_commit_pair(tfull_bar(as));
              } else {
                umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
                if (kb == num_kb - 1) umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
              }
            }
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: two warpgroups, ping-pong on tiles
    // Warpgroup g (warps 4+4g .. 7+4g) drains accumulator stage g, i.e. every tile with (tile & 1) == g, while the
    // other warpgroup drains the previous/next tile.  Warp w may touch TMEM lanes 32*(w%4) .. +31.
    const int g = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int row = ew * 32 + lane;    // row within the 128-row tile
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t bar_id = 1 + g;     // named barrier of this warpgroup (128 threads)
    const bool leader = (ew == 0 && lane == 0);
    const uint32_t as = g;
    uint32_t tile = 0, my_tiles = 0;
    if constexpr (EPI == EPI_STORE) {
      constexpr int CH_ELEMS = 128 / (int)sizeof(T);  // columns per 128-byte staging chunk (32 fp32 / 64 bf16)
      constexpr int NCHUNK = BLOCK_N / CH_ELEMS;
      uint32_t chunk_ctr = 0;
      const bool use_res = (p.residual != nullptr);
      // Residual tiles travel by TMA into the staging buffer the output chunk will be written to (same swizzled
      // layout, updated in place), prefetched one chunk ahead by the warpgroup leader.
      const int units_n2 = (p.num_n_blocks + p.n_blocks_per_unit - 1) / p.n_blocks_per_unit;
      int pf_t = g, pf_c = 0;           // leader only: next (local tile, chunk) whose residual has not been requested
      uint32_t pf_ctr = 0;
      auto prefetch_next = [&]() {      // leader only
        const int u2 = unit0 + pf_t * unit_stride;
        if (u2 >= p.num_units) return;
        const int mb2 = (u2 / units_n2) * CTAS + (int)cta_rank, nb2 = u2 % units_n2;
        const int n2 = nb2 * BLOCK_N + pf_c * CH_ELEMS;
        const uint32_t b2 = pf_ctr & 1;
        mbar_expect_tx(rfull_bar(g * 2 + b2), 16384);
        tma_load_2d(sEpi + (g * 2 + b2) * 16384, &tmR, rfull_bar(g * 2 + b2), n2, mb2 * 128);
        ++pf_ctr;
        if (++pf_c == NCHUNK || nb2 * BLOCK_N + pf_c * CH_ELEMS >= p.N) {
          pf_c = 0;
          pf_t += 2;
        }
      };
      if (use_res && leader) prefetch_next();
      for (int u = unit0; u < p.num_units; u += unit_stride, ++tile) {
        if ((tile & 1) != (uint32_t)g) continue;
        int mb, nb0, nbn;
        decode(u, mb, nb0, nbn);
        const uint32_t aphase = my_tiles & 1;
        ++my_tiles;
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        for (int c = 0; c < NCHUNK; ++c, ++chunk_ctr) {
          const int n0 = nb0 * BLOCK_N + c * CH_ELEMS;
          if (n0 >= p.N) break;  // uniform across the CTA
          const uint32_t buf = chunk_ctr & 1;
          const uint32_t st_row = sEpi + (g * 2 + buf) * 16384 + row * 128;
          // All TMEM loads of the chunk and the first bias vector are issued up front (asynchronous): their latency
          // overlaps the wait for the residual tile / the staging buffer below instead of following it.
          constexpr int NH = CH_ELEMS / 32;
          uint32_t racc[NH][32];
          float4 bb[8];
          auto load_bias = [&](int h) -> bool {
            const bool has = (p.bias != nullptr && n0 + h * 32 < p.N);
            if (has) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + h * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) bb[q] = __ldg(b4 + q);
            }
            return has;
          };
          bool has_bias = load_bias(0);
          // (plain 1x1 layers only: same-box A/B - those ran 2-5 % faster, the implicit-GEMM convolutions 2-3 % slower)
          constexpr bool EARLY = (AMODE == AMODE_2D);
          if constexpr (EARLY) {
#pragma unroll
            for (int h = 0; h < NH; ++h) tmem_ld_32x32(tmem_base + lane_addr + as * BLOCK_N + c * CH_ELEMS + h * 32, racc[h]);
          }
          uint4 rv[8];
          if (use_res) {
            if (leader) {
              tma_store_wait_read<0>();  // the other buffer's last store has finished reading it ...
              prefetch_next();           // ... so the next chunk's residual may land there
            }
            mbar_wait(rfull_bar(g * 2 + buf), (chunk_ctr >> 1) & 1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t a = st_row + (((uint32_t)q ^ (row & 7)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(rv[q].x), "=r"(rv[q].y), "=r"(rv[q].z), "=r"(rv[q].w)
                           : "r"(a));
            }
          } else {
            // the TMA store that last read this staging buffer must have finished reading it
            if (leader) tma_store_wait_read<1>();
            named_bar_sync(bar_id, 128);
          }
          if constexpr (EARLY) tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint32_t (&r)[32] = racc[h];
            if constexpr (!EARLY) {   // one 32-column read at a time, right before its use
              tmem_ld_32x32(tmem_base + lane_addr + as * BLOCK_N + c * CH_ELEMS + h * 32, r);
              tmem_ld_wait();
            }
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
            if (h + 1 < NH) has_bias = load_bias(h + 1);   // in flight during this half's residual add / pack / store
            if (use_res) {
              if constexpr (sizeof(T) == 4) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  v[4 * q] += __uint_as_float(rv[q].x);
                  v[4 * q + 1] += __uint_as_float(rv[q].y);
                  v[4 * q + 2] += __uint_as_float(rv[q].z);
                  v[4 * q + 3] += __uint_as_float(rv[q].w);
                }
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 t = rv[h * 4 + q];
                  const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    v[8 * q + 2 * e] += __uint_as_float(w4[e] << 16);
                    v[8 * q + 2 * e + 1] += __uint_as_float(w4[e] & 0xFFFF0000u);
                  }
                }
              }
            }
            if constexpr (sizeof(T) == 4) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float x = v[j];
                if (p.act == ACT_RELU) x = fmaxf(x, 0.f);
                if (p.act == ACT_RELU6) x = fminf(fmaxf(x, 0.f), 6.f);
                v[j] = x;
              }
              if (p.round_tf32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = round_tf32(v[j]);
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) {  // 8 x 16-byte units, XOR-swizzled like SWIZZLE_128B
                const uint32_t a = st_row + (((uint32_t)q ^ (row & 7)) << 4);
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
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                             "r"(w[3]));
              }
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (leader) {
            tma_store_2d(&tmD, sEpi + (g * 2 + buf) * 16384, n0, mb * 128);
            tma_store_commit();
          }
        }
        // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));  // the leader CTA's barrier
          else mbar_arrive(tempty_bar(as));
        }
      }
      if (leader) tma_store_wait_all();
    } else {
      // ---------------------------------------------------------------- EPI_KNN
      float* gn = reinterpret_cast<float*>(sEpi_gen) + g * BLOCK_N;  // this warpgroup's gallery-norm slice
      for (int u = unit0; u < p.num_units; u += unit_stride) {
        int mb, nb0, nbn;
        decode(u, mb, nb0, nbn);
        constexpr int NC = (EPI == EPI_KNN4) ? 4 : 2;  // candidates kept per (row, gallery split, warpgroup)
        float b1 = INFINITY, b2 = INFINITY;
        int i1 = -1, i2 = -1;
        [[maybe_unused]] float b3 = INFINITY, b4 = INFINITY;
        [[maybe_unused]] int i3 = -1, i4 = -1;
        for (int nb = nb0; nb < nb0 + nbn; ++nb, ++tile) {
          if ((tile & 1) != (uint32_t)g) continue;
          const uint32_t aphase = my_tiles & 1;
          ++my_tiles;
          named_bar_sync(bar_id, 128);  // everyone is done reading the previous tile's norms
          for (int j = row; j < BLOCK_N; j += 128) {
            const int n = nb * BLOCK_N + j;
            gn[j] = (n < p.N) ? __ldg(p.gnorm + n) : INFINITY;
          }
          named_bar_sync(bar_id, 128);
          mbar_wait(tfull_bar(as), aphase);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < BLOCK_N / 32; ++c) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld_32x32(tmem_base + lane_addr + as * BLOCK_N + c * 32, r);
            tmem_ld_wait();
            const int nbase = nb * BLOCK_N + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = fmaf(-2.f, __uint_as_float(r[j]), gn[c * 32 + j]);
              if constexpr (NC == 4) {
                if (s < b4) {  // sorted insert, strict '<': on equal scores the row seen first (lower index) stays ahead
                  if (s < b2) {
                    b4 = b3; i4 = i3;
                    b3 = b2; i3 = i2;
                    if (s < b1) {
                      b2 = b1; i2 = i1;
                      b1 = s; i1 = nbase + j;
                    } else {
                      b2 = s; i2 = nbase + j;
                    }
                  } else if (s < b3) {
                    b4 = b3; i4 = i3;
                    b3 = s; i3 = nbase + j;
                  } else {
                    b4 = s; i4 = nbase + j;
                  }
                }
              } else if (s < b2) {
                if (s < b1) {
                  b2 = b1;
                  i2 = i1;
                  b1 = s;
                  i1 = nbase + j;
                } else {
                  b2 = s;
                  i2 = nbase + j;
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0));
            else mbar_arrive(tempty_bar(as));
          }
        }
        const int m = mb * 128 + row;
        if (m < p.M) {
          // one (top-2) record per (row, gallery split, warpgroup)
          const int sp = nb0 / p.n_blocks_per_unit;
          const size_t o = (((size_t)m * p.splits + sp) * 2 + g) * NC;
          p.part_score[o] = b1;
          p.part_score[o + 1] = b2;
          p.part_idx[o] = i1;
          p.part_idx[o + 1] = i2;
          if constexpr (NC == 4) {
            p.part_score[o + 2] = b3;
            p.part_score[o + 3] = b4;
            p.part_idx[o + 2] = i3;
            p.part_idx[o + 3] = i4;
          }
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all();  // neither CTA's shared memory / TMEM goes away while the peer uses it
  else __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS, CTAS>(tmem_base);
}

}  // namespace hfr
