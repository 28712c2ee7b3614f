// Two dependent 1x1 convolutions in ONE persistent launch:   Y = act0(A0 W0^T + b0 (+ R))  ;  Z = act1(Y W1^T + b1)
//
// ResNet bottleneck chains end one block with the "increase" GEMM (+ residual + ReLU) and start the next with the
// "reduce" GEMM over the tensor just written.  As two launches Y travels HBM -> SM twice (written, then re-read) and
// every launch pays its own ramp and tail.  Here both GEMMs' tiles are work units of one persistent kernel, ordered so
// that the consumer tile of row block m runs `lag` row blocks after the producer tiles of m:
//     unit order:  [Y(0,*)] [Y(1,*)] ... [Y(g,*), Z(g-lag,*)] ...            (* = all column blocks of that GEMM)
// Y is still written once (the next block needs it as its residual), but the consumer reads it while it is L2-resident
// (lag x 128 rows, a few MB) - the HBM read of Y disappears, and so does one launch.
//
// Correct ordering between CTAs: a producer tile is published - done[m] += 1, release at gpu scope - once its TMA stores
// have completed.  In steady state the epilogue warpgroup does that while it finishes its NEXT tile (the stores are long
// complete by then, the wait costs nothing; waiting right after a tile's own stores serialised the epilogue on the store
// latency and made the pair 1.6x slower than two launches).  A warpgroup never blocks with an unpublished tile, though:
// if its next accumulator is not ready it publishes first.  The TMA producer warp of a consumer tile acquires
// done[m] == column blocks of Y before it requests the rows.  Dependencies only point to earlier units, every CTA walks
// its units in order and nobody blocks while holding an unpublished tile, so the unit with the smallest index is always
// runnable: no deadlock for any grid size or lag.  (The first deferred version published only from inside the next tile
// and deadlocked in the sparse tail of the schedule, where that next tile can depend on the unpublished one.)
//
// Same warp roles, rings, TMEM double buffering and epilogue as gemm_tc_kernel (1-CTA tiles, 2-D operands).
#pragma once
#include "gemm_tc.cuh"

namespace hfr {

struct ChainProblem {
  int N, K;              // output columns, reduction length
  int num_n_blocks;      // ceil(N / BLOCK_N)
  const float* bias;     // [N] or nullptr
  const void* residual;  // [M, N] of T or nullptr
  int act, round_tf32;
};
struct ChainParams {
  ChainProblem pr[2];    // 0: producer GEMM (writes Y), 1: consumer GEMM (its A operand is Y)
  int M, num_m_blocks;
  int lag;               // row blocks per super-block: the consumer works one super-block behind the producer
  int num_units;         // (ceil(num_m_blocks / lag) + 1) * lag * (pr[0].num_n_blocks + pr[1].num_n_blocks)
  unsigned* done;        // [num_m_blocks] finished producer tiles per row block, zero when the launch starts
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all_but() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// counters of every chained pair of a forward pass, cleared by one (PDL-ordered) kernel at the start of the pass
__global__ void zero_u32_kernel(uint4* p, size_t n16) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}

template <typename T, int BLOCK_N>
__global__ void __launch_bounds__(384, 1)
gemm_chain_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                  const __grid_constant__ CUtensorMap tmD0, const __grid_constant__ CUtensorMap tmR0,
                  const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ CUtensorMap tmD1, const __grid_constant__ CUtensorMap tmR1,
                  const ChainParams p) {
  using TR = GemmTraits<T>;
  using SM = GemmSmem<BLOCK_N, EPI_STORE, 1>;
  constexpr int STAGES = SM::kStages;
  constexpr int BK = TR::BK;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
  static_assert(BLOCK_N == 64 || BLOCK_N == 128, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * SM::kABytes;
  const uint32_t sEpi = smem_base + STAGES * SM::kStageBytes;
  uint8_t* sEpi_gen = smem_gen + STAGES * SM::kStageBytes;
  const uint32_t sBar = sEpi + SM::kEpiBytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sEpi_gen + SM::kEpiBytes + 192);
  const uint32_t dep_ready = sBar + 200u;   // consumer tiles of this CTA whose rows are known to be published
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 64u + 8u * s; };
  auto tfull_bar = [&](int s) { return sBar + 128u + 8u * s; };
  auto tempty_bar = [&](int s) { return sBar + 144u + 8u * s; };
  auto rfull_bar = [&](int s) { return sBar + 160u + 8u * s; };

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmB0); tma_prefetch_desc(&tmD0);
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmD1);
    if (p.pr[0].residual != nullptr) tma_prefetch_desc(&tmR0);
    if (p.pr[1].residual != nullptr) tma_prefetch_desc(&tmR1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    for (int s = 0; s < 4; ++s) mbar_init(rfull_bar(s), 1);
    st_release_cta_shared(dep_ready, 0u);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS, 1>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int nb0 = p.pr[0].num_n_blocks, nb1 = p.pr[1].num_n_blocks;
  const int G = nb0 + nb1;
  // unit -> (problem, row block, column block); false: the unit is a hole of the schedule (before the consumer starts
  // or after the producer has finished)
  // Phased order: super-blocks of p.lag row blocks; super-block s = [producer tiles of its rows][consumer tiles of the
  // rows of super-block s - 1].  Within a phase all tiles have the same K, so the load ring keeps its look-ahead (with
  // the two kinds interleaved tile by tile a K = 256 consumer tile took 4 of the 5 slots and the HBM-latency-bound
  // producer tiles around it lost their prefetch depth).
  const int R = p.lag, SP = R * nb0, S = R * G;
  auto decode = [&](int u, int& q, int& mb, int& nb) -> bool {
    const int sb = u / S, r = u - sb * S;
    if (r < SP) {
      const int i = r / nb0;
      q = 0; mb = sb * R + i; nb = r - i * nb0;
    } else {
      const int r2 = r - SP, i = r2 / nb1;
      q = 1; mb = (sb - 1) * R + i; nb = r2 - i * nb1;
      if (sb == 0) return false;
    }
    return mb >= 0 && mb < p.num_m_blocks;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0, consumer_idx = 0;   // consumer tiles of this CTA handed to the loader so far
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int q, mb, nb;
      if (!decode(u, q, mb, nb)) continue;
      if (q == 1) {   // rows of Y: warp 3 has seen every producer tile of this row block published
        if (lane == 0) {
          while (ld_acquire_cta_shared(dep_ready) <= consumer_idx) __nanosleep(20);
          fence_proxy_async_all();   // the published rows were written by the async proxy (TMA stores); so are our reads
        }
        __syncwarp();
        ++consumer_idx;
      }
      const CUtensorMap* ta = q ? &tmA1 : &tmA0;
      const CUtensorMap* tb = q ? &tmB1 : &tmB0;
      const int num_kb = (p.pr[q].K + BK - 1) / BK;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), SM::kStageBytes);
          tma_load_2d(sA + stage * SM::kABytes, ta, full_bar(stage), kb * BK, mb * 128);
          tma_load_2d(sB + stage * SM::kBBytes, tb, full_bar(stage), kb * BK, nb * BLOCK_N);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ dependency scout
    // Walks this CTA's consumer tiles in order, well ahead of the loader, and polls their row block's counter with
    // gpu-scope acquire loads (an L2 round trip each - on the loader's own path they cost ~1 us per consumer tile);
    // the loader then only reads a shared-memory count.
    if (lane == 0) {
      uint32_t seen = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int q, mb, nb;
        if (!decode(u, q, mb, nb) || q == 0) continue;
        while (ld_acquire_gpu(p.done + mb) < (unsigned)nb0) __nanosleep(64);
        ++seen;
        st_release_cta_shared(dep_ready, seen);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc(TR::kFmt, 128, BLOCK_N);
    const uint64_t a_desc0 = umma_desc_sw128(sA), b_desc0 = umma_desc_sw128(sB);
    int stage = 0;
    uint32_t phase = 0, tile = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int q, mb, nb;
      if (!decode(u, q, mb, nb)) continue;
      const int num_kb = (p.pr[q].K + BK - 1) / BK;
      const uint32_t as = tile & 1, aphase = (tile >> 1) & 1;
      ++tile;
      mbar_wait(tempty_bar(as), aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_off = (uint32_t)(stage * SM::kABytes) >> 4, b_off = (uint32_t)(stage * SM::kBBytes) >> 4;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma<TR::kTF32>(d_tmem, a_desc0 + a_off + 2u * k, b_desc0 + b_off + 2u * k, idesc, (kb | k) != 0);
          umma_commit(empty_bar(stage));
          if (kb == num_kb - 1) umma_commit(tfull_bar(as));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: two warpgroups ping-pong on tiles
    const int g = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t bar_id = 1 + g;
    const bool leader = (ew == 0 && lane == 0);
    const uint32_t as = g;
    constexpr int CH_ELEMS = 128 / (int)sizeof(T);
    constexpr int NCHUNK = BLOCK_N / CH_ELEMS;
    uint32_t tile = 0, my_tiles = 0, chunk_ctr = 0;
    // Residual tiles travel by TMA into the staging buffer the output chunk is written to, requested one chunk ahead by
    // the warpgroup leader.  The leader walks this warpgroup's future (tile, chunk) sequence with its own cursor; the
    // chunks of problems without a residual advance the staging-buffer parity but request nothing.
    int pf_u = blockIdx.x, pf_c = 0;
    uint32_t pf_tile = 0, pf_ctr = 0;
    bool pf_valid = false;
    int pf_q = 0, pf_mb = 0, pf_nb = 0;
    auto pf_seek = [&]() {   // move the cursor to this warpgroup's next tile (leader only)
      pf_valid = false;
      while (pf_u < p.num_units) {
        if (decode(pf_u, pf_q, pf_mb, pf_nb)) {
          const bool mine = (pf_tile & 1) == (uint32_t)g;
          ++pf_tile;
          if (mine) {
            pf_valid = true;
            pf_c = 0;
            return;
          }
        }
        pf_u += gridDim.x;
      }
    };
    auto pf_advance = [&]() {   // request the cursor's chunk (if its problem has a residual), then step the cursor
      if (!pf_valid) return;
      const ChainProblem& P = p.pr[pf_q];
      const int n2 = pf_nb * BLOCK_N + pf_c * CH_ELEMS;
      if (P.residual != nullptr) {
        const uint32_t b2 = pf_ctr & 1;
        mbar_expect_tx(rfull_bar(g * 2 + b2), 16384);
        tma_load_2d(sEpi + (g * 2 + b2) * 16384, pf_q ? &tmR1 : &tmR0, rfull_bar(g * 2 + b2), n2, pf_mb * 128);
      }
      ++pf_ctr;
      if (++pf_c == NCHUNK || pf_nb * BLOCK_N + pf_c * CH_ELEMS >= P.N) {
        pf_u += gridDim.x;
        pf_seek();
      }
    };
    int pub_mb = -1;   // row block of this warpgroup's last producer tile whose completion has not been published yet
    uint32_t res_waits0 = 0, res_waits1 = 0;   // completed waits on this warpgroup's two residual barriers (phase parity)
    if (leader) {
      pf_seek();
      pf_advance();
    }
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
      int q, mb, nb;
      if (!decode(u, q, mb, nb)) continue;
      const bool mine = (tile & 1) == (uint32_t)g;
      ++tile;
      if (!mine) continue;
      const ChainProblem& P = p.pr[q];
      const bool use_res = (P.residual != nullptr);
      const uint32_t aphase = my_tiles & 1;
      ++my_tiles;
      if (leader && pub_mb >= 0 && !mbar_try_wait(tfull_bar(as), aphase)) {
        // This tile's accumulator is not ready, i.e. we are about to block - possibly on a consumer tile that (through
        // other CTAs) waits for the very tile we have not published yet (sparse tail of the schedule: the next valid
        // tile of this warpgroup can be many waves away).  Never block holding an unpublished tile: publish now; the
        // wait for the stores overlaps the wait for the MMAs.
        tma_store_wait_all();
        fence_proxy_async_all();
        red_release_gpu_add(p.done + pub_mb, 1u);
        pub_mb = -1;
      }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      for (int c = 0; c < NCHUNK; ++c, ++chunk_ctr) {
        const int n0 = nb * BLOCK_N + c * CH_ELEMS;
        if (n0 >= P.N) break;
        const uint32_t buf = chunk_ctr & 1;
        const uint32_t st_row = sEpi + (g * 2 + buf) * 16384 + row * 128;
        uint4 rv[8];
        // the other buffer's last store has finished reading it: the next chunk's residual may land there
        if (leader) {
          tma_store_wait_read<0>();
          pf_advance();
        }
        if (use_res) {
          if (buf == 0) {
            mbar_wait(rfull_bar(g * 2), res_waits0 & 1);
            ++res_waits0;
          } else {
            mbar_wait(rfull_bar(g * 2 + 1), res_waits1 & 1);
            ++res_waits1;
          }
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const uint32_t a = st_row + (((uint32_t)qq ^ (row & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rv[qq].x), "=r"(rv[qq].y), "=r"(rv[qq].z), "=r"(rv[qq].w)
                         : "r"(a));
          }
        } else {
          named_bar_sync(bar_id, 128);   // the leader's wait covers this buffer for the whole warpgroup
        }
#pragma unroll
        for (int h = 0; h < CH_ELEMS / 32; ++h) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + lane_addr + as * BLOCK_N + c * CH_ELEMS + h * 32, r);
          tmem_ld_wait();
          float v[32];
          if (P.bias != nullptr && n0 + h * 32 < P.N) {
            const float4* b4 = reinterpret_cast<const float4*>(P.bias + n0 + h * 32);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const float4 bb = __ldg(b4 + qq);
              v[4 * qq] = __uint_as_float(r[4 * qq]) + bb.x;
              v[4 * qq + 1] = __uint_as_float(r[4 * qq + 1]) + bb.y;
              v[4 * qq + 2] = __uint_as_float(r[4 * qq + 2]) + bb.z;
              v[4 * qq + 3] = __uint_as_float(r[4 * qq + 3]) + bb.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          if (use_res) {
            if constexpr (sizeof(T) == 4) {
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                v[4 * qq] += __uint_as_float(rv[qq].x);
                v[4 * qq + 1] += __uint_as_float(rv[qq].y);
                v[4 * qq + 2] += __uint_as_float(rv[qq].z);
                v[4 * qq + 3] += __uint_as_float(rv[qq].w);
              }
            } else {
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                const uint4 t = rv[h * 4 + qq];
                const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[8 * qq + 2 * e] += __uint_as_float(w4[e] << 16);
                  v[8 * qq + 2 * e + 1] += __uint_as_float(w4[e] & 0xFFFF0000u);
                }
              }
            }
          }
          if constexpr (sizeof(T) == 4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = v[j];
              if (P.act == ACT_RELU) x = fmaxf(x, 0.f);
              if (P.act == ACT_RELU6) x = fminf(fmaxf(x, 0.f), 6.f);
              v[j] = P.round_tf32 ? round_tf32(x) : x;
            }
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const uint32_t a = st_row + (((uint32_t)qq ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * qq]), "f"(v[4 * qq + 1]),
                           "f"(v[4 * qq + 2]), "f"(v[4 * qq + 3]));
            }
          } else {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2_act(v[8 * qq + 2 * e], v[8 * qq + 2 * e + 1], P.act);
              const uint32_t a = st_row + (((uint32_t)(h * 4 + qq) ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (leader) {
          tma_store_2d(q ? &tmD1 : &tmD0, sEpi + (g * 2 + buf) * 16384, n0, mb * 128);
          tma_store_commit();
          if (c == 0 && pub_mb >= 0) {
            // Deferred publication of this warpgroup's PREVIOUS producer tile: every bulk group but the one just
            // committed has completed (not merely been read) - that tile's stores were issued a whole tile ago, so this
            // wait is free; waiting right after a tile's own stores serialised the epilogue on the store latency.
            tma_store_wait_all_but<1>();
            fence_proxy_async_all();
            red_release_gpu_add(p.done + pub_mb, 1u);
            pub_mb = -1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (q == 0) pub_mb = mb;   // (leader's copy is the one that matters)
    }
    if (leader) {
      tma_store_wait_all();
      if (pub_mb >= 0) {
        fence_proxy_async_all();
        red_release_gpu_add(p.done + pub_mb, 1u);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS, 1>(tmem_base);
}

}  // namespace hfr
