// The seam between two ResNet bottleneck blocks in ONE launch, with the intermediate operand never re-read from memory:
//
//     Y[M,N1] = act1(A[M,K1] W1^T + b1 (+ R))      block i's "increase" 1x1 convolution + shortcut + ReLU
//     Z[M,N2] = act2(Y W2^T + b2)                  block i+1's "reduce" 1x1 convolution over the tensor just produced
//
// As two launches Y is written (the next shortcut needs it) and read again by the reduce GEMM: N1 * sizeof(T) bytes per
// pixel of HBM traffic and a whole launch (ramp, tail, prologue) that this kernel removes.  A work unit is one 128-row
// block of M for ALL N1 columns.  Every 128-byte-wide chunk of Y that an epilogue warpgroup stages in shared memory for
// its TMA store (128 rows x 64 bf16 / 32 tf32, XOR-swizzled exactly like a SWIZZLE_128B operand tile) is at the same
// time one K-block of the second GEMM's A operand: the MMA warp multiplies it with the matching K-block of W2 straight
// out of the staging buffer into a second accumulator (TMEM columns 256...), and after the unit's last chunk the Z rows
// are drained like any other tile.  The K order of the second GEMM (chunks in ascending column order, four K-slices
// each) is the order the stand-alone kernel uses, so Z is bit-identical to two separate launches.
//
//   warp 0        TMA producer of the weight ring (16 KB slots):  per Y tile  K1/BK slots of W1 k-blocks, then the W2
//                 K-blocks of the PREVIOUS Y tile (N2 rows x 128 bytes each: two per slot for N2 = 64, two slots = two
//                 128-column halves for N2 = 256)
//   warp 3        TMA producer of the unit's A rows: all K1/BK k-blocks of a 128-row block stay RESIDENT for the unit's
//                 N1/128 Y tiles (the first version re-streamed them with W1 for every tile, 8 x 64 KB per unit in
//                 stage 4; resident rows + 16 KB slots: stage-2 seams -10 %, stage 4 unchanged); double-buffered
//                 across units when K1 <= 2 k-blocks
//   warp 1        MMA issuer:  GEMM1(tile q) into accumulator stage q & 1, then GEMM2 of tile q-1 (its chunks are being
//                 staged by an epilogue warpgroup while GEMM1(q) runs), commit -> sfree[buffer], last chunk -> zfull
//   warp 2        TMEM allocator (512 columns: 2 x 128 for Y tiles, N2 <= 256 for Z)
//   warps 4..11   two epilogue warpgroups ping-pong over the job sequence  Y(0) .. Y(T-1) Z | Y(0) ...  of the CTA's
//                 units (job index parity), NBUF staging buffers each
//
// Staging buffer life cycle (buffer = chunk counter % NBUF of its warpgroup):
//   [leader] previous TMA store has finished reading it  &&  previous GEMM2 that read it has retired (sfree)
//   -> residual chunk lands in it by TMA (rfull), PF chunks ahead  -> epilogue adds accumulator, writes the result in
//   place -> fence.proxy.async + warpgroup barrier -> TMA store of the chunk  +  arrive(sfull) -> GEMM2 reads it.
//
// Memory: Z is written while other CTAs still read A, A0 and R, so Z (like Y) must not share memory with any of them -
// the arena planner keeps the first layer's inputs live through the second layer (api.cu plan_arena, seam rule).
//
// Reference semantics being replaced: the `convN_M_1x1_increase` + `Add` + `Relu` and `convN_(M+1)_1x1_reduce` + `Relu`
// nodes of the frozen ResNet-50 graph, evaluated by sess.run (facerec_test.py:114-122).
#pragma once
#include "gemm_tc.cuh"

namespace hfr {

struct PairParams {
  int M, num_m_blocks;   // rows; 128-row blocks = work units
  int N1, K1;            // first GEMM (K1 in elements); N1 is a multiple of 128
  const float* bias1;    // [N1] or nullptr
  const void* residual;  // [M, N1] of T or nullptr
  int act1, round1;
  const float* bias2;    // [N2] or nullptr
  int act2, round2;
  int kb_split;          // the first kb_split k-blocks of A come from the concatenated operand A0 (tmA0), see GemmParams
  int na;                // A buffers (units in flight): 2 when K1 <= 2 k-blocks, else 1
  int stages;            // 16 KB slots of the weight ring (what is left of the shared memory, <= 8)
};

struct PairSmem {
  static constexpr int kSlotBytes = 16384;                 // one k-block: 128 rows x 128 bytes
  static constexpr int kBarBytes = 512;
  static constexpr int kMax = 227 * 1024;
  // dynamic shared memory: [1 KB alignment slack][na x num_kb A k-blocks][stages ring slots][2 x nbuf staging][barriers]
  static int stages_for(int num_kb, int na, int nbuf) {
    const int n = (kMax - 1024 - kBarBytes - (na * num_kb + 2 * nbuf) * kSlotBytes) / kSlotBytes;
    return n > 8 ? 8 : n;
  }
  static int total(int num_kb, int na, int nbuf, int stages) {
    return 1024 + (na * num_kb + stages + 2 * nbuf) * kSlotBytes + kBarBytes;
  }
};

template <typename T, int N2, int NBUF, int PF>
__global__ void __launch_bounds__(384, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmD1, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmD2,
                 const __grid_constant__ CUtensorMap tmA0, const PairParams p) {
  using TR = GemmTraits<T>;
  constexpr int SLOT = PairSmem::kSlotBytes;
  constexpr int BK = TR::BK;
  constexpr int CH_ELEMS = 128 / (int)sizeof(T);     // columns per 128-byte chunk = one K-block of the second GEMM
  constexpr int NCHUNK = 128 / CH_ELEMS;             // chunks per Y tile (2 bf16, 4 tf32)
  constexpr int ZCH = N2 / CH_ELEMS;                 // chunks of a Z tile
  constexpr int HALVES = (N2 == 256) ? 2 : 1;        // a W2 K-block of 256 rows travels as two 128-row slots
  constexpr int N2H = N2 / HALVES;                   // columns per second-GEMM MMA
  constexpr int B2_BYTES = N2H * 128;                // bytes of one W2 K-block (half) in its slot
  constexpr int CPS = (N2 == 64) ? 2 : 1;            // W2 K-blocks per slot
  constexpr uint32_t ACC2_COL = 256;
  static_assert(N2 == 64 || N2 == 128 || N2 == 256, "N2");
  static_assert(NBUF >= 2 && NBUF <= 4 && PF >= 1 && PF < NBUF, "staging buffers / prefetch distance");
  static_assert(NCHUNK % CPS == 0, "W2 K-blocks per slot");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int num_kb = (p.K1 + BK - 1) / BK;
  const int STAGES = p.stages, NA = p.na;
  const uint32_t a_bytes = (uint32_t)num_kb * SLOT;
  const uint32_t sA = smem_base;
  const uint32_t sRing = sA + (uint32_t)NA * a_bytes;
  const uint32_t sEpi = sRing + (uint32_t)STAGES * SLOT;
  const uint32_t sBar = sEpi + 2u * NBUF * SLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (sBar - smem_base) + 400);
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 64u + 8u * s; };
  auto tfull_bar = [&](int s) { return sBar + 128u + 8u * s; };
  auto tempty_bar = [&](int s) { return sBar + 144u + 8u * s; };
  auto afull_bar = [&](int s) { return sBar + 160u + 8u * s; };   // the unit's A rows are resident in A buffer s
  auto aempty_bar = [&](int s) { return sBar + 176u + 8u * s; };  // the unit's last GEMM1 has retired
  auto rfull_bar = [&](int s) { return sBar + 192u + 8u * s; };   // residual chunk landed in staging buffer s
  auto sfull_bar = [&](int s) { return sBar + 256u + 8u * s; };   // Y chunk staged in buffer s (GEMM2 may read it)
  auto sfree_bar = [&](int s) { return sBar + 320u + 8u * s; };   // the GEMM2 MMAs that read buffer s have retired
  const uint32_t zfull_bar = sBar + 384u, zempty_bar = sBar + 392u;

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const bool use_res = (p.residual != nullptr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmD1);
    tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmD2);
    if (use_res) tma_prefetch_desc(&tmR);
    if (p.kb_split > 0) tma_prefetch_desc(&tmA0);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
    }
    for (int s = 0; s < 2 * NBUF; ++s) {
      mbar_init(rfull_bar(s), 1);
      mbar_init(sfull_bar(s), 1);
      mbar_init(sfree_bar(s), 1);
    }
    mbar_init(zfull_bar, 1);
    mbar_init(zempty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512, 1>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int T1 = p.N1 / 128;                        // Y tiles per unit
  const int my_units = ((int)blockIdx.x < p.num_m_blocks) ? (p.num_m_blocks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp loops, one lane issues)
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    };
    auto load_w2 = [&](int j) {   // the W2 K-blocks that meet Y tile j's chunks
#pragma unroll
      for (int c = 0; c < NCHUNK; c += CPS) {
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_bar(stage), CPS * B2_BYTES);
#pragma unroll
            for (int cc = 0; cc < CPS; ++cc)
              tma_load_2d(sRing + stage * SLOT + cc * B2_BYTES, &tmB2, full_bar(stage), j * 128 + (c + cc) * CH_ELEMS, h * 128);
          }
          __syncwarp();
          advance();
        }
      }
    };
    int prev_j = -1;
    for (int ul = 0; ul < my_units; ++ul) {
      for (int j = 0; j < T1; ++j) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_bar(stage), SLOT);
            tma_load_2d(sRing + stage * SLOT, &tmB1, full_bar(stage), kb * BK, j * 128);
          }
          __syncwarp();
          advance();
        }
        if (prev_j >= 0) load_w2(prev_j);
        prev_j = j;
      }
    }
    if (prev_j >= 0) load_w2(prev_j);
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer of the units' A rows
    for (int ul = 0; ul < my_units; ++ul) {
      const int mb = (int)blockIdx.x + ul * (int)gridDim.x;
      const int ab = ul % NA;
      mbar_wait(aempty_bar(ab), (uint32_t)((ul / NA) & 1) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(afull_bar(ab), a_bytes);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (kb < p.kb_split) tma_load_2d(sA + (uint32_t)ab * a_bytes + kb * SLOT, &tmA0, afull_bar(ab), kb * BK, mb * 128);
          else tma_load_2d(sA + (uint32_t)ab * a_bytes + kb * SLOT, &tmA, afull_bar(ab), (kb - p.kb_split) * BK, mb * 128);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
    constexpr uint32_t idesc1 = umma_idesc(TR::kFmt, 128, 128);
    constexpr uint32_t idesc2 = umma_idesc(TR::kFmt, 128, N2H);
    const uint64_t ring_desc0 = umma_desc_sw128(sRing), epi_desc0 = umma_desc_sw128(sEpi), a_desc0 = umma_desc_sw128(sA);
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() {
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    };
    uint32_t ctr0 = 0, ctr1 = 0;      // chunk counters of the two epilogue warpgroups, as they will stand at each job
    uint32_t sfull_par = 0;           // per staging buffer: parity of the next sfull phase to wait for
    // second GEMM of one staged Y tile: (unit-local index ul, tile j, warpgroup wg, its chunk counter at the tile's start)
    auto gemm2 = [&](int ul, int j, uint32_t wg, uint32_t base) {
      if (j == 0) {   // a new unit: the previous unit's Z rows have left the second accumulator
        mbar_wait(zempty_bar, (uint32_t)(ul & 1) ^ 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const uint32_t bi = wg * NBUF + (base + c) % NBUF;
        mbar_wait(sfull_bar(bi), (sfull_par >> bi) & 1u);
        sfull_par ^= 1u << bi;
        const uint32_t a_off = (bi * (uint32_t)SLOT) >> 4;
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          if (c % CPS == 0) mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t b_off = (uint32_t)(stage * SLOT + (c % CPS) * B2_BYTES) >> 4;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma<TR::kTF32>(tmem_base + ACC2_COL + h * 128, epi_desc0 + a_off + 2u * k, ring_desc0 + b_off + 2u * k, idesc2,
                              (j | c | k) != 0);
            if (h == HALVES - 1) umma_commit(sfree_bar(bi));
            if (c % CPS == CPS - 1) umma_commit(empty_bar(stage));
            if (j == T1 - 1 && c == NCHUNK - 1 && h == HALVES - 1) umma_commit(zfull_bar);
          }
          __syncwarp();
          if (c % CPS == CPS - 1) advance();
        }
      }
    };
    int prev_ul = -1, prev_j = 0;
    uint32_t prev_wg = 0, prev_base = 0;
    uint32_t q = 0;                   // Y tiles issued so far (accumulator stage = q & 1)
    for (int ul = 0; ul < my_units; ++ul) {
      for (int j = 0; j < T1; ++j, ++q) {
        const uint32_t as = q & 1, aphase = (q >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 128;
        const int ab = ul % NA;
        if (j == 0) {   // the unit's A rows are resident
          mbar_wait(afull_bar(ab), (uint32_t)((ul / NA) & 1));
          tc_fence_after();
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_off = ((uint32_t)ab * a_bytes + (uint32_t)kb * SLOT) >> 4, b_off = (uint32_t)(stage * SLOT) >> 4;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma<TR::kTF32>(d_tmem, a_desc0 + a_off + 2u * k, ring_desc0 + b_off + 2u * k, idesc1, (kb | k) != 0);
            umma_commit(empty_bar(stage));
            if (kb == num_kb - 1) {
              umma_commit(tfull_bar(as));
              if (j == T1 - 1) umma_commit(aempty_bar(ab));   // the A buffer may take the next unit's rows
            }
          }
          __syncwarp();
          advance();
        }
        if (prev_ul >= 0) gemm2(prev_ul, prev_j, prev_wg, prev_base);
        // job bookkeeping: this Y tile is job ul * (T1 + 1) + j of the CTA; a Z job follows the unit's last Y tile
        const uint32_t job = (uint32_t)(ul * (T1 + 1) + j);
        prev_ul = ul;
        prev_j = j;
        prev_wg = job & 1;
        prev_base = prev_wg ? ctr1 : ctr0;
        if (prev_wg) ctr1 += NCHUNK; else ctr0 += NCHUNK;
        if (j == T1 - 1) {
          if ((job + 1) & 1) ctr1 += ZCH; else ctr0 += ZCH;
        }
      }
    }
    if (prev_ul >= 0) gemm2(prev_ul, prev_j, prev_wg, prev_base);
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: two warpgroups ping-pong on jobs
    const int g = (warp - 4) >> 2;
    const int ew = warp & 3;
    const int row = ew * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
    const uint32_t bar_id = 1 + g;
    const bool leader = (ew == 0 && lane == 0);
    const int JT = T1 + 1;
    const int total_jobs = my_units * JT;
    uint32_t ctr = 0;          // chunks this warpgroup has processed (buffer = ctr % NBUF)
    uint32_t rpar = 0;         // per buffer: parity of the next rfull phase
    uint32_t ypar = 0, yused = 0;   // per buffer: parity of its Y-chunk uses; has it held a Y chunk at all
    // leader only: cursor over this warpgroup's chunk sequence, PF chunks ahead of the one being processed
    int pf_i = g, pf_c = 0;
    uint32_t pf_ctr = 0;
    auto prefetch = [&]() {    // leader only: free the cursor's buffer, request its residual chunk, advance the cursor
      if (pf_i >= total_jobs) return;
      const uint32_t b = pf_ctr % NBUF, bi = g * NBUF + b;
      if (pf_ctr >= (uint32_t)NBUF) {
        tma_store_wait_read<NBUF - PF - 1>();    // the store of chunk pf_ctr - NBUF has finished reading the buffer
        if ((yused >> b) & 1u) mbar_wait(sfree_bar(bi), ((ypar >> b) & 1u) ^ 1u);   // ... and so has its GEMM2
      }
      const int ul = pf_i / JT, j = pf_i - ul * JT;
      if (j < T1 && use_res) {
        mbar_expect_tx(rfull_bar(bi), SLOT);
        tma_load_2d(sEpi + bi * SLOT, &tmR, rfull_bar(bi), j * 128 + pf_c * CH_ELEMS,
                    ((int)blockIdx.x + ul * (int)gridDim.x) * 128);
      }
      ++pf_ctr;
      if (++pf_c == (j < T1 ? NCHUNK : ZCH)) {
        pf_c = 0;
        pf_i += 2;
      }
    };
    if (leader) {
#pragma unroll
      for (int i = 0; i < PF; ++i) prefetch();
    }
    for (int i = g; i < total_jobs; i += 2) {
      const int ul = i / JT, j = i - ul * JT;
      const int mb = (int)blockIdx.x + ul * (int)gridDim.x;
      const bool isZ = (j == T1);
      uint32_t acc_col, as = 0;
      if (!isZ) {
        const uint32_t q = (uint32_t)(ul * T1 + j);
        as = q & 1;
        mbar_wait(tfull_bar(as), (q >> 1) & 1);
        acc_col = as * 128;
      } else {
        mbar_wait(zfull_bar, (uint32_t)(ul & 1));
        acc_col = ACC2_COL;
      }
      tc_fence_after();
      const int nch = isZ ? ZCH : NCHUNK;
      const float* bias = isZ ? p.bias2 : p.bias1;
      const int act = isZ ? p.act2 : p.act1;
      const int rnd = isZ ? p.round2 : p.round1;
      const bool res = use_res && !isZ;
      for (int c = 0; c < nch; ++c, ++ctr) {
        const int n0 = (isZ ? 0 : j * 128) + c * CH_ELEMS;
        const uint32_t b = ctr % NBUF, bi = g * NBUF + b;
        const uint32_t st_row = sEpi + bi * SLOT + row * 128;
        // All TMEM loads of the chunk and the first bias vector go out first (asynchronous): their latency overlaps the
        // leader's buffer hand-over and the wait for the residual chunk below.
        constexpr int NH = CH_ELEMS / 32;
        uint32_t racc[NH][32];
        float4 bb[8];
        auto load_bias = [&](int h) {
          if (bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(bias + n0 + h * 32);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) bb[qq] = __ldg(b4 + qq);
          }
        };
        load_bias(0);
#pragma unroll
        for (int h = 0; h < NH; ++h) tmem_ld_32x32(tmem_base + lane_addr + acc_col + c * CH_ELEMS + h * 32, racc[h]);
        if (leader) prefetch();   // the chunk PF ahead: frees its buffer (used NBUF - PF chunks ago), requests its residual
        uint4 rv[8];
        if (res) {
          mbar_wait(rfull_bar(bi), (rpar >> b) & 1u);
          rpar ^= 1u << b;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const uint32_t a = st_row + (((uint32_t)qq ^ (row & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rv[qq].x), "=r"(rv[qq].y), "=r"(rv[qq].z), "=r"(rv[qq].w)
                         : "r"(a));
          }
        }
        // (without a residual the buffer was freed by the leader PF chunks ago, ahead of a warpgroup barrier)
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          uint32_t (&r)[32] = racc[h];
          float v[32];
          if (bias != nullptr) {
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              v[4 * qq] = __uint_as_float(r[4 * qq]) + bb[qq].x;
              v[4 * qq + 1] = __uint_as_float(r[4 * qq + 1]) + bb[qq].y;
              v[4 * qq + 2] = __uint_as_float(r[4 * qq + 2]) + bb[qq].z;
              v[4 * qq + 3] = __uint_as_float(r[4 * qq + 3]) + bb[qq].w;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = __uint_as_float(r[jj]);
          }
          if (h + 1 < NH) load_bias(h + 1);   // in flight during this half's residual add / pack / store
          if (res) {
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
            for (int jj = 0; jj < 32; ++jj) {
              float x = v[jj];
              if (act == ACT_RELU) x = fmaxf(x, 0.f);
              if (act == ACT_RELU6) x = fminf(fmaxf(x, 0.f), 6.f);
              v[jj] = x;
            }
            if (rnd) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) v[jj] = round_tf32(v[jj]);
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
              for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2_act(v[8 * qq + 2 * e], v[8 * qq + 2 * e + 1], act);
              const uint32_t a = st_row + (((uint32_t)(h * 4 + qq) ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                           "r"(w[3]));
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (leader) {
          tma_store_2d(isZ ? &tmD2 : &tmD1, sEpi + bi * SLOT, n0, mb * 128);
          tma_store_commit();
          if (!isZ) mbar_arrive(sfull_bar(bi));   // the chunk is complete in shared memory: GEMM2 may read it
        }
        if (!isZ) {
          ypar ^= 1u << b;
          yused |= 1u << b;
        }
      }
      // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(isZ ? zempty_bar : tempty_bar(as));
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512, 1>(tmem_base);
}

}  // namespace hfr
