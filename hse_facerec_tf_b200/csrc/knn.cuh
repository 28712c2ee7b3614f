// 1-NN / k-NN helper kernels around the distance GEMM (gemm_tc.cuh, EPI_KNN / EPI_KNN4):
//   rows_prep_kernel     fp32 rows -> bf16 copy + squared norms (+ the shard's largest squared norm)
//   knn_finalize_kernel  merge the per-(split, warpgroup) candidate records, re-score the best ones in fp64, CERTIFY
//   knn_exact_kernel     fp64 brute force over the whole shard for the queries that could not be certified
//   knn_merge_kernel     merge of per-shard results after the all-gather
//
// Reference semantics: sklearn KNeighborsClassifier(n_neighbors, p=2).kneighbors (facerec_test.py:272-275,284-285):
// ArgKmin over d2 = |x|^2 + |y|^2 - 2 x.y evaluated in fp64 on the float32 rows (_argkmin.pyx.tp, _middle_term_computer),
// the heap keeps the first-seen row on exact ties.  Here: the result is the fp64 brute-force answer, ties to the lowest
// gallery index, for EVERY query - proven per query by the bound below or recomputed.
//
// Certification.  The GEMM's score of gallery row j is  s~_j = fl32(gn_j - 2 acc_j),  acc_j the tensor-core dot product
// of the operands rounded to bf16 (round to nearest, u = 2^-9) or truncated to tf32 by the MMA (u = 2^-10), gn_j the fp32
// squared norm.  Against the exact  s_j = |g_j|^2 - 2 q.g_j :
//     |s~_j - s_j| <= 2 ((2u + u^2) + c_acc) |q| |g_j|  +  c_norm |g_j|^2  =: E        (Cauchy-Schwarz on sum |q_i g_ji|)
// with c_acc covering the fp32 accumulation of D exact products (taken as max(D 2^-22, 2^-12): twice the round-to-nearest
// figure, tensor-core adders may truncate) and c_norm = (D + 4) 2^-24 the fp32 norm / final fma roundings.  E is
// evaluated with |g_j| <= the shard's largest norm.  Every row that was not re-scored either lost inside its bucket
// (s~ >= the bucket's last kept score) or lost the merge (s~ >= the first score not re-scored); with t the minimum of
// those, all such rows have  d2_j = |q|^2 + s_j >= |q|^2 + t - E.  The k-th re-scored distance being strictly below that
// value proves the answer; otherwise the query goes to knn_exact_kernel.
#pragma once
#include "ptx.cuh"

namespace hfr {

struct Neighbor {  // layout of hfr_neighbor (include/hfr.h)
  double dist2;
  long long index;
};

// rows fp32 -> bf16 copy (optional) + squared L2 norms in fp32 (optional) + max squared norm (optional).  Warp per row.
__global__ void __launch_bounds__(256) rows_prep_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                                                        float* __restrict__ nrm, float* __restrict__ max_nrm,
                                                        long long n, int d) {
  __shared__ float wmax[8];
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  if (row < n) {
    const float* xr = x + row * d;
    for (int j = lane * 4; j < d; j += 128) {
      const float4 v = *reinterpret_cast<const float4*>(xr + j);
      s = fmaf(v.x, v.x, s);
      s = fmaf(v.y, v.y, s);
      s = fmaf(v.z, v.z, s);
      s = fmaf(v.w, v.w, s);
      if (xb) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        *reinterpret_cast<uint2*>(xb + row * d + j) = o;
      }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0 && nrm) nrm[row] = s;
  }
  if (max_nrm) {  // block maximum, then one atomic per block (non-negative floats order like their bit patterns)
    if (lane == 0) wmax[threadIdx.x >> 5] = (row < n && s == s) ? s : 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = wmax[0];
      for (int i = 1; i < 8; ++i) m = fmaxf(m, wmax[i]);
      atomicMax(reinterpret_cast<int*>(max_nrm), __float_as_int(m));
    }
  }
}

// Squared distance of two float32 rows accumulated in fp64 by one warp (lane-strided partial sums, xor-shuffle tree):
// THE distance every result carries - the certified path and the exact pass both finish with it, so a query's output
// does not depend on the path it took (nor on how the gallery is sharded).
__device__ __forceinline__ double warp_dist2_fp64(const float* __restrict__ a, const float* __restrict__ b, int d, int lane) {
  double acc = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double df = (double)a[j] - (double)b[j];
    acc = fma(df, df, acc);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

struct KnnFinalizeParams {
  const float* q;            // [nq][d] fp32 queries
  const float* g;            // [n][d] fp32 gallery shard
  const float* part_score;   // [nq][nbuckets][NC]
  const int* part_idx;
  int nbuckets;              // gallery splits x 2 epilogue warpgroups
  long long nq;
  int d;
  long long row_offset;
  int k;                     // neighbours wanted (<= 4)
  double c_dot, c_norm;      // bound coefficients, see the file header
  const float* gmax2;        // largest squared gallery norm of the shard
  Neighbor* out;             // [nq][k]
  int* unc_list;             // queries that need the exact pass
  int* counters;             // [0] number of uncertified queries
  int partial;               // sharded gallery: out is [nq][k + 1] = the k re-scored candidates of THIS shard (always,
                             // certified or not) + {lower bound on d2 of every row of the shard that was not re-scored,
                             // index -2}; the certification happens after the exchange, against the GLOBAL k-th distance
};

// Warp per query.  NC = candidates per bucket record (2 / 4), KEEP = candidates re-scored in fp64.
template <int NC, int KEEP>
__global__ void __launch_bounds__(256) knn_finalize_kernel(const KnnFinalizeParams p) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.nq) return;
  const int ncand = p.nbuckets * NC;
  // each lane keeps its own KEEP best, then the warp extracts the global KEEP best one at a time
  float ls[KEEP];
  int li[KEEP];
#pragma unroll
  for (int t = 0; t < KEEP; ++t) {
    ls[t] = INFINITY;
    li[t] = -1;
  }
  float t_bucket = INFINITY;  // smallest "last kept score" over the buckets that dropped rows
  for (int j = lane; j < ncand; j += 32) {
    float s = p.part_score[row * ncand + j];
    int i = p.part_idx[row * ncand + j];
    if (i < 0) continue;
    if ((j % NC) == NC - 1) t_bucket = fminf(t_bucket, s);  // a full record: anything the bucket dropped scores >= s
#pragma unroll
    for (int t = 0; t < KEEP; ++t) {
      if (s < ls[t] || (s == ls[t] && i < li[t])) {
        const float ts = ls[t];
        const int ti = li[t];
        ls[t] = s;
        li[t] = i;
        s = ts;
        i = ti;
      }
    }
  }
  const float* qr = p.q + row * p.d;
  double qn = 0.0;
  for (int j = lane; j < p.d; j += 32) {
    const double v = (double)qr[j];
    qn = fma(v, v, qn);
  }
  for (int o = 16; o; o >>= 1) {
    qn += __shfl_xor_sync(0xffffffffu, qn, o);
    t_bucket = fminf(t_bucket, __shfl_xor_sync(0xffffffffu, t_bucket, o));
  }
  double bd[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
  int bi[4] = {-1, -1, -1, -1};
  float t_next = INFINITY;  // best approximate score that was NOT re-scored
  for (int c = 0; c <= KEEP; ++c) {
    float hs = ls[0];
    int hi = li[0];
    for (int o = 16; o; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, hs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, hi, o);
      if (oi >= 0 && (hi < 0 || os < hs || (os == hs && oi < hi))) {
        hs = os;
        hi = oi;
      }
    }
    if (hi < 0) break;
    if (c == KEEP) {
      t_next = hs;
      break;
    }
    if (li[0] == hi) {  // pop it from the owning lane
#pragma unroll
      for (int t = 0; t + 1 < KEEP; ++t) {
        ls[t] = ls[t + 1];
        li[t] = li[t + 1];
      }
      ls[KEEP - 1] = INFINITY;
      li[KEEP - 1] = -1;
    }
    const double acc = warp_dist2_fp64(qr, p.g + (long long)hi * p.d, p.d, lane);
    // sorted insert into the exact top-4 (ascending distance, ties to the lowest index)
    double cd = acc;
    int ci = hi;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (ci >= 0 && (bi[t] < 0 || cd < bd[t] || (cd == bd[t] && ci < bi[t]))) {
        const double td = bd[t];
        const int ti = bi[t];
        bd[t] = cd;
        bi[t] = ci;
        cd = td;
        ci = ti;
      }
    }
  }
  // certification (file header): every row not re-scored has d2 >= qn + t - E
  const double t = (double)fminf(t_bucket, t_next);
  const double gm2 = (double)__ldg(p.gmax2);
  const double E = (p.c_dot * sqrt(qn) * sqrt(gm2) + p.c_norm * gm2) * (1.0 + 1e-9) + 1e-300;
  const bool certified = (t == (double)INFINITY) || (bd[p.k - 1] < qn + t - E);
  if (p.partial) {
    if (lane == 0) {
      Neighbor* o = p.out + row * (p.k + 1);
      for (int j = 0; j < p.k; ++j) {
        Neighbor nb;
        nb.dist2 = bi[j] >= 0 ? bd[j] : (double)INFINITY;
        nb.index = bi[j] >= 0 ? p.row_offset + bi[j] : -1;
        o[j] = nb;
      }
      Neighbor bnd;
      bnd.dist2 = (t == (double)INFINITY) ? (double)INFINITY : qn + t - E;
      bnd.index = -2;
      o[p.k] = bnd;
    }
    return;
  }
  if (lane == 0) {
    for (int j = 0; j < p.k; ++j) {
      Neighbor nb;
      nb.dist2 = certified ? bd[j] : (double)INFINITY;   // uncertified: the exact pass starts from an empty list
      nb.index = (certified && bi[j] >= 0) ? p.row_offset + bi[j] : -1;
      p.out[row * p.k + j] = nb;
    }
    if (!certified) p.unc_list[atomicAdd(p.counters, 1)] = (int)row;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Exact pass: squared distances of the listed queries to EVERY row of the shard, in fp64 on the float32 rows
// (d2 = |q|^2 + |g|^2 - 2 q.g, each term accumulated in double: sklearn's formula).  Work item = 16 queries x a chunk of
// gallery rows; a CTA of 256 threads walks the chunk 256 rows at a time (one row per thread, 16 accumulators each),
// survivors of a running threshold go through a small shared-memory queue into the tile's top-4 lists, and at the end of
// the chunk one thread per query merges its list into the global result under a per-query lock.
constexpr int kExQ = 16;        // queries per tile
constexpr int kExRows = 256;    // gallery rows per iteration (= threads)
constexpr int kExK = 32;        // k-slab
// gallery rows per work item: chosen in the kernel from the number of listed queries so that there are ~4 items per CTA

struct KnnExactSmem {
  double qd[kExK][kExQ];            // query slab, converted once
  float gs[kExRows][kExK + 4];      // gallery slab (row pitch 36 floats: conflict-free 16-byte reads)
  double qn[kExQ];
  double top_d[kExQ][4];
  long long top_i[kExQ][4];
  double thr[kExQ];
  int qrow[kExQ];
  int cnt[kExQ];
  double pend_d[kExQ][kExRows];
  int pend_r[kExQ][kExRows];
};

__global__ void __launch_bounds__(256) knn_exact_kernel(const float* __restrict__ q, const float* __restrict__ g,
                                                         long long n, int d, long long row_offset, int k,
                                                         const int* __restrict__ unc_list, const int* __restrict__ counters,
                                                         int* __restrict__ locks, Neighbor* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char exact_smem_raw[];
  KnnExactSmem& sm = *reinterpret_cast<KnnExactSmem*>(exact_smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nu = counters[0];
  if (nu == 0) return;
  const int tiles = (nu + kExQ - 1) / kExQ;
  long long chunk_rows = (n * tiles / (4ll * gridDim.x) + kExRows - 1) / kExRows * kExRows;
  chunk_rows = max((long long)kExRows, min(chunk_rows, 8192ll));
  const long long chunks = (n + chunk_rows - 1) / chunk_rows;
  const long long items = (long long)tiles * chunks;
  for (long long w = blockIdx.x; w < items; w += gridDim.x) {
    const int tile = (int)(w % tiles);   // CTAs running side by side share a gallery chunk (L2), not a query tile
    const long long r_begin = (w / tiles) * chunk_rows;
    const long long r_end = min(r_begin + chunk_rows, n);
    __syncthreads();  // previous item's lists are no longer read
    if (tid < kExQ) {
      const int u = tile * kExQ + tid;
      sm.qrow[tid] = u < nu ? unc_list[u] : -1;
      sm.thr[tid] = INFINITY;
      sm.cnt[tid] = 0;
      for (int t = 0; t < 4; ++t) {
        sm.top_d[tid][t] = INFINITY;
        sm.top_i[tid][t] = -1;
      }
    }
    __syncthreads();
    // squared query norms in fp64: warp w handles queries 2w, 2w + 1
    for (int j = warp * 2; j < warp * 2 + 2; ++j) {
      const int qr = sm.qrow[j];
      double s = 0.0;
      if (qr >= 0)
        for (int c = lane; c < d; c += 32) {
          const double v = (double)q[(long long)qr * d + c];
          s = fma(v, v, s);
        }
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) sm.qn[j] = s;
    }
    for (long long r0 = r_begin; r0 < r_end; r0 += kExRows) {
      double acc[kExQ];
#pragma unroll
      for (int j = 0; j < kExQ; ++j) acc[j] = 0.0;
      double gn = 0.0;
      for (int k0 = 0; k0 < d; k0 += kExK) {
        __syncthreads();  // the previous slab has been consumed
        for (int e = tid; e < kExK * kExQ; e += 256) {
          const int j = e / kExK, kk = e % kExK;   // consecutive threads read consecutive floats of one query row
          const int qr = sm.qrow[j];
          sm.qd[kk][j] = (qr >= 0 && k0 + kk < d) ? (double)q[(long long)qr * d + k0 + kk] : 0.0;
        }
        for (int e = tid; e < kExRows * (kExK / 4); e += 256) {
          const int r = e / (kExK / 4), c4 = e % (kExK / 4);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          // d is a multiple of 4 (rows are 16-byte multiples), so a float4 never straddles the row end
          if (r0 + r < r_end && k0 + c4 * 4 < d) v = *reinterpret_cast<const float4*>(g + (r0 + r) * d + k0 + c4 * 4);
          *reinterpret_cast<float4*>(&sm.gs[r][c4 * 4]) = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int c4 = 0; c4 < kExK / 4; ++c4) {
          const float4 gv4 = *reinterpret_cast<const float4*>(&sm.gs[tid][c4 * 4]);
          const float gvf[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const double gv = (double)gvf[e];
            gn = fma(gv, gv, gn);
            const double2* qv = reinterpret_cast<const double2*>(&sm.qd[c4 * 4 + e][0]);
#pragma unroll
            for (int j2 = 0; j2 < kExQ / 2; ++j2) {
              const double2 qq = qv[j2];
              acc[2 * j2] = fma(qq.x, gv, acc[2 * j2]);
              acc[2 * j2 + 1] = fma(qq.y, gv, acc[2 * j2 + 1]);
            }
          }
        }
      }
      const long long row = r0 + tid;
      if (row < r_end) {
#pragma unroll
        for (int j = 0; j < kExQ; ++j) {
          const double d2 = fmax(sm.qn[j] + gn - 2.0 * acc[j], 0.0);
          if (sm.qrow[j] >= 0 && d2 <= sm.thr[j]) {   // '<=': equal distances go on to the index tie-break below
            const int pos = atomicAdd(&sm.cnt[j], 1);
            sm.pend_d[j][pos] = d2;
            sm.pend_r[j][pos] = (int)(row - r_begin);
          }
        }
      }
      __syncthreads();
      if (tid < kExQ) {  // one thread per query drains its queue into the sorted top-4 list
        const int c = sm.cnt[tid];
        for (int e = 0; e < c; ++e) {
          double cd = sm.pend_d[tid][e];
          long long ci = r_begin + sm.pend_r[tid][e];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const double td = sm.top_d[tid][t];
            const long long ti = sm.top_i[tid][t];
            if (ci >= 0 && (ti < 0 || cd < td || (cd == td && ci < ti))) {
              sm.top_d[tid][t] = cd;
              sm.top_i[tid][t] = ci;
              cd = td;
              ci = ti;
            }
          }
        }
        sm.cnt[tid] = 0;
        sm.thr[tid] = sm.top_d[tid][k - 1];
      }
      // (the next iteration's first __syncthreads orders these updates before the next threshold tests)
    }
    __syncthreads();
    if (tid < kExQ && sm.qrow[tid] >= 0) {
      // merge this chunk's list into the global result of the query; chunks of the same query run on other CTAs
      const int qr = sm.qrow[tid];
      while (atomicCAS(&locks[qr], 0, 1) != 0) __nanosleep(64);
      __threadfence();
      volatile Neighbor* o = out + (long long)qr * k;
      double md[4];
      long long mi[4];
      for (int t = 0; t < 4; ++t) {
        md[t] = t < k ? o[t].dist2 : (double)INFINITY;
        mi[t] = t < k ? o[t].index : -1;
      }
      for (int e = 0; e < k; ++e) {
        double cd = sm.top_d[tid][e];
        long long ci = sm.top_i[tid][e] < 0 ? -1 : sm.top_i[tid][e] + row_offset;
        for (int t = 0; t < 4; ++t) {
          if (ci >= 0 && (mi[t] < 0 || cd < md[t] || (cd == md[t] && ci < mi[t]))) {
            const double td = md[t];
            const long long ti = mi[t];
            md[t] = cd;
            mi[t] = ci;
            cd = td;
            ci = ti;
          }
        }
      }
      for (int t = 0; t < k; ++t) {
        o[t].dist2 = md[t];
        o[t].index = mi[t];
      }
      __threadfence();
      atomicExch(&locks[qr], 0);
    }
  }
}

// After the exact pass: the listed queries' neighbours are final as a SET; give them the same distance arithmetic as the
// certified path (warp_dist2_fp64) and re-sort by (dist2, index).  Warp per listed query.
__global__ void __launch_bounds__(256) knn_rescore_kernel(const float* __restrict__ q, const float* __restrict__ g, int d,
                                                           long long row_offset, int k, const int* __restrict__ unc_list,
                                                           const int* __restrict__ counters, Neighbor* __restrict__ out) {
  const int nu = counters[0];
  const int lane = threadIdx.x & 31;
  for (int u = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < nu; u += gridDim.x * (blockDim.x >> 5)) {
    const int qr = unc_list[u];
    Neighbor* o = out + (long long)qr * k;
    double md[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    long long mi[4] = {-1, -1, -1, -1};
    for (int e = 0; e < k; ++e) {
      long long ci = o[e].index;
      if (ci < 0) continue;
      double cd = warp_dist2_fp64(q + (long long)qr * d, g + (ci - row_offset) * d, d, lane);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (ci >= 0 && (mi[t] < 0 || cd < md[t] || (cd == md[t] && ci < mi[t]))) {
          const double td = md[t];
          const long long ti = mi[t];
          md[t] = cd;
          mi[t] = ci;
          cd = td;
          ci = ti;
        }
      }
    }
    __syncwarp();
    if (lane == 0)
      for (int t = 0; t < k; ++t) {
        o[t].dist2 = md[t];
        o[t].index = mi[t];
      }
  }
}

// Sharded gallery, step 2 (after the all-gather of the partial records [P][nq][k + 1], see KnnFinalizeParams::partial):
// the global k nearest among all shards' re-scored candidates, certified against the smallest of the shards' bounds -
// every row of every shard that was not re-scored is at least that far away.  A shard that does not hold a query's
// neighbour therefore costs nothing (certifying per shard sent ~5 % of such queries through the shard's exact pass).
// Uncertified queries are listed; every rank computes the same list from the same gathered records.
__global__ void knn_merge_certify_kernel(const Neighbor* __restrict__ parts, int nparts, long long nq, int k,
                                         Neighbor* __restrict__ out, int* __restrict__ unc_list, int* __restrict__ unc_count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  double md[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
  long long mi[4] = {-1, -1, -1, -1};
  double bound = INFINITY;
  for (int p = 0; p < nparts; ++p) {
    const Neighbor* rec = parts + ((size_t)p * nq + i) * (k + 1);
    bound = fmin(bound, rec[k].dist2);
    for (int e = 0; e < k; ++e) {
      double cd = rec[e].dist2;
      long long ci = rec[e].index;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (ci >= 0 && (mi[t] < 0 || cd < md[t] || (cd == md[t] && ci < mi[t]))) {
          const double td = md[t];
          const long long ti = mi[t];
          md[t] = cd;
          mi[t] = ci;
          cd = td;
          ci = ti;
        }
      }
    }
  }
  for (int t = 0; t < k; ++t) {
    Neighbor nb;
    nb.dist2 = md[t];
    nb.index = mi[t];
    out[i * k + t] = nb;
  }
  const bool certified = (bound == (double)INFINITY) || (md[k - 1] < bound);
  if (!certified) unc_list[atomicAdd(unc_count, 1)] = (int)i;
}

// Sharded gallery, step 3 (only when step 2 listed queries): the listed rows of a [nq][k] result start empty ...
__global__ void knn_clear_listed_kernel(const int* __restrict__ unc_list, const int* __restrict__ unc_count, int k,
                                        Neighbor* __restrict__ out) {
  const int nu = unc_count[0];
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nu * k; u += gridDim.x * blockDim.x) {
    Neighbor nb;
    nb.dist2 = INFINITY;
    nb.index = -1;
    out[(long long)unc_list[u / k] * k + u % k] = nb;
  }
}
// ... knn_exact_kernel + knn_rescore_kernel fill them from the local shard, and after the second all-gather the listed
// rows of the result are replaced by the merge of the shards' exact answers.
__global__ void knn_merge_listed_kernel(const Neighbor* __restrict__ parts, int nparts, long long nq, int k,
                                        const int* __restrict__ unc_list, const int* __restrict__ unc_count,
                                        Neighbor* __restrict__ out) {
  const int nu = unc_count[0];
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x) {
    const long long i = unc_list[u];
    double md[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    long long mi[4] = {-1, -1, -1, -1};
    for (int p = 0; p < nparts; ++p)
      for (int e = 0; e < k; ++e) {
        const Neighbor nb = parts[((size_t)p * nq + i) * k + e];
        double cd = nb.dist2;
        long long ci = nb.index;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (ci >= 0 && (mi[t] < 0 || cd < md[t] || (cd == md[t] && ci < mi[t]))) {
            const double td = md[t];
            const long long ti = mi[t];
            md[t] = cd;
            mi[t] = ci;
            cd = td;
            ci = ti;
          }
        }
      }
    for (int t = 0; t < k; ++t) {
      Neighbor nb;
      nb.dist2 = md[t];
      nb.index = mi[t];
      out[i * k + t] = nb;
    }
  }
}

// Merge P per-shard results (gathered as [P][nq][k] records) into the global k nearest: ascending (dist2, index).
__global__ void knn_merge_kernel(const Neighbor* __restrict__ parts, int nparts, long long nq, int k,
                                 Neighbor* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  double md[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
  long long mi[4] = {-1, -1, -1, -1};
  for (int p = 0; p < nparts; ++p)
    for (int e = 0; e < k; ++e) {
      const Neighbor nb = parts[((size_t)p * nq + i) * k + e];
      double cd = nb.dist2;
      long long ci = nb.index;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (ci >= 0 && (mi[t] < 0 || cd < md[t] || (cd == md[t] && ci < mi[t]))) {
          const double td = md[t];
          const long long ti = mi[t];
          md[t] = cd;
          mi[t] = ci;
          cd = td;
          ci = ti;
        }
      }
    }
  for (int t = 0; t < k; ++t) {
    Neighbor nb;
    nb.dist2 = md[t];
    nb.index = mi[t];
    out[i * k + t] = nb;
  }
}

}  // namespace hfr
