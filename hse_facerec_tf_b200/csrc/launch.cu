// Kernel launchers: tensor-map construction (driver entry points resolved at run time, no libcuda link dependency),
// tile-shape selection and grid sizing for the sm_100a kernels in kernels.cuh / gemm_tc.cuh.
#include "launch.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <utility>
#include <cstring>
#include <mutex>

#include "conv_window.cuh"
#include "gemm_pair.cuh"
#include "kernels.cuh"

namespace hfr {

static std::atomic<int64_t> g_launches{0};
int64_t launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }

void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw Error(-6, std::string(what) + ": " + cudaGetErrorString(e));
}
#define HFR_LAUNCH_CHECK(name)                    \
  do {                                            \
    g_launches.fetch_add(1);                      \
    cuda_check(cudaGetLastError(), "launch " name); \
  } while (0)

// Launch with the programmatic-dependent-launch attribute (every kernel launched through here calls pdl_wait()).
static bool pdl_enabled() {
  static const bool on = getenv("HFR_NO_PDL") == nullptr;
  return on;
}
// cluster_x > 1: launch as thread-block clusters of that many CTAs along x (the grid must be a multiple of it)
template <typename... KArgs, typename... Args>
static void launch_pdl_cluster(void (*kern)(KArgs...), int cluster_x, dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                               Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cuda_check(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...), "cudaLaunchKernelEx");
}
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  launch_pdl_cluster(kern, 1, grid, block, smem, s, std::forward<Args>(args)...);
}

int device_sm_count(int device) {
  static std::mutex mu;
  static int cache[64];
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64) throw Error(-1, "bad device index");
  if (cache[device] == 0) {
    cudaDeviceProp p;
    cuda_check(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties");
    if (p.major != 10) throw Error(-6, "device " + std::to_string(device) + " is sm_" + std::to_string(p.major) +
                                           std::to_string(p.minor) + "; this library contains sm_100a code only");
    cache[device] = p.multiProcessorCount;
  }
  return cache[device];
}

void use_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) throw Error(-6, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= n) throw Error(-1, "device index out of range");
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  device_sm_count(device);
}

static unsigned grid_for(long long total, int threads, long long cap = 148LL * 32) {
  long long g = (total + threads - 1) / threads;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// ---------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void* driver_fn(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cuda_check(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q), name);
  if (q != cudaDriverEntryPointSuccess || !fn) throw Error(-6, std::string("driver entry point not found: ") + name);
  return fn;
}

static CUtensorMapDataType tmap_dtype(int prec) {
  return prec == PREC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

// rank-N tiled map; dims/strides innermost first; strides in bytes for dims 1..rank-1
static CUtensorMap make_tiled(const void* ptr, int prec, int rank, const uint64_t* dims, const uint64_t* strides,
                              const uint32_t* box, CUtensorMapSwizzle swz) {
  static PFN_encodeTiled enc = (PFN_encodeTiled)driver_fn("cuTensorMapEncodeTiled");
  CUtensorMap m;
  uint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&m, tmap_dtype(prec), (cuuint32_t)rank, const_cast<void*>(ptr), (const cuuint64_t*)dims,
                   (const cuuint64_t*)strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-6, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}

// 2-D row-major [rows, cols] matrix, box = {128 bytes of columns, box_rows}, SWIZZLE_128B
static CUtensorMap make_tmap_2d(const void* ptr, int prec, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {cols * elt_size(prec)};
  const uint32_t box[2] = {(uint32_t)(128 / elt_size(prec)), box_rows};
  return make_tiled(ptr, prec, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// ---------------------------------------------------------------------------------------------- stem
template <typename TIn, typename T>
static void launch_stem_t(const StemArgs& a, cudaStream_t s) {
  StemParams p;
  p.B = a.B; p.H = a.H; p.W = a.W; p.Ho = a.Ho; p.Wo = a.Wo; p.KH = a.kh; p.KW = a.kw; p.stride = a.stride;
  p.pad_t = a.pad_t; p.pad_l = a.pad_l; p.Cout = a.cout; p.flip = a.flip; p.scale = a.scale;
  for (int i = 0; i < 3; ++i) p.mean[i] = a.mean[i];
  p.act = a.act; p.round_tf32 = a.round_tf32;
  const long long npix = (long long)a.B * a.Ho * a.Wo;
  dim3 grid((unsigned)((npix + 127) / 128), (unsigned)(a.cout / 32));
  const size_t smem = (size_t)a.kh * a.kw * 3 * 32 * sizeof(float);
  launch_pdl(stem_conv_kernel<TIn, T>, grid, dim3(128), smem, s, (const TIn*)a.x, a.w, a.bias, (T*)a.y, p);
  HFR_LAUNCH_CHECK("stem_conv");
}
void launch_stem(const StemArgs& a, int prec, cudaStream_t s) {
  if (a.cout % 32) throw Error(-1, "stem: cout must be a multiple of 32");
  if ((size_t)a.kh * a.kw * 3 * 32 * 4 > 48 * 1024) throw Error(-5, "stem: kernel window too large");
  if (prec == PREC_BF16) {
    if (a.in_u8) launch_stem_t<uint8_t, __nv_bfloat16>(a, s); else launch_stem_t<float, __nv_bfloat16>(a, s);
  } else {
    if (a.in_u8) launch_stem_t<uint8_t, float>(a, s); else launch_stem_t<float, float>(a, s);
  }
}

// ---------------------------------------------------------------------------------------------- depthwise
// persistent pipelined kernel: grid = resident CTAs (a multiple of the channel-block count), stages sized to ~52 KB of
// windows per CTA
template <typename T, int STRIDE, int VL>
static void launch_dw_pipe_t(const DwArgs& a, int prec, int device, cudaStream_t s) {
  constexpr int VN = Vec16<T>::N;
  constexpr int CBE = VL * VN;
  constexpr int TWI = 7 * STRIDE + 3;
  const int es = (int)sizeof(T);
  if (a.C % CBE) throw Error(-1, "depthwise: channels must be a multiple of the channel block");
  const uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
  const uint64_t strides[3] = {(uint64_t)a.C * es, (uint64_t)a.W * a.C * es, (uint64_t)a.H * a.W * a.C * es};
  const uint32_t box[4] = {(uint32_t)CBE, (uint32_t)TWI, (uint32_t)TWI, 1};
  CUtensorMap tm = make_tiled(a.x, prec, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  DwPipeParams p;
  p.C = a.C; p.Ho = a.Ho; p.Wo = a.Wo; p.pad_t = a.pad_t; p.pad_l = a.pad_l;
  p.tiles_w = (a.Wo + 7) / 8;
  p.tiles_h = (a.Ho + 7) / 8;
  p.cblocks = a.C / CBE;
  p.act = a.act; p.round_tf32 = a.round_tf32;
  p.stage_bytes = (TWI * TWI * CBE * es + 127) / 128 * 128;
  p.stages = (52 * 1024) / p.stage_bytes;
  if (p.stages < 2) p.stages = 2;
  if (p.stages > 6) p.stages = 6;
  const long long num_sp = (long long)a.B * p.tiles_w * p.tiles_h;
  if (num_sp * p.cblocks >= (1ll << 31)) throw Error(-1, "depthwise: too many tiles for one launch");
  p.num_sp = (int)num_sp;
  const long long num_tiles = num_sp * p.cblocks;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 128;
  auto kern = dwconv3x3_pipe_kernel<T, STRIDE, VL>;
  static std::atomic<int> ctas_per_sm[64];  // per instantiation; the stage count is a function of the instantiation only
  if (ctas_per_sm[device].load() == 0) {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "dw smem attribute");
    int occ = 0;
    cuda_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 16 * VL + 32, smem), "dw occupancy");
    if (occ < 1) throw Error(-6, "depthwise: kernel does not fit on an SM");
    ctas_per_sm[device].store(occ);
  }
  long long grid = (long long)device_sm_count(device) * ctas_per_sm[device].load();
  if (grid > num_tiles) grid = num_tiles;
  grid = grid / p.cblocks * p.cblocks;      // every CTA keeps one channel block
  if (grid < p.cblocks) grid = p.cblocks;
  launch_pdl(kern, dim3((unsigned)grid), dim3(16 * VL + 32), smem, s, tm, a.w, a.bias, (T*)a.y, p);
  HFR_LAUNCH_CHECK("dwconv3x3_pipe");
}

void launch_dw(const DwArgs& a, int prec, cudaStream_t s) {
  if (a.B > 65535) throw Error(-1, "depthwise: batch too large for one launch");
  if (a.stride != 1 && a.stride != 2) throw Error(-5, "depthwise: stride must be 1 or 2");
  int device = 0;
  cuda_check(cudaGetDevice(&device), "cudaGetDevice");
  const int es = prec == PREC_BF16 ? 2 : 4;
  int VL = (a.C * es) / 16;
  if (VL > 8) VL = 8;
  if (VL != 4 && VL != 8) throw Error(-1, "depthwise: channel count must give 64 or >=128 bytes per pixel");
#define HFR_DW_PIPE(T, S) \
  do { if (VL == 8) launch_dw_pipe_t<T, S, 8>(a, prec, device, s); else launch_dw_pipe_t<T, S, 4>(a, prec, device, s); } while (0)
  if (prec == PREC_BF16) { if (a.stride == 1) HFR_DW_PIPE(__nv_bfloat16, 1); else HFR_DW_PIPE(__nv_bfloat16, 2); }
  else                   { if (a.stride == 1) HFR_DW_PIPE(float, 1); else HFR_DW_PIPE(float, 2); }
#undef HFR_DW_PIPE
}

// ---------------------------------------------------------------------------------------------- GEMM (tcgen05)
// CTAS = 2: CTA-pair tiles (256 rows); p.num_units / p.num_m_blocks then count pairs, and the grid is two CTAs per unit.
template <typename T, int BLOCK_N, int EPI, int AMODE, int CTAS = 1>
static void launch_gemm_inst(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tD, const CUtensorMap& tR,
                             const CUtensorMap& tA0, const GemmParams& p, int device, cudaStream_t s) {
  using SM = GemmSmem<BLOCK_N, EPI, CTAS>;
  auto kern = gemm_tc_kernel<T, BLOCK_N, EPI, AMODE, CTAS>;
  static std::atomic<bool> configured[64];
  if (!configured[device].load()) {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal),
               "cudaFuncSetAttribute(gemm smem)");
    configured[device].store(true);
  }
  const int slots = device_sm_count(device) / CTAS;  // concurrently resident units
  int grid = (p.num_units < slots ? p.num_units : slots) * CTAS;
  if (grid < 1) return;
  launch_pdl_cluster(kern, CTAS, dim3(grid), dim3(384), (size_t)SM::kTotal, s, tA, tB, tD, tR, tA0, p);
  HFR_LAUNCH_CHECK("gemm_tc");
}

// CTA pairs halve the B (weight) bytes every SM pulls from L2 per k-block.  Measured per layer on ResNet-50 (B = 256):
// the implicit-GEMM convolutions (long K, A re-read per tap from L2) gain 8-12 %, 1x1 layers with K >= 768 gain 3-18 %,
// short-K 1x1 layers - HBM/epilogue-bound, where coupling two CTAs' epilogues only removes slack - lose 5-15 %.
// Used when the output width splits into 256- or 128-column pair tiles and there is at least one full wave of pairs.
static int pick_pair_block_n(int64_t M, int N, int K, int sms, bool im2col, bool multi_tap) {
  static const char* mode = getenv("HFR_PAIR");  // "0": never, "1": whenever the shape allows, unset: measured policy
  if (mode && mode[0] == '0') return 0;
  // (K = 768, the K-concatenated first block of stage 4: 87.6 us as single-CTA tiles, 71.6 us as pairs; K = 512: 39.5 vs 46)
  if (!(mode && mode[0] == '1') && !multi_tap && K < 768) return 0;
  const bool need_even_m_blocks = im2col;  // im2col base pixels past the last image are not loaded
  const int64_t mb = (M + 127) / 128;
  if (need_even_m_blocks && (mb & 1)) return 0;
  const int64_t pairs = (mb + 1) / 2;
  int bn = (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : 0;
  if (bn == 0) return 0;
  if (pairs * (N / bn) < sms / 2) return 0;
  if (bn == 256) {
    // wave quantisation: the launch lasts ceil(units / resident pairs) rounds of one tile each, a tile's time scales
    // with its width.  Few wide tiles can leave most of the last round idle (ResNet-50 stage 5: 98 tiles on 74 pairs =
    // 2 rounds at 66 %); take 128-column tiles when they shorten the launch by more than 10 %
    const int64_t slots = sms / 2;
    const int64_t u256 = pairs * (N / 256), u128 = pairs * (N / 128);
    const int64_t t256 = (u256 + slots - 1) / slots * 256, t128 = (u128 + slots - 1) / slots * 128;
    if (t128 * 10 < t256 * 9) bn = 128;
  }
  return bn;
}

static int pick_block_n(int64_t M, int N, int sms, int widest) {
  const int64_t mb = (M + 127) / 128;
  const int cands[3] = {256, 128, 64};
  static const int env_cap = getenv("HFR_BLOCK_N_MAX") ? atoi(getenv("HFR_BLOCK_N_MAX")) : 0;  // tuning experiments
  const int cap = env_cap ? env_cap : widest;
  for (int c : cands) {
    if (c > cap && c != 64) continue;
    if (c > N && c != 64) continue;
    if (N % c != 0 && c != 64) continue;
    if (mb * ((N + c - 1) / c) >= sms) return c;
  }
  return 64;
}

void gemm_tile_choice(int64_t M, int N, int K, int conv_taps, int sms, int* ctas, int* block_n) {
  const bool im2col = conv_taps > 0;
  if (const int pbn = pick_pair_block_n(M, N, K, sms, im2col, conv_taps > 1)) {
    *ctas = 2;
    *block_n = pbn;
    return;
  }
  *ctas = 1;
  *block_n = pick_block_n(M, N, sms, (!im2col || conv_taps == 1) ? 128 : 256);
}

template <typename T, int AMODE>
static void launch_gemm_store(const CUtensorMap& tA, const void* b, void* y, GemmParams p, int64_t M, int N, int K,
                              int prec, int device, cudaStream_t s, const void* a0 = nullptr, int K0 = 0) {
  // K is the TOTAL reduction length: K0 columns of the concatenated operand a0 (GemmParams::kb_split) + the main operand's
  CUtensorMap tD = make_tmap_2d(y, prec, (uint64_t)M, (uint64_t)N, 128);
  CUtensorMap tR = p.residual ? make_tmap_2d(p.residual, prec, (uint64_t)M, (uint64_t)N, 128) : tD;
  CUtensorMap tA0 = a0 ? make_tmap_2d(a0, prec, (uint64_t)M, (uint64_t)K0, 128) : tA;
  p.kb_split = a0 ? K0 / (128 / (int)elt_size(prec)) : 0;
  int ctas = 1, bn = 0;
  gemm_tile_choice(M, N, K, AMODE == AMODE_2D ? 0 : p.conv_kw * p.conv_kw, device_sm_count(device), &ctas, &bn);
  if (ctas == 2) {
    const int pbn = bn;
    CUtensorMap tB = make_tmap_2d(b, prec, (uint64_t)N, (uint64_t)K, (uint32_t)pbn / 2);  // each CTA loads half the rows
    p.num_m_blocks = (int)((M + 255) / 256);
    p.num_n_blocks = N / pbn;
    p.n_blocks_per_unit = 1;
    p.num_units = p.num_m_blocks * p.num_n_blocks;
    if (pbn == 256) launch_gemm_inst<T, 256, EPI_STORE, AMODE, 2>(tA, tB, tD, tR, tA0, p, device, s);
    else launch_gemm_inst<T, 128, EPI_STORE, AMODE, 2>(tA, tB, tD, tR, tA0, p, device, s);
    return;
  }
  // 1x1 layers are HBM / latency bound: 128-column tiles (twice the units, five 32 KB stages in flight instead of three
  // 48 KB ones) measured faster than 256-column ones on every ResNet-50 1x1 layer but two ties
  CUtensorMap tB = make_tmap_2d(b, prec, (uint64_t)N, (uint64_t)K, (uint32_t)bn);
  p.num_m_blocks = (int)((M + 127) / 128);
  p.num_n_blocks = (N + bn - 1) / bn;
  p.n_blocks_per_unit = 1;
  p.num_units = p.num_m_blocks * p.num_n_blocks;
  if (bn == 256) launch_gemm_inst<T, 256, EPI_STORE, AMODE>(tA, tB, tD, tR, tA0, p, device, s);
  else if (bn == 128) launch_gemm_inst<T, 128, EPI_STORE, AMODE>(tA, tB, tD, tR, tA0, p, device, s);
  else launch_gemm_inst<T, 64, EPI_STORE, AMODE>(tA, tB, tD, tR, tA0, p, device, s);
}

void launch_gemm(const GemmArgs& a, int prec, int device, cudaStream_t s) {
  if (a.M <= 0) return;
  if (prec == PREC_FP32) {
    dim3 grid((unsigned)((a.N + 63) / 64), (unsigned)((a.M + 63) / 64));
    sgemm_kernel<<<grid, 256, 0, s>>>((const float*)a.a, (const float*)a.b, a.bias, (const float*)a.residual,
                                       (float*)a.y, (int)a.M, a.N, a.K, a.act);
    HFR_LAUNCH_CHECK("sgemm");
    return;
  }
  const int es = (int)elt_size(prec);
  if ((a.K * es) % 16 || (a.N * es) % 128) throw Error(-1, "gemm: K rows must be multiples of 16 bytes, N rows of 128 bytes");
  if (a.M >= (1ll << 31)) throw Error(-1, "gemm: M too large");
  if (a.a0 && (a.K0 <= 0 || (a.K0 * es) % 128)) throw Error(-1, "gemm: the concatenated operand must be whole 128-byte k-blocks");
  const int K0 = a.a0 ? a.K0 : 0;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)a.M; p.N = a.N; p.K = K0 + a.K;
  p.bias = a.bias; p.residual = a.residual; p.act = a.act; p.round_tf32 = a.round_tf32;
  CUtensorMap tA = make_tmap_2d(a.a, prec, (uint64_t)a.M, (uint64_t)a.K, 128);
  if (prec == PREC_BF16) launch_gemm_store<__nv_bfloat16, AMODE_2D>(tA, a.b, a.y, p, a.M, a.N, p.K, prec, device, s, a.a0, K0);
  else launch_gemm_store<float, AMODE_2D>(tA, a.b, a.y, p, a.M, a.N, p.K, prec, device, s, a.a0, K0);
}

// ---------------------------------------------------------------------------------------------- fused GEMM pair
// HFR_SEAM=0: never fuse an 'increase' 1x1 convolution with the next block's 'reduce'; unset / 1: fuse every eligible pair.
// HFR_SEAM_BUFS=2|3|4 staging buffers per epilogue warpgroup (4: residual prefetched two chunks ahead) - A/B knob.
static int seam_mode() {
  const char* e = getenv("HFR_SEAM");   // read per call (host side, once per launch): tests flip it between models
  return e ? atoi(e) : 1;
}
static int pair_num_kb(int K, int prec) {   // 128-byte k-blocks of the first GEMM
  const int bk = 128 / (int)elt_size(prec);
  return (K + bk - 1) / bk;
}
bool gemm_pair_eligible(const GemmArgs& a, const GemmArgs& b, int prec, int device) {
  (void)device;
  if (seam_mode() == 0 || prec == PREC_FP32) return false;
  const int es = (int)elt_size(prec);
  if (a.M != b.M || b.a != a.y || a.N != b.K || a.M <= 0 || a.M >= (1ll << 31)) return false;
  if (b.residual != nullptr) return false;
  if ((a.K * es) % 16 || a.N % 128) return false;     // whole 128-column tiles of Y; TMA row pitch
  if (b.N != 64 && b.N != 128 && b.N != 256) return false;   // the second accumulator: one tile of N2 TMEM columns
  if (a.a0 && (a.K0 <= 0 || (a.K0 * es) % 128 || (a.K * es) % 128)) return false;
  if (pair_num_kb((a.a0 ? a.K0 : 0) + a.K, prec) > 4) return false;   // the unit's A rows stay resident in shared memory (<= 64 KB)
  return true;
}
// A buffers of a unit's resident rows: two (the next unit's rows arrive while this one computes) when they are small;
// HFR_SEAM_NA=1: a single buffer whenever that buys the fourth staging buffer (A/B knob)
static int pair_na(int num_kb, int nbuf) {
  static const int na_env = getenv("HFR_SEAM_NA") ? atoi(getenv("HFR_SEAM_NA")) : 0;
  if (num_kb > 2) return 1;
  if (na_env == 1 && PairSmem::stages_for(num_kb, 2, nbuf) < 4) return 1;
  return 2;
}
template <typename T, int N2, int NBUF, int PF>
static void launch_gemm_pair_inst(const GemmArgs& a, const GemmArgs& b, int prec, int device, cudaStream_t s) {
  auto kern = gemm_pair_kernel<T, N2, NBUF, PF>;
  static std::atomic<bool> configured[64];
  if (!configured[device].load()) {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::kMax),
               "cudaFuncSetAttribute(gemm pair smem)");
    configured[device].store(true);
  }
  PairParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)a.M; p.num_m_blocks = (int)((a.M + 127) / 128);
  const int K0 = a.a0 ? a.K0 : 0;
  p.N1 = a.N; p.K1 = K0 + a.K;
  p.kb_split = K0 / (128 / (int)elt_size(prec));
  p.bias1 = a.bias; p.residual = a.residual; p.act1 = a.act; p.round1 = a.round_tf32;
  p.bias2 = b.bias; p.act2 = b.act; p.round2 = b.round_tf32;
  const int num_kb = pair_num_kb(p.K1, prec);
  p.na = pair_na(num_kb, NBUF);
  p.stages = PairSmem::stages_for(num_kb, p.na, NBUF);
  const int grid = std::min(device_sm_count(device), p.num_m_blocks);
  CUtensorMap tA = make_tmap_2d(a.a, prec, (uint64_t)a.M, (uint64_t)a.K, 128);
  CUtensorMap tA0 = a.a0 ? make_tmap_2d(a.a0, prec, (uint64_t)a.M, (uint64_t)K0, 128) : tA;
  CUtensorMap tB1 = make_tmap_2d(a.b, prec, (uint64_t)a.N, (uint64_t)p.K1, 128);
  CUtensorMap tD1 = make_tmap_2d(a.y, prec, (uint64_t)a.M, (uint64_t)a.N, 128);
  CUtensorMap tR = a.residual ? make_tmap_2d(a.residual, prec, (uint64_t)a.M, (uint64_t)a.N, 128) : tD1;
  CUtensorMap tB2 = make_tmap_2d(b.b, prec, (uint64_t)b.N, (uint64_t)b.K, (uint32_t)std::min(N2, 128));
  CUtensorMap tD2 = make_tmap_2d(b.y, prec, (uint64_t)b.M, (uint64_t)b.N, 128);
  launch_pdl(kern, dim3(grid), dim3(384), (size_t)PairSmem::total(num_kb, p.na, NBUF, p.stages), s, tA, tB1, tD1, tR, tB2, tD2, tA0, p);
  HFR_LAUNCH_CHECK("gemm_pair");
}
// Staging buffers per epilogue warpgroup, residual prefetch distance, A buffers, ring slots and dynamic shared memory of
// the fused kernel for a first GEMM of `num_kb` 128-byte k-blocks (host logic, also behind hfr_debug_gemm_pair_config).
void gemm_pair_config(int num_kb, int* nbuf, int* pf, int* na, int* stages, int* smem_bytes) {
  static const int bufs_env = getenv("HFR_SEAM_BUFS") ? atoi(getenv("HFR_SEAM_BUFS")) : 0;
  int bufs = bufs_env ? bufs_env : 4;   // 4: residual chunks requested two ahead (stage 2: 192 -> 172 us per seam)
  // the weight ring keeps at least 4 slots: fewer staging buffers when the resident A rows are large
  while (bufs > 2 && PairSmem::stages_for(num_kb, pair_na(num_kb, bufs), bufs) < 4) --bufs;
  *nbuf = bufs >= 4 ? 4 : (bufs == 3 ? 3 : 2);
  *pf = *nbuf == 4 ? 2 : 1;
  *na = pair_na(num_kb, *nbuf);
  *stages = PairSmem::stages_for(num_kb, *na, *nbuf);
  *smem_bytes = PairSmem::total(num_kb, *na, *nbuf, *stages);
}
template <typename T, int N2>
static void launch_gemm_pair_n2(const GemmArgs& a, const GemmArgs& b, int prec, int device, cudaStream_t s) {
  int nbuf, pf, na, stages, smem;
  gemm_pair_config(pair_num_kb((a.a0 ? a.K0 : 0) + a.K, prec), &nbuf, &pf, &na, &stages, &smem);
  if (nbuf == 4) launch_gemm_pair_inst<T, N2, 4, 2>(a, b, prec, device, s);
  else if (nbuf == 3) launch_gemm_pair_inst<T, N2, 3, 1>(a, b, prec, device, s);
  else launch_gemm_pair_inst<T, N2, 2, 1>(a, b, prec, device, s);
}
template <typename T>
static void launch_gemm_pair_t(const GemmArgs& a, const GemmArgs& b, int prec, int device, cudaStream_t s) {
  if (b.N == 64) launch_gemm_pair_n2<T, 64>(a, b, prec, device, s);
  else if (b.N == 128) launch_gemm_pair_n2<T, 128>(a, b, prec, device, s);
  else launch_gemm_pair_n2<T, 256>(a, b, prec, device, s);
}
void launch_gemm_pair(const GemmArgs& a, const GemmArgs& b, int prec, int device, cudaStream_t s) {
  if (!gemm_pair_eligible(a, b, prec, device)) throw Error(-5, "gemm pair: the two layers do not form an eligible pair");
  if (prec == PREC_BF16) launch_gemm_pair_t<__nv_bfloat16>(a, b, prec, device, s);
  else launch_gemm_pair_t<float>(a, b, prec, device, s);
}

// ---------------------------------------------------------------------------------------------- implicit-GEMM conv
void launch_conv(const ConvArgs& a, int prec, int device, cudaStream_t s) {
  if (prec == PREC_FP32) throw Error(-5, "KxK convolutions run on the tensor-core path only (tf32 / bf16 precision)");
  static PFN_encodeIm2col enc = (PFN_encodeIm2col)driver_fn("cuTensorMapEncodeIm2col");
  const int es = (int)elt_size(prec);
  const int bk = 128 / es;
  if (a.cin % bk) throw Error(-1, "conv: input channels must be a multiple of the 128-byte K block");
  if (a.kh != a.kw) throw Error(-5, "conv: square kernels only");
  const int64_t M = (int64_t)a.B * a.Ho * a.Wo;
  const uint64_t dims[4] = {(uint64_t)a.cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
  const uint64_t strides[3] = {(uint64_t)a.cin * es, (uint64_t)a.W * a.cin * es, (uint64_t)a.H * a.W * a.cin * es};
  // bounding box of filter-window origins (fprop): lower = -pad_before, upper = pad_after - (k-1)*dil, with the
  // trailing pad chosen so that exactly Ho x Wo window positions exist at the traversal stride
  const int span = (a.kw - 1) * a.dil;
  const int pad_r = (a.Wo - 1) * a.stride + span + 1 - a.W - a.pad_l;
  const int pad_b = (a.Ho - 1) * a.stride + span + 1 - a.H - a.pad_t;
  int lower[2] = {-a.pad_l, -a.pad_t};
  int upper[2] = {pad_r - span, pad_b - span};
  uint32_t estr[4] = {1, (uint32_t)a.stride, (uint32_t)a.stride, 1};
  CUtensorMap tA;
  CUresult r = enc(&tA, tmap_dtype(prec), 4, const_cast<void*>(a.x), (const cuuint64_t*)dims,
                   (const cuuint64_t*)strides, lower, upper, (cuuint32_t)bk, 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-6, "cuTensorMapEncodeIm2col failed with CUresult " + std::to_string((int)r));
  {
    // Same workaround CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp): drivers up to 13.1 encode im2col maps of
    // tensors smaller than 128 KiB with a descriptor bit that must be cleared.
    int drv = 0;
    cuda_check(cudaDriverGetVersion(&drv), "cudaDriverGetVersion");
    const uint64_t tensor_bytes = (uint64_t)a.B * a.H * a.W * a.cin * es;
    if (drv <= 13010 && tensor_bytes < 131072) reinterpret_cast<uint64_t*>(&tA)[1] &= ~(1ull << 21);
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  if (a.a0 && (a.K0 <= 0 || a.K0 % bk)) throw Error(-1, "conv: the concatenated operand must be whole 128-byte k-blocks");
  const int K0 = a.a0 ? a.K0 : 0;
  p.M = (int)M; p.N = a.cout; p.K = K0 + a.kh * a.kw * a.cin;
  p.bias = a.bias; p.residual = a.residual; p.act = a.act; p.round_tf32 = a.round_tf32;
  p.conv_kw = a.kw; p.conv_cblocks = a.cin / bk; p.conv_wo = a.Wo; p.conv_ho = a.Ho;
  p.conv_stride = a.stride; p.conv_pad_w = a.pad_l; p.conv_pad_h = a.pad_t; p.conv_dil = a.dil;
  if (prec == PREC_BF16)
    launch_gemm_store<__nv_bfloat16, AMODE_IM2COL>(tA, a.w, a.y, p, M, a.cout, p.K, prec, device, s, a.a0, K0);
  else
    launch_gemm_store<float, AMODE_IM2COL>(tA, a.w, a.y, p, M, a.cout, p.K, prec, device, s, a.a0, K0);
}

// ---------------------------------------------------------------------------------------------- tensor-core stem
void launch_stem_tc(const StemTcArgs& a, int device, cudaStream_t s) {
  static PFN_encodeIm2col enc = (PFN_encodeIm2col)driver_fn("cuTensorMapEncodeIm2col");
  if (a.cout > 64 || a.cout % 32) throw Error(-5, "tensor-core stem: cout must be 32 or 64");
  const int Hs = a.Ho + a.ka - 1, Ws = a.Wo + a.kb - 1;
  {
    const long long total = (long long)a.B * Hs * Ws;
    launch_pdl(stem_s2d_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, a.x, (__nv_bfloat16*)a.scratch, a.B, a.H, a.W,
               Hs, Ws, a.pt2, a.pl2, a.use_window);
    HFR_LAUNCH_CHECK("stem_s2d");
  }
  if (a.use_window) {
    WinArgs w;
    w.x = a.scratch; w.w = a.w2; w.bias = a.bias; w.y = a.y;
    w.B = a.B; w.H = Hs; w.W = Ws; w.cin = 16; w.Ho = a.Ho; w.Wo = a.Wo; w.cout = a.cout;
    w.kh = a.ka; w.kw = a.kb; w.pad_t = 0; w.pad_l = 0; w.act = a.act;
    w.plane_major = 1;
    w.out_f32 = a.out_f32; w.passes = a.passes; w.round_tf32 = a.round_tf32;
    launch_conv_window(w, device, s);
    return;
  }
  const int64_t M = (int64_t)a.B * a.Ho * a.Wo;
  const uint64_t dims[4] = {16, (uint64_t)Ws, (uint64_t)Hs, (uint64_t)a.B};
  const uint64_t strides[3] = {32, (uint64_t)Ws * 32, (uint64_t)Hs * Ws * 32};
  int lower[2] = {0, 0};
  int upper[2] = {-(a.kb - 1), -(a.ka - 1)};
  uint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tA;
  CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.scratch, (const cuuint64_t*)dims,
                   (const cuuint64_t*)strides, lower, upper, 16, 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-6, "cuTensorMapEncodeIm2col(stem) failed with CUresult " + std::to_string((int)r));
  {
    int drv = 0;
    cuda_check(cudaDriverGetVersion(&drv), "cudaDriverGetVersion");
    if (drv <= 13010 && (uint64_t)a.B * Hs * Ws * 32 < 131072) reinterpret_cast<uint64_t*>(&tA)[1] &= ~(1ull << 21);
  }
  // weights [taps][cout][16]: 3-D map, box {16, 64, 4} (rows >= cout and taps >= ka*kb are zero-filled)
  const int taps = a.ka * a.kb;
  const uint64_t wd[3] = {16, (uint64_t)a.cout, (uint64_t)taps};
  const uint64_t ws[2] = {32, (uint64_t)a.cout * 32};
  const uint32_t wbox[3] = {16, 64, 4};
  CUtensorMap tB = make_tiled(a.w2, PREC_BF16, 3, wd, ws, wbox, CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tD = make_tmap_2d(a.y, PREC_BF16, (uint64_t)M, (uint64_t)a.cout, 128);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = a.cout; p.K = taps * 16;
  p.bias = a.bias; p.act = a.act;
  p.conv_kw = a.kb; p.conv_kh = a.ka; p.conv_taps = taps; p.conv_wo = a.Wo; p.conv_ho = a.Ho;
  p.conv_stride = 1; p.conv_dil = 1;
  p.num_m_blocks = (int)((M + 127) / 128);
  p.num_n_blocks = 1;
  p.n_blocks_per_unit = 1;
  p.num_units = p.num_m_blocks;
  launch_gemm_inst<__nv_bfloat16, 64, EPI_STORE, AMODE_STEM16>(tA, tB, tD, tD, tA, p, device, s);
}

// ---------------------------------------------------------------------------------------------- window conv
// shared-memory plan of the window kernel for `stages` ring slots and `passes` weight sweeps
static void window_smem(int cin, int kh, int kw, int* w_bytes, int* win_bytes, int* win_stride, int* total,
                        bool shifted = false, int stages = kWinStagesMin, int passes = 1) {
  const int planes = cin / 8;
  *w_bytes = kh * kw * planes * 64 * 16 * passes;
  *win_bytes = shifted ? kw * planes * (16 + kh - 1) * 128
                       : planes * (((16 + kh - 1) * (8 + kw - 1) * 16 + 127) / 128 * 128);  // planes at a 128-B pitch
  *win_stride = (*win_bytes + 1023) / 1024 * 1024;
  *total = 1024 + *w_bytes + stages * *win_stride + 4 * 16384 + 256;
}
bool conv_window_fits(int cin, int kh, int kw) {
  if (cin % 16) return false;
  int wb, wn, ws, tot;
  window_smem(cin, kh, kw, &wb, &wn, &ws, &tot);
  return tot <= 227 * 1024;
}
template <int TH, int TW, int KS, typename TOut = __nv_bfloat16, int PASSES = 1>
static void launch_window_inst(int grid, int total, int device, cudaStream_t s, const CUtensorMap& tX, const CUtensorMap& tW,
                               const CUtensorMap& tD, const WinParams& p, int w_bytes, int win_stride) {
  auto kern = conv_window_kernel<TH, TW, KS, TOut, PASSES>;
  static std::atomic<bool> configured[64];  // per instantiation
  if (!configured[device].load()) {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
               "cudaFuncSetAttribute(window smem)");
    configured[device].store(true);
  }
  launch_pdl(kern, dim3(grid), dim3(384), (size_t)total, s, tX, tW, tD, p, w_bytes, win_stride);
}
void launch_conv_window(const WinArgs& a, int device, cudaStream_t s) {
  if (a.cout > 64 || a.cout % 32 || !conv_window_fits(a.cin, a.kh, a.kw)) throw Error(-5, "window conv: unsupported shape");
  int w_bytes, win_bytes, win_stride, total;
  static const bool no_shifted = getenv("HFR_NO_SHIFTED") != nullptr;
  const bool shifted = a.plane_major && !no_shifted;
  if (a.passes != 1 && a.passes != 2) throw Error(-1, "window conv: passes must be 1 or 2");
  // (passes == 2: a second sweep of taps - the bf16 residual of the weights - is resident next to the first)
  static const int stage_cap = getenv("HFR_WIN_STAGES") ? atoi(getenv("HFR_WIN_STAGES")) : kWinStagesMax;
  int stages = kWinStagesMin;
  window_smem(a.cin, a.kh, a.kw, &w_bytes, &win_bytes, &win_stride, &total, shifted, stages, a.passes);
  if (total > 227 * 1024) throw Error(-5, "window conv: shared memory budget exceeded");
  while (stages < kWinStagesMax && stages < stage_cap) {   // as deep a ring as fits: the depth hides the window load latency
    int wb, wn, ws, tot;
    window_smem(a.cin, a.kh, a.kw, &wb, &wn, &ws, &tot, shifted, stages + 1, a.passes);
    if (tot > 227 * 1024) break;
    ++stages;
    total = tot;
  }
  const int planes = a.cin / 8;
  const int ww = 8 + a.kw - 1, wh = 16 + a.kh - 1;
  CUtensorMap tX;
  const int plane_px = wh * ww * 16;
  if (a.plane_major) {
    if (ww * 8 > 256) throw Error(-5, "window conv: window row too long for one TMA box");
    const uint64_t xd[4] = {(uint64_t)a.W * 8, (uint64_t)a.H, (uint64_t)planes, (uint64_t)a.B};
    const uint64_t xs[3] = {(uint64_t)a.W * 16, (uint64_t)a.H * a.W * 16, (uint64_t)planes * a.H * a.W * 16};
    const uint32_t xbox[4] = {(uint32_t)(shifted ? 8 : ww) * 8, (uint32_t)wh, (uint32_t)planes, 1};
    tX = make_tiled(a.x, PREC_BF16, 4, xd, xs, xbox, CU_TENSOR_MAP_SWIZZLE_NONE);
  } else {
    const uint64_t xd[4] = {(uint64_t)a.cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
    const uint64_t xs[3] = {(uint64_t)a.cin * 2, (uint64_t)a.W * a.cin * 2, (uint64_t)a.H * a.W * a.cin * 2};
    const uint32_t xbox[4] = {8, (uint32_t)ww, (uint32_t)wh, 1};
    tX = make_tiled(a.x, PREC_BF16, 4, xd, xs, xbox, CU_TENSOR_MAP_SWIZZLE_NONE);
  }
  const int chunks = a.kh * a.kw * planes * a.passes;
  const uint64_t wd[3] = {8, 64, (uint64_t)chunks};
  const uint64_t wst[2] = {16, 1024};
  const uint32_t wbox[3] = {8, 64, 2};
  CUtensorMap tW = make_tiled(a.w, PREC_BF16, 3, wd, wst, wbox, CU_TENSOR_MAP_SWIZZLE_NONE);
  const uint64_t yd[4] = {(uint64_t)a.cout, (uint64_t)a.Wo, (uint64_t)a.Ho, (uint64_t)a.B};
  const uint64_t yes = a.out_f32 ? 4 : 2;
  const uint64_t ys[3] = {(uint64_t)a.cout * yes, (uint64_t)a.Wo * a.cout * yes, (uint64_t)a.Ho * a.Wo * a.cout * yes};
  const uint32_t ybox[4] = {a.out_f32 ? 32u : 64u, 8, 16, 1};   // 128-byte rows either way
  CUtensorMap tD = make_tiled(a.y, a.out_f32 ? PREC_TF32 : PREC_BF16, 4, yd, ys, ybox, CU_TENSOR_MAP_SWIZZLE_128B);
  WinParams p;
  memset(&p, 0, sizeof(p));
  p.tiles_x = (a.Wo + 7) / 8;
  p.tiles_y = (a.Ho + 15) / 16;
  p.num_tiles = a.B * p.tiles_x * p.tiles_y;
  p.taps_h = a.kh; p.taps_w = a.kw; p.planes = planes; p.ww = ww; p.wh = wh;
  p.pad_l = a.pad_l; p.pad_t = a.pad_t; p.N = a.cout; p.bias = a.bias; p.act = a.act; p.w_chunks = chunks;
  p.plane_major = a.plane_major;
  p.plane_pitch = shifted ? wh * 128 : a.plane_major ? plane_px : (plane_px + 127) / 128 * 128;
  p.shifted = shifted;
  p.copy_pitch = planes * wh * 128;
  p.round_tf32 = a.round_tf32;
  p.stages = stages;
  const int grid = p.num_tiles < device_sm_count(device) ? p.num_tiles : device_sm_count(device);
  if (grid < 1) return;
  if (a.out_f32 || a.passes == 2) {  // experimental tf32-mode stem: fp32 output, two weight sweeps
    if (!(a.out_f32 && a.passes == 2)) throw Error(-5, "window conv: fp32 output comes with two weight sweeps");
    if (a.kh == 4 && a.kw == 4 && planes == 2)
      launch_window_inst<4, 4, 1, float, 2>(grid, total, device, s, tX, tW, tD, p, w_bytes, win_stride);
    else
      launch_window_inst<0, 0, 0, float, 2>(grid, total, device, s, tX, tW, tD, p, w_bytes, win_stride);
    HFR_LAUNCH_CHECK("conv_window_f32");
    return;
  }
  // fully unrolled MMA issue for the shapes the networks use: stem after space-to-depth (4x4 taps x 16 channels),
  // 3x3 over 64 and 32 channels
  if (a.kh == 4 && a.kw == 4 && planes == 2) launch_window_inst<4, 4, 1>(grid, total, device, s, tX, tW, tD, p, w_bytes, win_stride);
  else if (a.kh == 3 && a.kw == 3 && planes == 8) launch_window_inst<3, 3, 4>(grid, total, device, s, tX, tW, tD, p, w_bytes, win_stride);
  else if (a.kh == 3 && a.kw == 3 && planes == 4) launch_window_inst<3, 3, 2>(grid, total, device, s, tX, tW, tD, p, w_bytes, win_stride);
  else launch_window_inst<0, 0, 0>(grid, total, device, s, tX, tW, tD, p, w_bytes, win_stride);
  HFR_LAUNCH_CHECK("conv_window");
}

// ---------------------------------------------------------------------------------------------- simple kernels

void launch_maxpool(const PoolArgs& a, int prec, cudaStream_t s) {
  PoolParams p;
  p.B = a.B; p.H = a.H; p.W = a.W; p.C = a.C; p.Ho = a.Ho; p.Wo = a.Wo; p.k = a.k; p.stride = a.stride;
  p.pad_t = a.pad_t; p.pad_l = a.pad_l; p.explicit_zero = a.explicit_zero;
  const int vn = prec == PREC_BF16 ? 8 : 4;
  if (a.C % vn) throw Error(-1, "maxpool: channels must be a multiple of the 16-byte vector");
  if (prec == PREC_BF16 && a.k == 3 && a.stride == 2) {
    const long long pairs = (long long)a.B * a.Ho * ((a.Wo + 1) / 2) * (a.C / 8);
    launch_pdl(maxpool3x3s2_bf16_kernel, dim3(grid_for(pairs, 256)), dim3(256), 0, s, (const __nv_bfloat16*)a.x,
               (__nv_bfloat16*)a.y, p);
    HFR_LAUNCH_CHECK("maxpool3x3s2");
    return;
  }
  const long long total = (long long)a.B * a.Ho * a.Wo * (a.C / vn);
  if (prec == PREC_BF16)
    launch_pdl(maxpool_kernel<__nv_bfloat16>, dim3(grid_for(total, 256)), dim3(256), 0, s, (const __nv_bfloat16*)a.x, (__nv_bfloat16*)a.y, p);
  else
    launch_pdl(maxpool_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, s, (const float*)a.x, (float*)a.y, p);
  HFR_LAUNCH_CHECK("maxpool");
}

void launch_subsample(const void* x, void* y, int B, int H, int W, int C, int Ho, int Wo, int stride, int prec,
                      cudaStream_t s) {
  const int vn = prec == PREC_BF16 ? 8 : 4;
  const long long total = (long long)B * Ho * Wo * (C / vn);
  if (prec == PREC_BF16)
    launch_pdl(subsample_kernel<__nv_bfloat16>, dim3(grid_for(total, 256)), dim3(256), 0, s, (const __nv_bfloat16*)x,
               (__nv_bfloat16*)y, B, H, W, C, Ho, Wo, stride);
  else
    launch_pdl(subsample_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, s, (const float*)x, (float*)y, B, H, W, C,
               Ho, Wo, stride);
  HFR_LAUNCH_CHECK("subsample");
}

void launch_gap(const void* x, float* y, int B, int HW, int C, int prec, cudaStream_t s) {
  const int vn = prec == PREC_BF16 ? 8 : 4;
  dim3 grid((unsigned)B, (unsigned)((C / vn + 31) / 32));
  if (prec == PREC_BF16) launch_pdl(gap_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, (const __nv_bfloat16*)x, y, HW, C);
  else launch_pdl(gap_kernel<float>, grid, dim3(256), 0, s, (const float*)x, y, HW, C);
  HFR_LAUNCH_CHECK("gap");
}

template <int COLS>
static void launch_fc_t(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act,
                        cudaStream_t s) {
  const size_t smem = ((size_t)8 * K + 8 * 256 + 16) * sizeof(float);
  cuda_check(cudaFuncSetAttribute(fc_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
             "cudaFuncSetAttribute(fc smem)");
  dim3 grid((unsigned)((B + 7) / 8), (unsigned)((N + COLS - 1) / COLS));
  launch_pdl(fc_kernel<COLS>, grid, dim3(256), smem, s, x, w, bias, y, B, K, N, act);
  HFR_LAUNCH_CHECK("fc");
}
void launch_fc(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act,
               cudaStream_t s) {
  if (act == FC_SOFTMAX && N > 256) throw Error(-5, "dense+softmax: at most 256 classes");
  if (((size_t)8 * K + 8 * 256 + 16) * sizeof(float) > 200 * 1024) throw Error(-5, "dense: input dimension too large");
  if (act == FC_SOFTMAX) {   // the whole row must live in one CTA
    if (N <= 64) launch_fc_t<64>(x, w, bias, y, B, K, N, act, s);
    else if (N <= 128) launch_fc_t<128>(x, w, bias, y, B, K, N, act, s);
    else launch_fc_t<256>(x, w, bias, y, B, K, N, act, s);
  } else {
    launch_fc_t<32>(x, w, bias, y, B, K, N, act, s);   // 8 k-slices per column: short dependent-load chains
  }
}

static size_t heads_smem(int K, int n1) {
  const int slices = kHeadThreads / (n1 / 4);
  const size_t part = std::max((size_t)slices * kHeadRows * n1, (size_t)4 * kHeadRows * kHeadMaxCols);
  return ((size_t)kHeadRows * K + part + (size_t)kHeadRows * kHeadMaxHidden + (size_t)kHeadRows * kHeadMaxCols) * sizeof(float);
}
bool dense_heads_supported(int K, int n1, int n_heads, const int* n) {
  if (n1 <= 0 || n1 > kHeadMaxHidden || n1 % 4 || kHeadThreads % (n1 / 4) || n_heads < 0 || n_heads > kHeadMaxHeads) return false;
  int cols = 0;
  for (int h = 0; h < n_heads; ++h) cols += n[h];
  return cols <= kHeadMaxCols && heads_smem(K, n1) <= 200 * 1024;
}
void launch_dense_heads(const HeadsArgs& a, cudaStream_t s) {
  if (!dense_heads_supported(a.K, a.n1, a.n_heads, a.n)) throw Error(-5, "dense heads: unsupported shape");
  HeadsParams p;
  memset(&p, 0, sizeof(p));
  p.x = a.x; p.w1 = a.w1; p.b1 = a.b1; p.hidden = a.hidden; p.B = a.B; p.K = a.K; p.N1 = a.n1; p.act1 = a.act1;
  p.n_heads = a.n_heads;
  for (int h = 0; h < a.n_heads; ++h) {
    p.w[h] = a.w[h]; p.b[h] = a.b[h]; p.y[h] = a.y[h]; p.n[h] = a.n[h]; p.act[h] = a.act[h];
  }
  const size_t smem = heads_smem(a.K, a.n1);
  cuda_check(cudaFuncSetAttribute(dense_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
             "cudaFuncSetAttribute(dense heads smem)");
  launch_pdl(dense_heads_kernel, dim3((unsigned)((a.B + kHeadRows - 1) / kHeadRows)), dim3(kHeadThreads), smem, s, p);
  HFR_LAUNCH_CHECK("dense_heads");
}

void launch_crop_resize(const uint8_t* frames, int H, int W, const int* boxes, int n, uint8_t* out, int oh, int ow,
                        cudaStream_t s) {
  if (n <= 0) return;
  const long long total = (long long)n * oh * ow;
  crop_resize_u8_kernel<<<grid_for(total, 256), 256, 0, s>>>(frames, H, W, boxes, n, out, oh, ow);
  HFR_LAUNCH_CHECK("crop_resize_u8");
}

size_t resize_pil_table_ints(int n, int oh, int ow, int taps_h, int taps_v) {
  return (size_t)n * ((size_t)ow * (2 + taps_h) + (size_t)oh * (2 + taps_v));
}
void launch_resize_pil(const uint8_t* images, const long long* desc, int* tab, int n, uint8_t* out, int oh, int ow,
                       int taps_h, int taps_v, cudaStream_t s) {
  if (n <= 0) return;
  if (taps_v > kPilMaxTaps || taps_h > kPilMaxTaps) throw Error(-5, "PIL resize: more than 31x reduction is not supported");
  const size_t smem = (size_t)taps_v * ow * 3;
  if (smem > 200 * 1024) throw Error(-5, "PIL resize: output row cache exceeds shared memory");
  if (n > 65535) throw Error(-1, "PIL resize: at most 65535 images per call");
  if (smem > 48 * 1024)   // per device and cheap: set on every call rather than cached in a process-wide flag
    cuda_check(cudaFuncSetAttribute(resize_pil_bilinear_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
               "cudaFuncSetAttribute(pil resize smem)");
  pil_coeff_kernel<<<dim3((unsigned)((ow + oh + 127) / 128), (unsigned)n), 128, 0, s>>>(desc, tab, oh, ow, taps_h, taps_v);
  HFR_LAUNCH_CHECK("pil_coeff");
  resize_pil_bilinear_u8_kernel<<<dim3((unsigned)oh, (unsigned)n), 256, smem, s>>>(images, desc, tab, out, oh, ow, taps_h,
                                                                                  taps_v);
  HFR_LAUNCH_CHECK("resize_pil_bilinear_u8");
}

void launch_gemm_x3(const float* a, const float* b, const float* bias, float* y, int64_t M, int N, int K, int64_t ldy, int act,
                    int device, cudaStream_t s) {
  if (M <= 0 || N <= 0) return;
  if (K % 4 || ldy < N) throw Error(-1, "gemm x3: K must be a multiple of 4 and ldy >= N");
  const int Np = (N + 31) / 32 * 32;            // output rows of whole 128-byte lines
  const int64_t K3 = 3 * (int64_t)K;
  const int64_t rows_per_pass = std::max<int64_t>(128, (int64_t)(192u << 20) / (K3 * 4));   // <= 192 MB of split A rows
  const int64_t mp = std::min(M, rows_per_pass);
  float *a3 = nullptr, *b3 = nullptr, *yt = nullptr, *biasp = nullptr;
  cuda_check(cudaMallocAsync((void**)&a3, (size_t)mp * K3 * 4, s), "cudaMallocAsync(x3 A)");
  cuda_check(cudaMallocAsync((void**)&b3, (size_t)Np * K3 * 4, s), "cudaMallocAsync(x3 B)");
  cuda_check(cudaMallocAsync((void**)&yt, (size_t)mp * Np * 4, s), "cudaMallocAsync(x3 Y)");
  cuda_check(cudaMemsetAsync(b3, 0, (size_t)Np * K3 * 4, s), "cudaMemsetAsync");
  split_tf32_kernel<<<grid_for((long long)N * K, 256), 256, 0, s>>>(b, b3, (long long)N, K, 1);
  HFR_LAUNCH_CHECK("split_tf32");
  if (bias && Np != N) {   // the epilogue reads Np bias values
    cuda_check(cudaMallocAsync((void**)&biasp, (size_t)Np * 4, s), "cudaMallocAsync(x3 bias)");
    cuda_check(cudaMemsetAsync(biasp, 0, (size_t)Np * 4, s), "cudaMemsetAsync");
    cuda_check(cudaMemcpyAsync(biasp, bias, (size_t)N * 4, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync(bias)");
  }
  for (int64_t m0 = 0; m0 < M; m0 += mp) {
    const int64_t mm = std::min(mp, M - m0);
    split_tf32_kernel<<<grid_for((long long)mm * K, 256), 256, 0, s>>>(a + m0 * K, a3, (long long)mm, K, 0);
    HFR_LAUNCH_CHECK("split_tf32");
    GemmArgs g;
    g.a = a3; g.b = b3; g.bias = biasp ? biasp : bias; g.residual = nullptr; g.y = yt; g.M = mm; g.N = Np; g.K = (int)K3;
    g.act = act; g.round_tf32 = 0;
    launch_gemm(g, PREC_TF32, device, s);
    cuda_check(cudaMemcpy2DAsync(y + m0 * ldy, (size_t)ldy * 4, yt, (size_t)Np * 4, (size_t)N * 4, (size_t)mm,
                                 cudaMemcpyDeviceToDevice, s), "cudaMemcpy2DAsync(x3 out)");
  }
  if (biasp) cuda_check(cudaFreeAsync(biasp, s), "cudaFreeAsync");
  cuda_check(cudaFreeAsync(yt, s), "cudaFreeAsync");
  cuda_check(cudaFreeAsync(b3, s), "cudaFreeAsync");
  cuda_check(cudaFreeAsync(a3, s), "cudaFreeAsync");
}

void launch_pairwise_dist(const float* x, const float* y, int64_t n, int64_t m, int d, const float* year_x,
                          const float* born_x, const float* year_y, const float* born_y, float age_w, float* out,
                          int device, cudaStream_t s) {
  if (n <= 0 || m <= 0) return;
  if (d % 4) throw Error(-1, "pairwise distances: the dimension must be a multiple of 4");
  // cross terms G = X Y^T on the tensor cores (3xTF32), norms in fp32, then d2 = |x|^2 + |y|^2 - 2 G with the
  // cancelling elements recomputed directly (pairwise_post_kernel)
  const int64_t mpad = (m + 31) / 32 * 32;
  float *G = nullptr, *nx = nullptr, *ny = nullptr;
  cuda_check(cudaMallocAsync((void**)&G, (size_t)n * mpad * 4, s), "cudaMallocAsync(pairwise G)");
  cuda_check(cudaMallocAsync((void**)&nx, (size_t)n * 4, s), "cudaMallocAsync(norms)");
  launch_rows_prep(x, nullptr, nx, nullptr, n, d, s);
  if (y != x) {
    cuda_check(cudaMallocAsync((void**)&ny, (size_t)m * 4, s), "cudaMallocAsync(norms)");
    launch_rows_prep(y, nullptr, ny, nullptr, m, d, s);
  }
  launch_gemm_x3(x, y, nullptr, G, n, (int)m, d, mpad, 0, device, s);
  pairwise_post_kernel<<<grid_for((long long)n * m, 256), 256, 0, s>>>(x, y, G, (int)mpad, nx, ny ? ny : nx, (long long)n,
                                                                        (long long)m, d, year_x, born_x, year_y, born_y,
                                                                        age_w, x == y ? 1 : 0, out);
  HFR_LAUNCH_CHECK("pairwise_post");
  if (ny) cuda_check(cudaFreeAsync(ny, s), "cudaFreeAsync");
  cuda_check(cudaFreeAsync(nx, s), "cudaFreeAsync");
  cuda_check(cudaFreeAsync(G, s), "cudaFreeAsync");
}

void launch_age_post(const float* probs, float* age, int B, int N, cudaStream_t s) {
  age_post_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(probs, age, B, N);
  HFR_LAUNCH_CHECK("age_post");
}

void launch_l2norm(const float* x, float* y, int64_t n, int d, cudaStream_t s) {
  if (n <= 0) return;
  l2norm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(x, y, (long long)n, d);
  HFR_LAUNCH_CHECK("l2norm");
}

void launch_cast_to_f32(const void* x, float* y, int64_t n, int prec, cudaStream_t s) {
  if (prec == PREC_BF16) {
    cast_to_f32_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, s>>>((const __nv_bfloat16*)x, y, (long long)n);
    HFR_LAUNCH_CHECK("cast_to_f32");
  } else {
    cuda_check(cudaMemcpyAsync(y, x, (size_t)n * 4, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync");
  }
}
void launch_cast_from_f32(const float* x, void* y, int64_t n, int prec, cudaStream_t s) {
  if (prec == PREC_BF16)
    cast_f32_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, s>>>(x, (__nv_bfloat16*)y, (long long)n, 0);
  else
    cast_f32_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(x, (float*)y, (long long)n, prec == PREC_TF32);
  HFR_LAUNCH_CHECK("cast_from_f32");
}

// ---------------------------------------------------------------------------------------------- 1-NN
void launch_rows_prep(const float* x, void* xb, float* norms, float* max_norm, int64_t n, int d, cudaStream_t s) {
  if (n <= 0) return;
  if (d % 4) throw Error(-1, "1-NN: dimension must be a multiple of 4");
  rows_prep_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(x, (__nv_bfloat16*)xb, norms, max_norm, (long long)n, d);
  HFR_LAUNCH_CHECK("rows_prep");
}

void knn_plan(int64_t nq, int64_t n, int* splits, int* n_blocks_per_unit) {
  const int64_t nb = (n + 255) / 256;
  const int64_t mb = (nq + 127) / 128;
  int64_t sp = (1184 + mb - 1) / mb;           // enough units to balance 148 persistent CTAs
  const int64_t sp_min = (nb + 63) / 64;       // at most 64 n-blocks (16384 gallery rows, L2-sized) per unit
  if (sp < sp_min) sp = sp_min;
  if (sp > nb) sp = nb;
  if (sp < 1) sp = 1;
  const int64_t per = (nb + sp - 1) / sp;
  *n_blocks_per_unit = (int)per;
  *splits = (int)((nb + per - 1) / per);
}

void launch_knn_gemm(const KnnGemmArgs& a, int prec, int device, cudaStream_t s) {
  if (prec != PREC_BF16 && prec != PREC_TF32) throw Error(-1, "1-NN: precision must be tf32 or bf16");
  const int es = (int)elt_size(prec);
  if ((a.d * es) % 16) throw Error(-1, "1-NN: rows must be multiples of 16 bytes");
  if (a.nq >= (1ll << 31) || a.n >= (1ll << 31)) throw Error(-1, "1-NN: shard too large for 32-bit indices");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)a.nq; p.N = (int)a.n; p.K = a.d;
  p.num_m_blocks = (int)((a.nq + 127) / 128);
  p.num_n_blocks = (int)((a.n + 255) / 256);
  p.n_blocks_per_unit = a.n_blocks_per_unit;
  p.splits = a.splits;
  p.num_units = p.num_m_blocks * a.splits;
  p.gnorm = a.gnorm; p.part_score = a.part_score; p.part_idx = a.part_idx;
  CUtensorMap tA = make_tmap_2d(a.q, prec, (uint64_t)a.nq, (uint64_t)a.d, 128);
  // CTA pairs (256 query rows x 256 gallery rows per tile, each CTA stages half of the gallery rows): fewer shared-memory
  // and L2 bytes per flop; used when there is at least one wave of pair units
  static const char* pair_mode = getenv("HFR_PAIR");
  const int64_t pairs = (a.nq + 255) / 256;
  const bool pair = !(pair_mode && pair_mode[0] == '0') && pairs * a.splits >= device_sm_count(device) / 2;
  if (pair) {
    p.num_m_blocks = (int)pairs;
    p.num_units = p.num_m_blocks * a.splits;
  }
  CUtensorMap tB = make_tmap_2d(a.g, prec, (uint64_t)a.n, (uint64_t)a.d, pair ? 128 : 256);
#define HFR_KNN_LAUNCH(T, EPI)                                                                       \
  do {                                                                                               \
    if (pair) launch_gemm_inst<T, 256, EPI, AMODE_2D, 2>(tA, tB, tA, tA, tA, p, device, s);              \
    else launch_gemm_inst<T, 256, EPI, AMODE_2D>(tA, tB, tA, tA, tA, p, device, s);                      \
  } while (0)
  if (a.cand == 4) {
    if (prec == PREC_BF16) HFR_KNN_LAUNCH(__nv_bfloat16, EPI_KNN4); else HFR_KNN_LAUNCH(float, EPI_KNN4);
  } else {
    if (prec == PREC_BF16) HFR_KNN_LAUNCH(__nv_bfloat16, EPI_KNN); else HFR_KNN_LAUNCH(float, EPI_KNN);
  }
#undef HFR_KNN_LAUNCH
}

static void configure_knn_exact(int device) {
  static std::atomic<bool> configured[64];
  if (!configured[device].load()) {
    cuda_check(cudaFuncSetAttribute(knn_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KnnExactSmem)),
               "cudaFuncSetAttribute(knn exact smem)");
    configured[device].store(true);
  }
}
void launch_knn_finalize(const KnnFinalizeArgs& a, int device, cudaStream_t s) {
  if (a.nq <= 0) return;
  if (a.k < 1 || a.k > 4) throw Error(-1, "k-NN: k must be 1..4");
  KnnFinalizeParams p;
  p.q = a.q; p.g = a.g; p.part_score = a.part_score; p.part_idx = a.part_idx;
  p.nbuckets = a.splits * 2;
  p.nq = a.nq; p.d = a.d; p.row_offset = a.row_offset; p.k = a.k;
  // rounding bound of the approximate scores (knn.cuh): operands rounded to bf16 (u = 2^-9) or truncated to tf32 by the
  // MMA (u = 2^-10), fp32 accumulation of d exact products, fp32 norms, the final fma
  const double u = a.precision == PREC_BF16 ? 1.0 / 512 : 1.0 / 1024;
  const double c_acc = std::max((double)a.d / 4194304.0, 1.0 / 4096);
  p.c_dot = 2.0 * (2.0 * u + u * u + c_acc) + 1.0 / 8388608;
  p.c_norm = (double)(a.d + 4) / 16777216.0;
  p.gmax2 = a.gmax2;
  p.out = (Neighbor*)a.out; p.unc_list = a.unc_list; p.counters = a.counters; p.partial = a.partial;
  const unsigned grid = (unsigned)((a.nq + 7) / 8);
  if (a.cand == 2) knn_finalize_kernel<2, 4><<<grid, 256, 0, s>>>(p);
  else knn_finalize_kernel<4, 8><<<grid, 256, 0, s>>>(p);
  HFR_LAUNCH_CHECK("knn_finalize");
  if (a.partial) return;   // sharded gallery: certification and the exact pass follow the exchange
  // exact pass over the queries the bound could not certify (usually none: the kernel reads the count and returns)
  configure_knn_exact(device);
  knn_exact_kernel<<<(unsigned)(2 * device_sm_count(device)), 256, sizeof(KnnExactSmem), s>>>(
      a.q, a.g, (long long)a.n, a.d, (long long)a.row_offset, a.k, a.unc_list, a.counters, a.locks, (Neighbor*)a.out);
  HFR_LAUNCH_CHECK("knn_exact");
  knn_rescore_kernel<<<64, 256, 0, s>>>(a.q, a.g, a.d, (long long)a.row_offset, a.k, a.unc_list, a.counters, (Neighbor*)a.out);
  HFR_LAUNCH_CHECK("knn_rescore");
}

void launch_knn_merge_certify(const void* parts, int n_parts, int64_t nq, int k, void* out, int* unc_list, int* unc_count,
                              cudaStream_t s) {
  cuda_check(cudaMemsetAsync(unc_count, 0, 4, s), "cudaMemsetAsync");
  if (nq <= 0) return;
  knn_merge_certify_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>((const Neighbor*)parts, n_parts, (long long)nq, k,
                                                                        (Neighbor*)out, unc_list, unc_count);
  HFR_LAUNCH_CHECK("knn_merge_certify");
}
void launch_knn_exact_listed(const float* q, const float* g, int64_t n, int d, int64_t row_offset, int k, const int* unc_list,
                             const int* unc_count, int* locks, void* out, int device, cudaStream_t s) {
  configure_knn_exact(device);
  knn_clear_listed_kernel<<<64, 256, 0, s>>>(unc_list, unc_count, k, (Neighbor*)out);
  HFR_LAUNCH_CHECK("knn_clear_listed");
  knn_exact_kernel<<<(unsigned)(2 * device_sm_count(device)), 256, sizeof(KnnExactSmem), s>>>(
      q, g, (long long)n, d, (long long)row_offset, k, unc_list, unc_count, locks, (Neighbor*)out);
  HFR_LAUNCH_CHECK("knn_exact");
  knn_rescore_kernel<<<64, 256, 0, s>>>(q, g, d, (long long)row_offset, k, unc_list, unc_count, (Neighbor*)out);
  HFR_LAUNCH_CHECK("knn_rescore");
}
void launch_knn_merge_listed(const void* parts, int n_parts, int64_t nq, int k, const int* unc_list, const int* unc_count,
                             void* out, cudaStream_t s) {
  knn_merge_listed_kernel<<<64, 256, 0, s>>>((const Neighbor*)parts, n_parts, (long long)nq, k, unc_list, unc_count,
                                             (Neighbor*)out);
  HFR_LAUNCH_CHECK("knn_merge_listed");
}

void launch_knn_merge(const void* parts, int n_parts, int64_t nq, int k, void* out, cudaStream_t s) {
  if (nq <= 0) return;
  knn_merge_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>((const Neighbor*)parts, n_parts, (long long)nq, k,
                                                                 (Neighbor*)out);
  HFR_LAUNCH_CHECK("knn_merge");
}

}  // namespace hfr
