// Host-callable launchers for every kernel of the hot path (implemented in launch.cu).  All functions enqueue work on
// `stream` and throw hfr::Error on failure; none of them synchronises.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace hfr {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void cuda_check(cudaError_t e, const char* what);

enum { PREC_FP32 = 0, PREC_TF32 = 1, PREC_BF16 = 2 };
inline size_t elt_size(int prec) { return prec == PREC_BF16 ? 2 : 4; }

int device_sm_count(int device);
void use_device(int device);  // cudaSetDevice + sm_100 check
int64_t launch_count();
void count_launch();                         // for kernels launched outside launch.cu (mtcnn.cu)
void set_last_error(const std::string& m);   // thread-local message behind hfr_last_error() (defined in api.cu)

struct StemArgs {
  const void* x;  // [B,H,W,3] u8 or f32
  int in_u8;
  const float* w;     // [kh][kw][3][cout]
  const float* bias;  // [cout] or null
  void* y;            // [B,Ho,Wo,cout] of T
  int B, H, W, Ho, Wo, kh, kw, stride, pad_t, pad_l, cout;
  int flip;
  float scale, mean[3];
  int act, round_tf32;
};
void launch_stem(const StemArgs& a, int prec, cudaStream_t s);

// Stem on the tensor cores (bf16, uint8 input, stride 2, even H/W): space-to-depth staging + STEM16 implicit GEMM.
struct StemTcArgs {
  const uint8_t* x;     // [B,H,W,3]
  void* scratch;        // [B,Hs,Ws,16] bf16, Hs = Ho + ka - 1, Ws = Wo + kb - 1
  const void* w2;       // [ka*kb][cout][16] bf16 (built by stem_tc_weights)
  const float* bias;    // [cout] or null
  void* y;              // [B,Ho,Wo,cout] bf16
  int B, H, W, Ho, Wo, cout, ka, kb, pt2, pl2;   // pt2/pl2: even padding used by the staging kernel
  int act;
  int use_window;       // 1: w2 is in window layout [taps*2][64][8] (conv_window_kernel); 0: [taps][cout][16] (STEM16 im2col)
  // experimental tf32-mode stem (window kernel only): fp32 output, weights as hi + lo bf16 terms (w2 holds both sweeps)
  int out_f32 = 0, passes = 1, round_tf32 = 0;
};
void launch_stem_tc(const StemTcArgs& a, int device, cudaStream_t s);

// Window implicit-GEMM convolution (conv_window.cuh): stride 1, bf16, cout <= 64, weights [taps*cin/8][64][8] bf16.
struct WinArgs {
  const void* x;      // [B,H,W,cin] bf16
  const void* w;      // packed by pack_window_weights
  const float* bias;
  void* y;            // [B,Ho,Wo,cout] bf16
  int B, H, W, cin, Ho, Wo, cout, kh, kw, pad_t, pad_l, act;
  int plane_major;    // x is [B][cin/8][H][W][8] instead of NHWC
  int out_f32 = 0;    // y is fp32 (experimental)
  int passes = 1;     // 2: w holds a second sweep of taps with the weights' bf16 residual
  int round_tf32 = 0;
};
bool conv_window_fits(int cin, int kh, int kw);
void launch_conv_window(const WinArgs& a, int device, cudaStream_t s);

struct DwArgs {
  const void* x;      // [B,H,W,C] of T
  const float* w;     // [9][C]
  const float* bias;  // [C]
  void* y;            // [B,Ho,Wo,C]
  int B, H, W, C, Ho, Wo, stride, pad_t, pad_l, act, round_tf32;
};
void launch_dw(const DwArgs& a, int prec, cudaStream_t s);

struct GemmArgs {
  const void* a;         // [M,K] of T (row-major, K contiguous)
  const void* b;         // [N,K] of T
  const float* bias;     // [N] or null
  const void* residual;  // [M,N] of T or null
  void* y;               // [M,N] of T
  int64_t M;
  int N, K;
  int act, round_tf32;
  // K-concatenated operand: y = act([a0 | a] b^T + bias) with b = [N, K0 + K]; a0 is a plain [M, K0] matrix
  const void* a0 = nullptr;
  int K0 = 0;
};
void launch_gemm(const GemmArgs& a, int prec, int device, cudaStream_t s);
// Two dependent 1x1 convolutions in one launch (gemm_pair.cuh): second.a must be first.y, same M, second.N in
// {64, 128, 256}, no residual on the second; the second GEMM's A operand never leaves shared memory.
bool gemm_pair_eligible(const GemmArgs& first, const GemmArgs& second, int prec, int device);
void launch_gemm_pair(const GemmArgs& first, const GemmArgs& second, int prec, int device, cudaStream_t s);
void gemm_pair_config(int num_kb, int* nbuf, int* pf, int* na, int* stages, int* smem_bytes);
// fp32-accurate y[M, N] = act(a[M,K] b[N,K]^T + bias) on the tensor cores (3xTF32, kernels.cuh); ldy >= N is y's row
// pitch in floats.  Needs K % 4 == 0; scratch comes from the stream-ordered allocator.
void launch_gemm_x3(const float* a, const float* b, const float* bias, float* y, int64_t M, int N, int K, int64_t ldy, int act,
                    int device, cudaStream_t s);

struct ConvArgs {  // KxK convolution as implicit GEMM (NHWC), weights [cout][kh*kw][cin]
  const void* x;
  const void* w;
  const float* bias;
  const void* residual;
  void* y;
  int B, H, W, cin, Ho, Wo, cout, kh, kw, stride, pad_t, pad_l, dil;
  int act, round_tf32;
  // K-concatenated operand (see GemmArgs): a0 = plain [B*Ho*Wo, K0] matrix, w = [cout][K0 + kh*kw*cin]
  const void* a0 = nullptr;
  int K0 = 0;
};
void launch_conv(const ConvArgs& a, int prec, int device, cudaStream_t s);

struct PoolArgs {
  const void* x;
  void* y;
  int B, H, W, C, Ho, Wo, k, stride, pad_t, pad_l, explicit_zero;
};
void launch_maxpool(const PoolArgs& a, int prec, cudaStream_t s);
void launch_subsample(const void* x, void* y, int B, int H, int W, int C, int Ho, int Wo, int stride, int prec,
                      cudaStream_t s);
void launch_gap(const void* x, float* y, int B, int HW, int C, int prec, cudaStream_t s);
// act: 0 none, 1 relu, 3 sigmoid, 4 softmax
void launch_fc(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act, cudaStream_t s);
// The dense tail in one launch: hidden = act1(x W1 + b1) [B, n1], then up to 4 heads head_h = act_h(hidden W_h + b_h).
// Supported when dense_heads_supported() says so (n1 <= 256 and a multiple of 4, the heads' columns sum to <= 128).
struct HeadsArgs {
  const float* x; const float* w1; const float* b1; float* hidden;
  int B, K, n1, act1, n_heads;
  const float* w[4]; const float* b[4]; float* y[4];
  int n[4], act[4];
};
bool dense_heads_supported(int K, int n1, int n_heads, const int* n);
void launch_dense_heads(const HeadsArgs& a, cudaStream_t s);
// crop + cv2-exact bilinear resize of uint8 RGB boxes: boxes [n][5] = (frame, x1, y1, x2, y2), device int32
void launch_crop_resize(const uint8_t* frames, int H, int W, const int* boxes, int n, uint8_t* out, int oh, int ow,
                        cudaStream_t s);
// Pillow-exact bilinear resize of n uint8 RGB images of arbitrary size (desc: device int64 [n][4] = byte offset, height,
// width, row pitch); taps_h / taps_v = the largest tap counts over the images; tab: device scratch of
// resize_pil_table_ints(n, oh, ow, taps_h, taps_v) ints for the coefficient tables
size_t resize_pil_table_ints(int n, int oh, int ow, int taps_h, int taps_v);
void launch_resize_pil(const uint8_t* images, const long long* desc, int* tab, int n, uint8_t* out, int oh, int ow,
                       int taps_h, int taps_v, cudaStream_t s);
void launch_age_post(const float* probs, float* age, int B, int N, cudaStream_t s);
void launch_l2norm(const float* x, float* y, int64_t n, int d, cudaStream_t s);
void launch_cast_to_f32(const void* x, float* y, int64_t n, int prec, cudaStream_t s);
void launch_cast_from_f32(const float* x, void* y, int64_t n, int prec, cudaStream_t s);

// 1-NN
void launch_rows_prep(const float* x, void* x_bf16_or_null, float* norms_or_null, float* max_norm_or_null, int64_t n, int d,
                      cudaStream_t s);
struct KnnGemmArgs {
  const void* q;  // [nq, d] of T
  const void* g;  // [n, d] of T
  const float* gnorm;
  float* part_score;  // [nq][splits][2]
  int* part_idx;
  int64_t nq, n;
  int d, splits, n_blocks_per_unit;
  int cand = 2;  // candidates per (query, split, warpgroup): 2 (1-NN) or 4 (k-NN, k <= 4)
};
void knn_plan(int64_t nq, int64_t n, int* splits, int* n_blocks_per_unit);
// Tile shape the GEMM launcher picks for an M x N x K problem on a device with `sms` SMs (host logic only, testable
// without a GPU): *ctas = 1 or 2 (CTA pair), *block_n = 64 / 128 / 256.  conv_taps = kh * kw of an implicit-GEMM
// convolution (0: plain 2-D GEMM).
void gemm_tile_choice(int64_t M, int N, int K, int conv_taps, int sms, int* ctas, int* block_n);
void launch_knn_gemm(const KnnGemmArgs& a, int prec, int device, cudaStream_t s);
// merge of the candidate records + fp64 re-scoring + certification + (for uncertified queries) the exact fp64 pass;
// out: device hfr_neighbor [nq][k]
struct KnnFinalizeArgs {
  const float* q;
  const float* g;
  const float* part_score;
  const int* part_idx;
  int splits, cand;
  int64_t nq, n;
  int d;
  int64_t row_offset;
  int k, precision;
  const float* gmax2;
  void* out;
  int* unc_list;
  int* counters;
  int* locks;
  int partial;   // 1: out is [nq][k + 1] candidates + bound record, no certification / exact pass (sharded gallery)
};
void launch_knn_finalize(const KnnFinalizeArgs& a, int device, cudaStream_t s);
// sharded gallery (knn.cuh): global merge + certification of gathered partial records; exact pass over listed queries;
// merge of the gathered exact answers into the listed rows
void launch_knn_merge_certify(const void* parts, int n_parts, int64_t nq, int k, void* out, int* unc_list, int* unc_count,
                              cudaStream_t s);
void launch_knn_exact_listed(const float* q, const float* g, int64_t n, int d, int64_t row_offset, int k, const int* unc_list,
                             const int* unc_count, int* locks, void* out, int device, cudaStream_t s);
void launch_knn_merge_listed(const void* parts, int n_parts, int64_t nq, int k, const int* unc_list, const int* unc_count,
                             void* out, cudaStream_t s);
// pairwise euclidean distances (+ optional album age penalty), fp32 direct differences; y == x: the diagonal is forced
// to exact zeros
void launch_pairwise_dist(const float* x, const float* y, int64_t n, int64_t m, int d, const float* year_x,
                          const float* born_x, const float* year_y, const float* born_y, float age_w, float* out,
                          int device, cudaStream_t s);
// parts: device hfr_neighbor [n_parts][nq][k] -> out [nq][k]
void launch_knn_merge(const void* parts, int n_parts, int64_t nq, int k, void* out, cudaStream_t s);

}  // namespace hfr
