// CUDA-core kernels of the hot path (everything that is not a dense contraction): input staging + stem convolution,
// depthwise 3x3 (TMA halo staging), pooling, the dense heads, L2 normalisation and the 1-NN helper kernels.
// All activations are NHWC.  T is the activation storage type: __nv_bfloat16 (bf16 mode) or float (tf32 / fp32 modes).
#pragma once
#include "gemm_tc.cuh"
#include "knn.cuh"

namespace hfr {

template <typename T>
struct Vec16;  // 16-byte vector of T
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_RELU) x = fmaxf(x, 0.f);
  if (act == ACT_RELU6) x = fminf(fmaxf(x, 0.f), 6.f);
  return x;
}
template <typename T>
__device__ __forceinline__ float finish(float x, int act, int rtf32) {
  x = apply_act(x, act);
  if (sizeof(T) == 4 && rtf32) x = round_tf32(x);
  return x;
}

// ------------------------------------------------------------------------------------------------------------------
// Stem: direct KHxKW convolution over a 3-channel image with the reference's pre-processing fused into the load
// (facerec_test.py:96-110 / facial_analysis.py:102-107):  v[c] = in[flip ? 2-c : c] * scale - mean[c].
// Zero padding is applied after pre-processing, as in the TF graph.  One thread = one output pixel x 32 channels.
struct StemParams {
  int B, H, W, Ho, Wo, KH, KW, stride, pad_t, pad_l, Cout;
  int flip;
  float scale, mean[3];
  int act, round_tf32;
};

template <typename TIn, typename T>
__global__ void __launch_bounds__(128) stem_conv_kernel(const TIn* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, T* __restrict__ y,
                                                        const StemParams p) {
  extern __shared__ float s_w[];  // [KH*KW*3][32] slice of the weights for this channel group
  const int cg = blockIdx.y;      // channel group of 32
  const int ntap = p.KH * p.KW * 3;
  for (int i = threadIdx.x; i < ntap * 32; i += blockDim.x) {
    const int t = i >> 5, c = i & 31;
    s_w[i] = w[(size_t)t * p.Cout + cg * 32 + c];
  }
  pdl_launch_dependents();
  pdl_wait();
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npix = (long long)p.B * p.Ho * p.Wo;
  if (pix >= npix) return;
  const int ox = (int)(pix % p.Wo);
  const int oy = (int)((pix / p.Wo) % p.Ho);
  const int b = (int)(pix / ((long long)p.Wo * p.Ho));
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = bias ? bias[cg * 32 + c] : 0.f;
  const TIn* xb = x + (size_t)b * p.H * p.W * 3;
  for (int r = 0; r < p.KH; ++r) {
    const int iy = oy * p.stride - p.pad_t + r;
    if (iy < 0 || iy >= p.H) continue;
    for (int s = 0; s < p.KW; ++s) {
      const int ix = ox * p.stride - p.pad_l + s;
      if (ix < 0 || ix >= p.W) continue;
      const TIn* px = xb + ((size_t)iy * p.W + ix) * 3;
      const float i0 = (float)px[0], i1 = (float)px[1], i2 = (float)px[2];
      const float in3[3] = {p.flip ? i2 : i0, i1, p.flip ? i0 : i2};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = in3[c] * p.scale - p.mean[c];
        const float4* wr = reinterpret_cast<const float4*>(s_w + ((r * p.KW + s) * 3 + c) * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 ww = wr[q];
          acc[4 * q] = fmaf(v, ww.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
        }
      }
    }
  }
  T* yo = y + (size_t)pix * p.Cout + cg * 32;
  constexpr int VN = Vec16<T>::N;
#pragma unroll
  for (int q = 0; q < 32 / VN; ++q) {
    float v[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) v[e] = finish<T>(acc[q * VN + e], p.act, p.round_tf32);
    Vec16<T>::store(yo + q * VN, v);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Input staging ("next" row 1 of the scope table): crop boxes out of an RGB uint8 frame and resize every crop to the
// network input size, bit-exactly like cv2.resize(..., INTER_LINEAR) on uint8 - what age_gender_fun does per face
// (facial_analysis.py:95) after process_image cropped the box (facial_analysis.py:236-267).
// OpenCV's 8-bit path: 11-bit fixed-point weights  a = rint(w * 2048)  from  f = (float)((d + 0.5) * scale - 0.5),
// horizontal pass in int32, vertical pass  ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// Along x the weight is forced to (2048, 0) outside [0, w-1); along y only the row index is clamped.
struct ResizeCoef {
  int i0, i1, w0, w1;
};
__device__ __forceinline__ ResizeCoef resize_coef(int d, int src, int dst, bool is_x) {
  const double scale = (double)src / (double)dst;
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);   // no FMA contraction: must match the host
  int s = (int)floorf(f);
  f -= (float)s;
  if (is_x) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  ResizeCoef c;
  c.w0 = __float2int_rn(__fmul_rn(1.f - f, 2048.f));
  c.w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  c.i0 = min(max(s, 0), src - 1);
  c.i1 = min(max(s + 1, 0), src - 1);
  return c;
}
// frames [F,H,W,3] u8; boxes [n][5] int32 = (frame, x1, y1, x2, y2) with 0 <= x1 < x2 <= W; out [n,oh,ow,3] u8
__global__ void crop_resize_u8_kernel(const uint8_t* __restrict__ frames, int H, int W, const int* __restrict__ boxes,
                                      int n, uint8_t* __restrict__ out, int oh, int ow) {
  const long long total = (long long)n * oh * ow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % ow);
    const int dy = (int)((i / ow) % oh);
    const int b = (int)(i / ((long long)ow * oh));
    const int* bx = boxes + (size_t)b * 5;
    const int x1 = bx[1], y1 = bx[2], cw = bx[3] - bx[1], ch = bx[4] - bx[2];
    const uint8_t* src = frames + ((size_t)bx[0] * H + y1) * W * 3 + (size_t)x1 * 3;
    const ResizeCoef cx = resize_coef(dx, cw, ow, true), cy = resize_coef(dy, ch, oh, false);
    const uint8_t* r0 = src + (size_t)cy.i0 * W * 3;
    const uint8_t* r1 = src + (size_t)cy.i1 * W * 3;
    uint8_t* o = out + (size_t)i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = (int)r0[cx.i0 * 3 + c] * cx.w0 + (int)r0[cx.i1 * 3 + c] * cx.w1;
      const int h1 = (int)r1[cx.i0 * 3 + c] * cx.w0 + (int)r1[cx.i1 * 3 + c] * cx.w1;
      const int v = (((cy.w0 * (h0 >> 4)) >> 16) + ((cy.w1 * (h1 >> 4)) >> 16) + 2) >> 2;
      o[c] = (uint8_t)min(max(v, 0), 255);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// scipy.misc.imresize(img, size, interp='bilinear') (facerec_test.py:84,93) = Pillow's Image.resize(BILINEAR) on uint8:
// separable triangle filter widened by the scale factor when shrinking, coefficients normalised in double and quantised
// to 22-bit fixed point, horizontal pass first with its result rounded to uint8, then the vertical pass (Pillow
// libImaging/Resample.c, restated in oracle/resize.py).  Bit-exact: the double arithmetic uses the explicit _rn
// intrinsics so that no multiply-add is contracted differently from the host library.
// One CTA per (image, output row): the input rows that row needs are resampled horizontally into shared memory as
// uint8, then combined vertically.  desc[i] = {byte offset, height, width, row pitch}.
constexpr int kPilPrecisionBits = 32 - 8 - 2;
constexpr int kPilMaxTaps = 64;  // filter taps per output sample: 2 * ceil(max(scale, 1)) + 1  (scale <= 31)

struct PilAxis {
  double scale, support, ss;
};
__device__ __forceinline__ PilAxis pil_axis(int in_size, int out_size) {
  PilAxis a;
  a.scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fs = a.scale < 1.0 ? 1.0 : a.scale;
  a.support = fs;  // bilinear support 1.0 * filterscale
  a.ss = __ddiv_rn(1.0, fs);
  return a;
}
// window [x0, x0 + n) and fixed-point coefficients of output sample xx
__device__ __forceinline__ void pil_coeffs(const PilAxis& a, int in_size, int xx, int& x0, int& n, int* k) {
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), a.scale);
  x0 = (int)__dadd_rn(__dadd_rn(center, -a.support), 0.5);
  if (x0 < 0) x0 = 0;
  int x1 = (int)__dadd_rn(__dadd_rn(center, a.support), 0.5);
  if (x1 > in_size) x1 = in_size;
  n = x1 - x0;
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    double t = __dmul_rn(__dadd_rn(__dadd_rn((double)(x + x0), -center), 0.5), a.ss);
    if (t < 0.0) t = -t;
    const double w = t < 1.0 ? __dadd_rn(1.0, -t) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  for (int x = 0; x < n; ++x) {
    double t = __dmul_rn(__dadd_rn(__dadd_rn((double)(x + x0), -center), 0.5), a.ss);
    if (t < 0.0) t = -t;
    double w = t < 1.0 ? __dadd_rn(1.0, -t) : 0.0;
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    const double q = __dmul_rn(w, (double)(1 << kPilPrecisionBits));
    k[x] = w < 0.0 ? (int)__dadd_rn(-0.5, q) : (int)__dadd_rn(0.5, q);
  }
}
__device__ __forceinline__ int pil_clip8(int acc) {
  const int v = acc >> kPilPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Pass 1: coefficient tables, once per image and axis (the double-precision arithmetic lives only here).
// tab layout per image: [ow] x (x0, count, k[taps_h])  then  [oh] x (y0, count, k[taps_v])   (ints)
__global__ void pil_coeff_kernel(const long long* __restrict__ desc, int* __restrict__ tab, int oh, int ow, int taps_h,
                                 int taps_v) {
  const int img = blockIdx.y;
  const int H = (int)desc[img * 4 + 1], W = (int)desc[img * 4 + 2];
  const int per_img = ow * (2 + taps_h) + oh * (2 + taps_v);
  int* t_h = tab + (size_t)img * per_img;
  int* t_v = t_h + ow * (2 + taps_h);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ow + oh; i += gridDim.x * blockDim.x) {
    const bool horiz = i < ow;
    const int o = horiz ? i : i - ow;
    const int in_size = horiz ? W : H, out_size = horiz ? ow : oh;
    int* row = horiz ? t_h + (size_t)o * (2 + taps_h) : t_v + (size_t)o * (2 + taps_v);
    int k[kPilMaxTaps];
    int x0, n;
    if (in_size == out_size) {  // Pillow skips the pass: identity
      x0 = o;
      n = 1;
      k[0] = 1 << kPilPrecisionBits;
    } else {
      pil_coeffs(pil_axis(in_size, out_size), in_size, o, x0, n, k);
    }
    row[0] = x0;
    row[1] = n;
    for (int x = 0; x < n; ++x) row[2 + x] = k[x];
  }
}

// Pass 2: one CTA per (image, output row): the input rows that row needs are resampled horizontally into shared memory
// as uint8 (Pillow rounds between the passes), then combined vertically.
__global__ void __launch_bounds__(256) resize_pil_bilinear_u8_kernel(const uint8_t* __restrict__ images,
                                                                     const long long* __restrict__ desc,
                                                                     const int* __restrict__ tab,
                                                                     uint8_t* __restrict__ out, int oh, int ow,
                                                                     int taps_h, int taps_v) {
  extern __shared__ uint8_t pil_rows[];  // [taps_v][ow][3] horizontally resampled rows
  const int img = blockIdx.y, oy = blockIdx.x;
  const long long off = desc[img * 4 + 0];
  const long long pitch = desc[img * 4 + 3];
  const uint8_t* src = images + off;
  const int per_img = ow * (2 + taps_h) + oh * (2 + taps_v);
  const int* t_h = tab + (size_t)img * per_img;
  const int* t_v = t_h + ow * (2 + taps_h) + (size_t)oy * (2 + taps_v);
  const int y0 = t_v[0], ny = t_v[1];
  for (int ox = threadIdx.x; ox < ow; ox += blockDim.x) {
    const int* th = t_h + (size_t)ox * (2 + taps_h);
    const int x0 = th[0], nx = th[1];
    for (int j = 0; j < ny; ++j) {
      const uint8_t* row = src + (long long)(y0 + j) * pitch + (long long)x0 * 3;
      int a0 = 1 << (kPilPrecisionBits - 1), a1 = a0, a2 = a0;
      for (int x = 0; x < nx; ++x) {
        const int kx = th[2 + x];
        a0 += (int)row[3 * x] * kx;
        a1 += (int)row[3 * x + 1] * kx;
        a2 += (int)row[3 * x + 2] * kx;
      }
      uint8_t* d = pil_rows + ((size_t)j * ow + ox) * 3;
      d[0] = (uint8_t)pil_clip8(a0);
      d[1] = (uint8_t)pil_clip8(a1);
      d[2] = (uint8_t)pil_clip8(a2);
    }
  }
  __syncthreads();
  uint8_t* orow = out + (((size_t)img * oh + oy) * ow) * 3;
  for (int i = threadIdx.x; i < ow * 3; i += blockDim.x) {
    int acc = 1 << (kPilPrecisionBits - 1);
    for (int j = 0; j < ny; ++j) acc += (int)pil_rows[(size_t)j * ow * 3 + i] * t_v[2 + j];
    orow[i] = (uint8_t)pil_clip8(acc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Stem on the tensor cores, step 1: space-to-depth staging of the uint8 image for a stride-2 KHxKW convolution.
//   S[b, Y, X, (dy*2+dx)*3 + j] = u8[b, 2Y+dy-pt, 2X+dx-pl, j]   (0 outside the image)      j = raw channel 0..2
//   S[b, Y, X, 12..15]          = 1 inside the image, 0 in the padding ("valid" channels: they carry the folded
//                                 -mean term, so mean subtraction stays exact and padding stays exactly zero)
// pt/pl are even and H/W are even, so a 2x2 block is entirely inside or entirely outside the image.  uint8 values are
// exact in bf16.  The stride-2 conv then is a stride-1 (KH'/2)x(KW'/2) conv over S: 16 channels = one MMA K-step.
// plane_major: S is stored as [B][2 planes][Hs][Ws][8] (window kernel: whole halo window in one TMA box) instead of
// [B][Hs][Ws][16].
__global__ void __launch_bounds__(256) stem_s2d_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ s,
                                                       int B, int H, int W, int Hs, int Ws, int pt, int pl,
                                                       int plane_major) {
  const long long total = (long long)B * Hs * Ws;
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % Ws);
    const int Y = (int)((i / Ws) % Hs);
    const int b = (int)(i / ((long long)Ws * Hs));
    const int iy = 2 * Y - pt, ix = 2 * X - pl;
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
    if (iy >= 0 && iy + 1 < H && ix >= 0 && ix + 1 < W) {
      const uint8_t* p0 = x + (((size_t)b * H + iy) * W + ix) * 3;
      const uint8_t* p1 = p0 + (size_t)W * 3;
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        v[e] = (float)p0[e];
        v[6 + e] = (float)p1[e];
      }
      v[12] = v[13] = v[14] = v[15] = 1.f;
    }
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      w[e] = *reinterpret_cast<uint32_t*>(&t);
    }
    if (plane_major) {
      const size_t plane = (size_t)Hs * Ws;
      uint4* dst = reinterpret_cast<uint4*>(s) + ((size_t)b * 2 * plane + (size_t)Y * Ws + X);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[plane] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
      uint4* dst = reinterpret_cast<uint4*>(s + (size_t)i * 16);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Depthwise 3x3 (+ per-channel scale folded into the taps, bias, ReLU6), stride 1 or 2, TF SAME/explicit padding.
// HBM-bound: every input byte is fetched once by TMA into a shared-memory halo tile (out-of-range coordinates are
// zero-filled by the TMA unit, which is exactly the zero padding), every output byte is written once as 16-byte
// vectors.  CTA = 8x8 output pixels x (16 bytes * VL) channels; a thread owns one 16-byte channel vector of a
// vertical strip of 4 outputs and slides the 3-row window through registers.
// Persistent, software-pipelined kernel: a producer warp keeps `stages` halo windows in
// flight through a TMA ring while 16*VL compute threads drain them, so an SM holds stages x CTAs windows of loads
// outstanding instead of one per resident CTA, tiles cost no CTA launch, and the per-CTA constants (this CTA's channel
// block of weights and bias) are staged once: the launcher makes the grid a multiple of the channel-block count, so a
// CTA's tiles  t = blockIdx.x + k * gridDim.x  all share  cb = blockIdx.x % cblocks.  VL (16-byte vectors per pixel) is a
// template parameter: every shared-memory offset is an immediate.
struct DwPipeParams {
  int C, Ho, Wo, pad_t, pad_l, tiles_w, tiles_h, cblocks, act, round_tf32, stages, stage_bytes;
  int num_sp;  // spatial tiles: batch x tiles_h x tiles_w (each exists once per channel block)
};

template <typename T, int STRIDE, int VL>
__global__ void __launch_bounds__(16 * VL + 32) dwconv3x3_pipe_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                      const float* __restrict__ w,     // [9][C]
                                                                      const float* __restrict__ bias,  // [C]
                                                                      T* __restrict__ y, const DwPipeParams p) {
  constexpr int VN = Vec16<T>::N;
  constexpr int CBE = VL * VN;                 // channels per tile
  constexpr int NCOMP = 16 * VL;               // compute threads: 16 strips of 4 vertically adjacent outputs x VL vectors
  constexpr int TWI = 7 * STRIDE + 3, THI = 7 * STRIDE + 3;
  constexpr int NROWS = 3 * STRIDE + 3;        // input rows touched by 4 vertically adjacent outputs
  constexpr int kMaxStages = 6;
  extern __shared__ uint8_t dwp_smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages];
  __shared__ __align__(16) float w_s[9 * CBE];
  uint8_t* ring = dwp_smem_raw + ((128u - (smem_u32(dwp_smem_raw) & 127u)) & 127u);  // TMA destinations: 128-B aligned
  const uint32_t ring_u = smem_u32(ring);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8u * kMaxStages;
  const int warp = uniform_warp_idx();
  const int cb = (int)(blockIdx.x % (unsigned)p.cblocks);
  const int c0 = cb * CBE;
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8u * s, 1);
      mbar_init(empty0 + 8u * s, NCOMP / 32);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 9 * CBE; i += blockDim.x) w_s[i] = __ldg(w + (size_t)(i / CBE) * p.C + c0 + i % CBE);
  pdl_wait();       // loads of, and stores over, tensors of the predecessor only after it has completed
  __syncthreads();
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  // this CTA's tiles: channel block cb of spatial tiles sp0, sp0 + sp_step, ...  (32-bit arithmetic only: the loop
  // runs in every thread, a 64-bit division costs as much as a third of a tile's FMAs)
  const int sp0 = (int)(blockIdx.x / (unsigned)p.cblocks), sp_step = (int)(gridDim.x / (unsigned)p.cblocks);
  auto decode = [&](int sp, int& b, int& oy0, int& ox0) {
    b = sp / tiles_per_img;
    const int r = sp - b * tiles_per_img;
    const int ty = r / p.tiles_w;
    oy0 = ty * 8;
    ox0 = (r - ty * p.tiles_w) * 8;
  };
  if (warp == NCOMP / 32) {
    // ------------------------------------------------------------ producer
    int stage = 0;
    uint32_t phase = 0;
    for (int sp = sp0; sp < p.num_sp; sp += sp_step) {
      int b, oy0, ox0;
      decode(sp, b, oy0, ox0);
      mbar_wait(empty0 + 8u * stage, phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full0 + 8u * stage, (uint32_t)(TWI * THI * CBE * sizeof(T)));
        tma_load_4d(ring_u + (uint32_t)(stage * p.stage_bytes), &tmX, full0 + 8u * stage, c0, ox0 * STRIDE - p.pad_l,
                    oy0 * STRIDE - p.pad_t, b);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------ compute
    const int v = threadIdx.x % VL;
    const int strip = threadIdx.x / VL;     // 0..15
    const int lx = strip & 7;               // output column within the tile
    const int ly0 = (strip >> 3) * 4;       // first of this thread's 4 output rows
    const int c = c0 + v * VN;
    float bs[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) bs[e] = __ldg(bias + c + e);
    const float* wv = w_s + v * VN;
    const int in_off = ((ly0 * STRIDE) * TWI + lx * STRIDE) * CBE + v * VN;  // elements
    const size_t row_pitch = (size_t)p.Wo * p.C;
    int stage = 0;
    uint32_t phase = 0;
    for (int sp = sp0; sp < p.num_sp; sp += sp_step) {
      int b, oy0, ox0;
      decode(sp, b, oy0, ox0);
      mbar_wait(full0 + 8u * stage, phase);
      const T* xin = reinterpret_cast<const T*>(ring + (size_t)stage * p.stage_bytes) + in_off;
      float acc[4][VN];
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int e = 0; e < VN; ++e) acc[o][e] = bs[e];
      // one filter column at a time: only 3 taps x VN weights are live
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        float wk[3][VN];
#pragma unroll
        for (int kr = 0; kr < 3; ++kr) {
#pragma unroll
          for (int q = 0; q < VN / 4; ++q) {
            const float4 t4 = *reinterpret_cast<const float4*>(wv + (kr * 3 + s) * CBE + 4 * q);
            wk[kr][4 * q] = t4.x; wk[kr][4 * q + 1] = t4.y; wk[kr][4 * q + 2] = t4.z; wk[kr][4 * q + 3] = t4.w;
          }
        }
#pragma unroll
        for (int r = 0; r < NROWS; ++r) {
          float xv[VN];
          Vec16<T>::load(xin + (r * TWI + s) * CBE, xv);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int kr = r - o * STRIDE;  // filter row this input row hits for output o
            if (kr >= 0 && kr < 3) {
#pragma unroll
              for (int e = 0; e < VN; e += 2)   // packed fp32 FMA: same rounding as fmaf, half the issue slots
                ffma2(acc[o][e], acc[o][e + 1], xv[e], xv[e + 1], wk[kr][e], wk[kr][e + 1]);
            }
          }
        }
      }
      // the window is consumed: hand the slot back before the (long-latency) stores
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(empty0 + 8u * stage);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
      const int ox = ox0 + lx, oy = oy0 + ly0;
      if (ox < p.Wo) {
        T* dst = y + (((size_t)b * p.Ho + oy) * p.Wo + ox) * p.C + c;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          if (oy + o < p.Ho) {
            if constexpr (sizeof(T) == 2) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) w4[e] = pack_bf16x2_act(acc[o][2 * e], acc[o][2 * e + 1], p.act);
              *reinterpret_cast<uint4*>(dst + o * row_pitch) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            } else {
              float ov[VN];
#pragma unroll
              for (int e = 0; e < VN; ++e) ov[e] = finish<T>(acc[o][e], p.act, p.round_tf32);
              Vec16<T>::store(dst + o * row_pitch, ov);
            }
          }
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Max pooling (k x k, stride s) over NHWC, 16-byte channel vectors.  Out-of-range taps are skipped (TF SAME) unless
// explicit_zero is set (TF Pad followed by a VALID MaxPool: the pad pixels are real zeros and take part in the max).
struct PoolParams {
  int B, H, W, C, Ho, Wo, k, stride, pad_t, pad_l, explicit_zero;
};
template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, const PoolParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  const int cv = p.C / VN;
  const long long total = (long long)p.B * p.Ho * p.Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long t = i / cv;
    const int ox = (int)(t % p.Wo);
    t /= p.Wo;
    const int oy = (int)(t % p.Ho);
    const int b = (int)(t / p.Ho);
    float m[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) m[e] = -INFINITY;
    for (int r = 0; r < p.k; ++r) {
      const int iy = oy * p.stride - p.pad_t + r;
      for (int s = 0; s < p.k; ++s) {
        const int ix = ox * p.stride - p.pad_l + s;
        if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) {
          if (p.explicit_zero) {
#pragma unroll
            for (int e = 0; e < VN; ++e) m[e] = fmaxf(m[e], 0.f);
          }
          continue;
        }
        float xv[VN];
        Vec16<T>::load(x + (((size_t)b * p.H + iy) * p.W + ix) * p.C + v * VN, xv);
#pragma unroll
        for (int e = 0; e < VN; ++e) m[e] = fmaxf(m[e], xv[e]);
      }
    }
    Vec16<T>::store(y + (size_t)i * VN, m);
  }
}

// bf16 3x3 / stride-2 pooling (the ResNet pool1 shape): one thread produces two horizontally adjacent outputs from a
// 3-row x 5-column patch - column maxima first, the middle column shared - entirely in packed bf16x2 max instructions
// (max is exact in any format, so no widening).  5.6 loads and ~30 ALU instructions per output vector instead of 9 and ~150.
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.x) : "r"(a.x), "r"(b.x));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.y) : "r"(a.y), "r"(b.y));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.z) : "r"(a.z), "r"(b.z));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.w) : "r"(a.w), "r"(b.w));
  return r;
}
__global__ void maxpool3x3s2_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                         const PoolParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = p.C / 8;
  const int wp = (p.Wo + 1) / 2;  // output pairs per row
  const long long total = (long long)p.B * p.Ho * wp * cv;
  const uint32_t fill = p.explicit_zero ? 0u : 0xFF80FF80u;  // padding value: 0 (explicit Pad op) or -inf
  const uint4 pad4 = make_uint4(fill, fill, fill, fill);
  const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long t = i / cv;
    const int oxp = (int)(t % wp);
    t /= wp;
    const int oy = (int)(t % p.Ho);
    const int b = (int)(t / p.Ho);
    const int ix0 = oxp * 4 - p.pad_l, iy0 = oy * 2 - p.pad_t;
    uint4 col[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) col[c] = ninf;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = iy0 + r;
      const bool row_ok = iy >= 0 && iy < p.H;
      const __nv_bfloat16* row = x + (((size_t)b * p.H + (row_ok ? iy : 0)) * p.W) * p.C + v * 8;
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        const int ix = ix0 + c;
        uint4 val = pad4;
        if (row_ok && ix >= 0 && ix < p.W) val = *reinterpret_cast<const uint4*>(row + (size_t)ix * p.C);
        col[c] = bf16x8_max(col[c], val);
      }
    }
    const int ox = oxp * 2;
    __nv_bfloat16* out = y + (((size_t)b * p.Ho + oy) * p.Wo + ox) * p.C + v * 8;
    *reinterpret_cast<uint4*>(out) = bf16x8_max(bf16x8_max(col[0], col[1]), col[2]);
    if (ox + 1 < p.Wo) *reinterpret_cast<uint4*>(out + p.C) = bf16x8_max(bf16x8_max(col[2], col[3]), col[4]);
  }
}

// Spatial subsampling y[b,oy,ox,:] = x[b,oy*s,ox*s,:] - the gather in front of a strided 1x1 convolution.
template <typename T>
__global__ void subsample_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int Ho,
                                 int Wo, int s) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  const int cv = C / VN;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    long long t = i / cv;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const uint4 val = *reinterpret_cast<const uint4*>(x + (((size_t)b * H + oy * s) * W + ox * s) * C + v * VN);
    *reinterpret_cast<uint4*>(y + (size_t)i * VN) = val;
  }
}

// Global average pool [B, HW, C] (T) -> [B, C] fp32.  grid (B, C/(VN*32)); 32 channel-vector lanes x 8 pixel slices
// per CTA, slices combined through shared memory.  Matches TF Mean(axis=[1,2]) / AvgPool(HxW, VALID).
template <typename T>
__global__ void __launch_bounds__(256) gap_kernel(const T* __restrict__ x, float* __restrict__ y, int HW, int C) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  __shared__ float red[8][32][VN];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int cvec = blockIdx.y * 32 + lane;
  const int b = blockIdx.x;
  float acc[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) acc[e] = 0.f;
  if (cvec * VN < C) {
    for (int px = slice; px < HW; px += 8) {
      float xv[VN];
      Vec16<T>::load(x + ((size_t)b * HW + px) * C + cvec * VN, xv);
#pragma unroll
      for (int e = 0; e < VN; ++e) acc[e] += xv[e];
    }
  }
#pragma unroll
  for (int e = 0; e < VN; ++e) red[slice][lane][e] = acc[e];
  __syncthreads();
  if (slice == 0 && cvec * VN < C) {
    const float inv = 1.f / (float)HW;
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += red[q][lane][e];
      y[(size_t)b * C + cvec * VN + e] = s * inv;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Dense head: y[B,N] = act(x[B,K] * W[K,N] + bias), fp32.  act: 0 none, 1 relu, 3 sigmoid, 4 softmax (whole row in
// one CTA, N <= COLS).  CTA = 8 batch rows x COLS columns; the K dimension is split over KS = 256/COLS thread slices
// (independent, unrolled loads keep many weight rows in flight) and reduced through shared memory.
enum { FC_NONE = 0, FC_RELU = 1, FC_SIGMOID = 3, FC_SOFTMAX = 4 };
template <int COLS>
__global__ void __launch_bounds__(256) fc_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                 const float* __restrict__ bias, float* __restrict__ y, int B, int K,
                                                 int N, int act) {
  constexpr int KS = 256 / COLS;
  extern __shared__ float s_fc[];  // [8][K] inputs | [KS][8][COLS] partials | [8][2] softmax stats
  float* sx = s_fc;
  float* sp = s_fc + 8 * K;
  float* st = sp + KS * 8 * COLS;
  const int b0 = blockIdx.x * 8;
  const int col = threadIdx.x % COLS, ks = threadIdx.x / COLS;
  const int n = blockIdx.y * COLS + col;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < 8 * K; i += 256) {
    const int r = i / K, k = i - r * K;
    sx[i] = (b0 + r < B) ? x[(size_t)(b0 + r) * K + k] : 0.f;
  }
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  if (n < N) {
    const int kper = (K + KS - 1) / KS;
    const int k0 = ks * kper, k1 = min(K, k0 + kper);
    int k = k0;
    for (; k + 16 <= k1; k += 16) {  // 16 independent weight loads in flight per thread
      float ww[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) ww[u] = __ldg(w + (size_t)(k + u) * N + n);
#pragma unroll
      for (int u = 0; u < 16; ++u)
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = fmaf(sx[r * K + k + u], ww[u], acc[r]);
    }
    for (; k < k1; ++k) {
      const float ww = __ldg(w + (size_t)k * N + n);
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = fmaf(sx[r * K + k], ww, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) sp[(ks * 8 + r) * COLS + col] = acc[r];
  __syncthreads();
  // threads of slice 0 own the final values
  if (ks == 0) {
    const float bb = (bias && n < N) ? bias[n] : 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float v = bb;
#pragma unroll
      for (int q = 0; q < KS; ++q) v += sp[(q * 8 + r) * COLS + col];
      if (act == FC_RELU) v = fmaxf(v, 0.f);
      if (act == FC_SIGMOID) v = 1.f / (1.f + expf(-v));
      acc[r] = v;
      if (act == FC_SOFTMAX) sp[r * COLS + col] = (n < N) ? v : -INFINITY;  // slice 0's own slots: safe to overwrite
    }
  }
  if (act == FC_SOFTMAX) {
    __syncthreads();
    const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;  // warp r reduces row r
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, sp[wrp * COLS + j]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) sum += expf(sp[wrp * COLS + j] - mx);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      st[wrp * 2] = mx;
      st[wrp * 2 + 1] = sum;
    }
    __syncthreads();
    if (ks == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = expf(acc[r] - st[r * 2]) / st[r * 2 + 1];
    }
  }
  if (ks == 0 && n < N) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (b0 + r < B) y[(size_t)(b0 + r) * N + n] = acc[r];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The whole dense tail of the age/gender network in ONE launch (age_gender_train.py:91-98, facial_analysis.py:109):
//   hidden = act1(x[B,K] W1[K,N1] + b1)            'feats'  Dense(256, relu)
//   head_h = act_h(hidden W_h[N1,N_h] + b_h)        'age_pred' Dense(100, softmax) | 'gender_pred' Dense(1, sigmoid)
// CTA = 4 batch rows, 512 threads.  Layer 1: thread = (4 output columns as one float4 of W1's row, one of 512/(N1/4)
// K slices): every weight load is 16 bytes and 1/8 of the K loop long, the slices are reduced through shared memory.
// Layer 2: the heads' columns side by side (<= 128 in total), thread = (column, one of 4 K slices).  Softmax / sigmoid
// finish in shared memory.  Three dependent launches of ~40 us (latency-bound weight streaming) become one of ~15 us.
constexpr int kHeadRows = 4, kHeadThreads = 512, kHeadMaxHeads = 4, kHeadMaxCols = 128, kHeadMaxHidden = 256;
struct HeadsParams {
  const float* x;        // [B][K]
  const float* w1;       // [K][N1]
  const float* b1;       // [N1] or null
  float* hidden;         // [B][N1]
  int B, K, N1, act1;
  int n_heads;
  const float* w[kHeadMaxHeads];   // [N1][n[h]]
  const float* b[kHeadMaxHeads];
  float* y[kHeadMaxHeads];         // [B][n[h]]
  int n[kHeadMaxHeads], act[kHeadMaxHeads];
};

__global__ void __launch_bounds__(kHeadThreads) dense_heads_kernel(const HeadsParams p) {
  extern __shared__ __align__(16) float s_heads[];
  float* sx = s_heads;                                   // [4][K]
  float* part = sx + kHeadRows * p.K;                    // [slices][4][N1] (layer 1) / [4 slices][4][128] (layer 2)
  float* hid = part + (kHeadThreads / (p.N1 / 4)) * kHeadRows * p.N1;   // [4][N1]
  float* logit = hid + kHeadRows * kHeadMaxHidden;       // [4][128]
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * kHeadRows;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = tid; i < kHeadRows * p.K; i += kHeadThreads) {
    const int r = i / p.K, k = i - r * p.K;
    sx[i] = (b0 + r < p.B) ? p.x[(size_t)(b0 + r) * p.K + k] : 0.f;
  }
  __syncthreads();
  // ---- layer 1
  const int ncv = p.N1 / 4, slices = kHeadThreads / ncv;
  {
    const int cv = tid % ncv, ks = tid / ncv;
    float acc[kHeadRows][4];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    if (ks < slices) {
      const int kper = (p.K + slices - 1) / slices;
      const int k0 = ks * kper, k1 = min(p.K, k0 + kper);
      const float4* w4 = reinterpret_cast<const float4*>(p.w1) + cv;
      int k = k0;
      for (; k + 8 <= k1; k += 8) {   // 8 independent 16-byte weight loads in flight per thread
        float4 ww[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ww[u] = __ldg(w4 + (size_t)(k + u) * ncv);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r) {
            const float xv = sx[r * p.K + k + u];
            acc[r][0] = fmaf(xv, ww[u].x, acc[r][0]);
            acc[r][1] = fmaf(xv, ww[u].y, acc[r][1]);
            acc[r][2] = fmaf(xv, ww[u].z, acc[r][2]);
            acc[r][3] = fmaf(xv, ww[u].w, acc[r][3]);
          }
      }
      for (; k < k1; ++k) {
        const float4 ww = __ldg(w4 + (size_t)k * ncv);
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) {
          const float xv = sx[r * p.K + k];
          acc[r][0] = fmaf(xv, ww.x, acc[r][0]);
          acc[r][1] = fmaf(xv, ww.y, acc[r][1]);
          acc[r][2] = fmaf(xv, ww.z, acc[r][2]);
          acc[r][3] = fmaf(xv, ww.w, acc[r][3]);
        }
      }
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r)
        *reinterpret_cast<float4*>(&part[(ks * kHeadRows + r) * p.N1 + cv * 4]) =
            make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
  }
  __syncthreads();
  for (int i = tid; i < kHeadRows * p.N1; i += kHeadThreads) {
    const int r = i / p.N1, n = i - r * p.N1;
    float v = p.b1 ? p.b1[n] : 0.f;
    for (int q = 0; q < slices; ++q) v += part[(q * kHeadRows + r) * p.N1 + n];   // fixed order: deterministic
    if (p.act1 == FC_RELU) v = fmaxf(v, 0.f);
    hid[r * kHeadMaxHidden + n] = v;
    if (b0 + r < p.B) p.hidden[(size_t)(b0 + r) * p.N1 + n] = v;
  }
  __syncthreads();
  if (p.n_heads == 0) return;
  // ---- layer 2: all heads' columns side by side
  int off[kHeadMaxHeads + 1];
  off[0] = 0;
#pragma unroll
  for (int h = 0; h < kHeadMaxHeads; ++h) off[h + 1] = off[h] + (h < p.n_heads ? p.n[h] : 0);
  const int ncols = off[p.n_heads];
  {
    const int col = tid & (kHeadMaxCols - 1), ks = tid / kHeadMaxCols;   // 4 K slices
    float acc[kHeadRows] = {0.f, 0.f, 0.f, 0.f};
    if (col < ncols) {
      int h = 0;
      while (col >= off[h + 1]) ++h;
      const int c = col - off[h], nh = p.n[h];
      const float* wh = p.w[h] + c;
      const int kper = (p.N1 + 3) / 4;
      const int k0 = ks * kper, k1 = min(p.N1, k0 + kper);
      int k = k0;
      for (; k + 8 <= k1; k += 8) {
        float ww[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) ww[u] = __ldg(wh + (size_t)(k + u) * nh);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r) acc[r] = fmaf(hid[r * kHeadMaxHidden + k + u], ww[u], acc[r]);
      }
      for (; k < k1; ++k) {
        const float ww = __ldg(wh + (size_t)k * nh);
#pragma unroll
        for (int r = 0; r < kHeadRows; ++r) acc[r] = fmaf(hid[r * kHeadMaxHidden + k], ww, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) part[(ks * kHeadRows + r) * kHeadMaxCols + col] = acc[r];
  }
  __syncthreads();
  for (int i = tid; i < kHeadRows * kHeadMaxCols; i += kHeadThreads) {
    const int r = i / kHeadMaxCols, col = i - r * kHeadMaxCols;
    float v = -INFINITY;
    if (col < ncols) {
      int h = 0;
      while (col >= off[h + 1]) ++h;
      v = p.b[h] ? p.b[h][col - off[h]] : 0.f;
      for (int q = 0; q < 4; ++q) v += part[(q * kHeadRows + r) * kHeadMaxCols + col];
    }
    logit[i] = v;
  }
  __syncthreads();
  // ---- activations: warp (r, h) finishes row r of head h
  const int warp = tid >> 5, lane = tid & 31;
  for (int job = warp; job < kHeadRows * p.n_heads; job += kHeadThreads / 32) {
    const int r = job / p.n_heads, h = job - r * p.n_heads;
    if (b0 + r >= p.B) continue;
    const float* lg = logit + r * kHeadMaxCols + off[h];
    float* out = p.y[h] + (size_t)(b0 + r) * p.n[h];
    if (p.act[h] == FC_SOFTMAX) {
      float mx = -INFINITY;
      for (int j = lane; j < p.n[h]; j += 32) mx = fmaxf(mx, lg[j]);
      for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int j = lane; j < p.n[h]; j += 32) sum += expf(lg[j] - mx);
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      for (int j = lane; j < p.n[h]; j += 32) out[j] = expf(lg[j] - mx) / sum;
    } else {
      for (int j = lane; j < p.n[h]; j += 32) {
        float v = lg[j];
        if (p.act[h] == FC_RELU) v = fmaxf(v, 0.f);
        if (p.act[h] == FC_SIGMOID) v = 1.f / (1.f + expf(-v));
        out[j] = v;
      }
    }
  }
}

// facial_analysis.py:113-124: indices = argsort(p)[::-1][:2] (ties -> higher index first), age = 1 + sum(i*p_i)/sum(p_i).
__global__ void age_post_kernel(const float* __restrict__ probs, float* __restrict__ age, int B, int N) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float p1 = -INFINITY, p2 = -INFINITY;
  int i1 = -1, i2 = -1;
  auto push = [&](float v, int i) {
    if (i < 0) return;
    if (v > p1 || (v == p1 && i > i1)) {
      p2 = p1; i2 = i1; p1 = v; i1 = i;
    } else if (v > p2 || (v == p2 && i > i2)) {
      p2 = v; i2 = i;
    }
  };
  for (int j = lane; j < N; j += 32) push(probs[(size_t)row * N + j], j);
  for (int o = 16; o; o >>= 1) {
    const float q1 = __shfl_xor_sync(0xffffffffu, p1, o), q2 = __shfl_xor_sync(0xffffffffu, p2, o);
    const int j1 = __shfl_xor_sync(0xffffffffu, i1, o), j2 = __shfl_xor_sync(0xffffffffu, i2, o);
    push(q1, j1);
    push(q2, j2);
  }
  if (lane == 0) {
    const float s = p1 + p2;
    age[row] = 1.f + ((float)i1 * (p1 / s) + (float)i2 * (p2 / s));
  }
}

// sklearn.preprocessing.normalize(X, 'l2') (facerec_test.py:262,405): fp32 norms, zero rows stay zero.  Warp per row.
__global__ void l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int d) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int j = lane; j < d; j += 32) s = fmaf(xr[j], xr[j], s);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float nrm = sqrtf(s);
  for (int j = lane; j < d; j += 32) y[row * d + j] = nrm == 0.f ? xr[j] : xr[j] / nrm;
}

// ------------------------------------------------------------------------------------------------------------------
// fp32 CUDA-core GEMM (precision mode 0): C = act(A[M,K] * B[N,K]^T + bias (+ residual)).  64x64 tile, 4x4 per thread.
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                    const float* __restrict__ bias, const float* __restrict__ res,
                                                    float* __restrict__ C, int M, int N, int K, int act) {
  __shared__ float sa[16][64 + 4], sb[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      sa[k][r] = (m0 + r < M && k0 + k < K) ? A[(size_t)(m0 + r) * K + k0 + k] : 0.f;
      sb[k][r] = (n0 + r < N && k0 + k < K) ? Bm[(size_t)(n0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sb[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (res) v += res[(size_t)m * N + n];
      C[(size_t)m * N + n] = apply_act(v, act);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// fp32-accurate contractions on the tensor cores ("3xTF32"): x = hi + lo with hi = tf32(x), lo = tf32(x - hi) carries 21
// mantissa bits in two tf32 numbers; a . b ~ hiA.hiB + loA.hiB + hiA.loB (the dropped loA.loB term is 2^-22 relative).
// The three products are ONE tf32 GEMM over a K dimension three times as long: A' = [hi | lo | hi], B' = [hi | hi | lo].
// mode 0 writes the A' arrangement, mode 1 the B' arrangement.  out: [rows][3K].
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int K, int mode) {
  const long long total = rows * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / K;
    const int k = (int)(i - r * K);
    const float v = x[i];
    const float hi = round_tf32(v);
    const float lo = round_tf32(v - hi);
    float* o = out + r * 3 * K + k;
    o[0] = hi;
    o[K] = mode == 0 ? lo : hi;
    o[2 * K] = mode == 0 ? hi : lo;
  }
}

// Pairwise euclidean distance matrix for clustering (scope row 8f-3): out[i,j] = sqrt(sum_k (x[i,k] - y[j,k])^2), the
// expression the reference evaluates per pair (process_photos.py:46-48; facial_clustering_test.py:396-400 calls
// sklearn's pairwise_distances, which upcasts to fp64).  The cross terms come from the tensor cores (3xTF32 GEMM above,
// G = X Y^T with fp32-level accuracy); d2 = |x|^2 + |y|^2 - 2 G cancels when the rows are close, so every element whose
// d2 is below 5 % of |x|^2 + |y|^2 - where the expanded form has lost more than ~1.5 digits - is recomputed from the
// direct differences (duplicates, the diagonal, near-identical faces: a few elements per row).  Optional album
// penalty (process_photos.py:49-52): + w * (a_i - a_j)^2 / (a_i + a_j) with a = max(year_i, year_j) - born, clipped at 0.
__global__ void __launch_bounds__(256) pairwise_post_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ G, int ldg,
                                                            const float* __restrict__ nx, const float* __restrict__ ny,
                                                            long long n, long long m, int d,
                                                            const float* __restrict__ year_x, const float* __restrict__ born_x,
                                                            const float* __restrict__ year_y, const float* __restrict__ born_y,
                                                            float age_w, int zero_diag, float* __restrict__ out) {
  const long long total = n * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / m, j = e - i * m;
    const float a = nx[i], b = ny[j];
    // y == x: both (i, j) and (j, i) read the upper-triangle product, so the matrix is exactly symmetric
    const float gij = (zero_diag && i > j) ? G[j * ldg + i] : G[i * ldg + j];
    float d2 = a + b - 2.f * gij;
    if (d2 < 0.05f * (a + b)) {   // cancellation would cost more than ~2 of the 7 digits: direct differences (rare)
      const float4* xr = reinterpret_cast<const float4*>(x + i * d);
      const float4* yr = reinterpret_cast<const float4*>(y + j * d);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int k = 0; k < d / 4; ++k) {
        const float4 p = xr[k], q = yr[k];
        const float d0 = p.x - q.x, d1 = p.y - q.y, d2_ = p.z - q.z, d3 = p.w - q.w;
        s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2_, d2_, s2); s3 = fmaf(d3, d3, s3);
      }
      d2 = (s0 + s1) + (s2 + s3);
    }
    float v = sqrtf(fmaxf(d2, 0.f));
    if (zero_diag && i == j) v = 0.f;
    if (year_x != nullptr) {
      const float my = fmaxf(year_x[i], year_y[j]);
      const float ai = my - born_x[i], aj = my - born_y[j];
      v = fmaxf(v + age_w * ((ai - aj) * (ai - aj) / (ai + aj)), 0.f);
    }
    out[e] = v;
  }
}

template <typename T>
__global__ void cast_f32_kernel(const float* __restrict__ x, T* __restrict__ y, long long n, int rtf32) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    if (sizeof(T) == 4 && rtf32) v = round_tf32(v);
    y[i] = (T)v;
  }
}
template <typename T>
__global__ void cast_to_f32_kernel(const T* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = (float)x[i];
}

}  // namespace hfr
