// MTCNN face detector networks (scope row 8f-4): the three sess.run lambdas of FacialImageProcessing.load_mtcnn
// (facial_analysis.py:334-352) over the reference's mtcnn.pb - P-Net (fully convolutional, any image size), R-Net
// (24x24 crops), O-Net (48x48 crops).  The cascade around them (pyramid, NMS, box regression) is host code
// (hse_facerec_tf_b200/detection.py), as in the reference.
//
// The nets are tiny (3-128 channels, 1.2 M parameters in total, ~13 MMAC per O-Net candidate): fp32 CUDA-core kernels,
// latency-bound by construction; nothing here is GEMM-shaped enough for the tensor cores (K = 27 ... 576, N = 10 ... 128).
//   conv_small_kernel   direct KxK convolution, VALID/SAME, NHWC fp32, + bias + PReLU (Relu(x) - alpha*Relu(-x), the
//                       graph's Relu/Neg/Relu/Neg/Mul/Add chain); weights of a 16-output-channel block in shared memory
//   pool_small_kernel   max pooling with TF SAME/VALID semantics
//   fc_small_kernel     dense layer on the flattened NHWC map (+ bias + PReLU)
//   softmax_small_kernel  softmax over the channel dimension (the graph's Max/Sub/Exp/Sum/RealDiv chain)
// Layer geometry is read from the graph itself (kernel shapes, ksize/strides/padding attributes), only the node names
// are fixed - they are the names load_mtcnn binds.
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/hfr.h"
#include "graph.h"
#include "launch.h"

using namespace hfr;

namespace {

__global__ void __launch_bounds__(128) conv_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, const float* __restrict__ alpha,
                                                         float* __restrict__ y, int N, int H, int W, int Cin, int Ho, int Wo,
                                                         int Cout, int kh, int kw, int pad_t, int pad_l) {
  extern __shared__ float ws[];  // [kh*kw*Cin][16]
  const int co0 = blockIdx.y * 16;
  const int K = kh * kw * Cin;
  for (int i = threadIdx.x; i < K * 16; i += blockDim.x) {
    const int k = i >> 4, c = i & 15;
    ws[i] = (co0 + c < Cout) ? w[(size_t)k * Cout + co0 + c] : 0.f;
  }
  __syncthreads();
  const long long npix = (long long)N * Ho * Wo;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int ox = (int)(pix % Wo);
  const int oy = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((long long)Wo * Ho));
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = (bias && co0 + c < Cout) ? bias[co0 + c] : 0.f;
  for (int r = 0; r < kh; ++r) {
    const int iy = oy + r - pad_t;
    if (iy < 0 || iy >= H) continue;
    for (int s = 0; s < kw; ++s) {
      const int ix = ox + s - pad_l;
      if (ix < 0 || ix >= W) continue;
      const float* xp = x + (((size_t)n * H + iy) * W + ix) * Cin;
      const float* wp = ws + (size_t)((r * kw + s) * Cin) * 16;
      for (int c = 0; c < Cin; ++c) {
        const float xv = xp[c];
        const float4* w4 = reinterpret_cast<const float4*>(wp + c * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = w4[q];
          acc[4 * q] = fmaf(xv, ww.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(xv, ww.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, ww.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, ww.w, acc[4 * q + 3]);
        }
      }
    }
  }
  float* yp = y + (size_t)pix * Cout + co0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (co0 + c >= Cout) break;
    float v = acc[c];
    if (alpha) v = fmaxf(v, 0.f) - alpha[co0 + c] * fmaxf(-v, 0.f);
    yp[c] = v;
  }
}

__global__ void pool_small_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int Ho,
                                  int Wo, int k, int stride, int pad_t, int pad_l) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float m = -INFINITY;
    for (int r = 0; r < k; ++r) {
      const int iy = oy * stride - pad_t + r;
      if (iy < 0 || iy >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int ix = ox * stride - pad_l + s;
        if (ix < 0 || ix >= W) continue;
        m = fmaxf(m, x[(((size_t)n * H + iy) * W + ix) * C + c]);
      }
    }
    y[i] = m;
  }
}

__global__ void fc_small_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                const float* __restrict__ alpha, float* __restrict__ y, int rows, int K, int Nout) {
  const long long total = (long long)rows * Nout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Nout);
    const long long r = i / Nout;
    const float* xr = x + r * K;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = 0;
    for (; k + 4 <= K; k += 4) {   // fixed summation order: four interleaved partial sums
      a0 = fmaf(xr[k], w[(size_t)k * Nout + n], a0);
      a1 = fmaf(xr[k + 1], w[(size_t)(k + 1) * Nout + n], a1);
      a2 = fmaf(xr[k + 2], w[(size_t)(k + 2) * Nout + n], a2);
      a3 = fmaf(xr[k + 3], w[(size_t)(k + 3) * Nout + n], a3);
    }
    for (; k < K; ++k) a0 = fmaf(xr[k], w[(size_t)k * Nout + n], a0);
    float v = (a0 + a1) + (a2 + a3) + (bias ? bias[n] : 0.f);
    if (alpha) v = fmaxf(v, 0.f) - alpha[n] * fmaxf(-v, 0.f);
    y[i] = v;
  }
}

__global__ void softmax_small_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int C) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float* xr = x + r * C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, xr[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(xr[c] - m);
    for (int c = 0; c < C; ++c) y[r * C + c] = expf(xr[c] - m) / s;
  }
}

struct DevArr {
  float* p = nullptr;
  size_t n = 0;
};

struct Step {
  enum Kind { CONV, POOL, FC, SOFTMAX } kind = CONV;
  int kh = 1, kw = 1, cin = 0, cout = 0;
  bool same = false;      // TF padding attribute
  int pool_k = 0, pool_s = 0;
  DevArr w, b, alpha;     // alpha.p == nullptr: no PReLU
  int src = -1;           // index of the producing step (-1: the network input)
  int out_slot = -1;      // >= 0: this step's result is output `out_slot` of the net
};

struct Net {
  std::vector<Step> steps;
  int n_out = 0;
};

unsigned blocks_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  return (unsigned)std::max<long long>(1, std::min<long long>(g, 148 * 16));
}

}  // namespace

struct hfr_mtcnn {
  int device = 0;
  Net nets[3];
  std::vector<float*> owned;
  ~hfr_mtcnn() {
    for (float* p : owned) cudaFree(p);
  }
};

namespace {

const HTensor& const_of(const Graph& g, const std::string& name) {
  const GNode* n = g.find(name);
  if (!n) throw Error(HFR_ERR_NOT_FOUND, "mtcnn: node '" + name + "' not found in graph");
  const AttrVal* v = n->attr("value");
  if (n->op != "Const" || !v || v->kind != AttrVal::TENSOR) throw Error(HFR_ERR_FORMAT, "mtcnn: '" + name + "' is not a constant");
  return v->tensor;
}

DevArr upload_arr(hfr_mtcnn* m, const std::vector<float>& v) {
  DevArr d;
  d.n = v.size();
  cuda_check(cudaMalloc((void**)&d.p, std::max<size_t>(v.size(), 1) * 4), "cudaMalloc(mtcnn weights)");
  m->owned.push_back(d.p);
  if (!v.empty()) cuda_check(cudaMemcpy(d.p, v.data(), v.size() * 4, cudaMemcpyHostToDevice), "cudaMemcpy(mtcnn weights)");
  return d;
}

// conv / dense layer `scope/name` with an optional PReLU `scope/prelu`; geometry from the kernel's shape
Step make_layer(hfr_mtcnn* m, const Graph& g, const std::string& scope, const std::string& name, const std::string& prelu,
                bool dense, int src) {
  Step s;
  const HTensor& w = const_of(g, scope + "/" + name + "/weights");
  const HTensor& b = const_of(g, scope + "/" + name + "/biases");
  if (dense) {
    if (w.shape.size() != 2) throw Error(HFR_ERR_FORMAT, "mtcnn: dense kernel of " + name + " is not rank 2");
    s.kind = Step::FC;
    s.cin = (int)w.shape[0];
    s.cout = (int)w.shape[1];
  } else {
    if (w.shape.size() != 4) throw Error(HFR_ERR_FORMAT, "mtcnn: conv kernel of " + name + " is not rank 4");
    s.kind = Step::CONV;
    s.kh = (int)w.shape[0]; s.kw = (int)w.shape[1]; s.cin = (int)w.shape[2]; s.cout = (int)w.shape[3];
    const GNode* cn = g.find(scope + "/" + name + "/Conv2D");
    if (!cn) throw Error(HFR_ERR_NOT_FOUND, "mtcnn: node '" + scope + "/" + name + "/Conv2D' not found");
    const AttrVal* pd = cn->attr("padding");
    const AttrVal* st = cn->attr("strides");
    s.same = pd && pd->s == "SAME";
    if (st && st->shape.size() == 4 && (st->shape[1] != 1 || st->shape[2] != 1)) throw Error(HFR_ERR_UNSUPPORTED, "mtcnn: strided convolution");
    if ((size_t)s.kh * s.kw * s.cin * 16 * 4 > 48 * 1024) throw Error(HFR_ERR_UNSUPPORTED, "mtcnn: convolution window too large");
  }
  if ((int64_t)b.f.size() != s.cout) throw Error(HFR_ERR_FORMAT, "mtcnn: bias length of " + name);
  s.w = upload_arr(m, w.f);
  s.b = upload_arr(m, b.f);
  if (!prelu.empty()) {
    const HTensor& a = const_of(g, scope + "/" + prelu + "/alpha");
    if ((int64_t)a.f.size() != s.cout) throw Error(HFR_ERR_FORMAT, "mtcnn: PReLU alpha length of " + prelu);
    s.alpha = upload_arr(m, a.f);
  }
  s.src = src;
  return s;
}

Step make_pool(const Graph& g, const std::string& node, int src) {
  const GNode* n = g.find(node);
  if (!n || n->op != "MaxPool") throw Error(HFR_ERR_NOT_FOUND, "mtcnn: MaxPool '" + node + "' not found");
  const AttrVal* ks = n->attr("ksize");
  const AttrVal* st = n->attr("strides");
  const AttrVal* pd = n->attr("padding");
  if (!ks || ks->shape.size() != 4 || !st || st->shape.size() != 4 || ks->shape[1] != ks->shape[2] || st->shape[1] != st->shape[2])
    throw Error(HFR_ERR_FORMAT, "mtcnn: bad pooling attributes of " + node);
  Step s;
  s.kind = Step::POOL;
  s.pool_k = (int)ks->shape[1];
  s.pool_s = (int)st->shape[1];
  s.same = pd && pd->s == "SAME";
  s.src = src;
  return s;
}

Step make_softmax(int src, int slot) {
  Step s;
  s.kind = Step::SOFTMAX;
  s.src = src;
  s.out_slot = slot;
  return s;
}

void build(hfr_mtcnn* m, const Graph& g) {
  // the tensors load_mtcnn binds (facial_analysis.py:336-347): P-Net (conv4-2/BiasAdd, prob1), R-Net (conv5-2, prob1),
  // O-Net (conv6-2, conv6-3, prob1)
  {
    Net& n = m->nets[0];
    auto& s = n.steps;
    s.push_back(make_layer(m, g, "pnet", "conv1", "PReLU1", false, -1));
    s.push_back(make_pool(g, "pnet/pool1", 0));
    s.push_back(make_layer(m, g, "pnet", "conv2", "PReLU2", false, 1));
    s.push_back(make_layer(m, g, "pnet", "conv3", "PReLU3", false, 2));
    s.push_back(make_layer(m, g, "pnet", "conv4-2", "", false, 3));
    s.back().out_slot = 0;
    s.push_back(make_layer(m, g, "pnet", "conv4-1", "", false, 3));
    s.push_back(make_softmax(5, 1));
    n.n_out = 2;
  }
  {
    Net& n = m->nets[1];
    auto& s = n.steps;
    s.push_back(make_layer(m, g, "rnet", "conv1", "prelu1", false, -1));
    s.push_back(make_pool(g, "rnet/pool1", 0));
    s.push_back(make_layer(m, g, "rnet", "conv2", "prelu2", false, 1));
    s.push_back(make_pool(g, "rnet/pool2", 2));
    s.push_back(make_layer(m, g, "rnet", "conv3", "prelu3", false, 3));
    s.push_back(make_layer(m, g, "rnet", "conv4", "prelu4", true, 4));
    s.push_back(make_layer(m, g, "rnet", "conv5-2", "", true, 5));
    s.back().out_slot = 0;
    s.push_back(make_layer(m, g, "rnet", "conv5-1", "", true, 5));
    s.push_back(make_softmax(7, 1));
    n.n_out = 2;
  }
  {
    Net& n = m->nets[2];
    auto& s = n.steps;
    s.push_back(make_layer(m, g, "onet", "conv1", "prelu1", false, -1));
    s.push_back(make_pool(g, "onet/pool1", 0));
    s.push_back(make_layer(m, g, "onet", "conv2", "prelu2", false, 1));
    s.push_back(make_pool(g, "onet/pool2", 2));
    s.push_back(make_layer(m, g, "onet", "conv3", "prelu3", false, 3));
    s.push_back(make_pool(g, "onet/pool3", 4));
    s.push_back(make_layer(m, g, "onet", "conv4", "prelu4", false, 5));
    s.push_back(make_layer(m, g, "onet", "conv5", "prelu5", true, 6));
    s.push_back(make_layer(m, g, "onet", "conv6-2", "", true, 7));
    s.back().out_slot = 0;
    s.push_back(make_layer(m, g, "onet", "conv6-3", "", true, 7));
    s.back().out_slot = 1;
    s.push_back(make_layer(m, g, "onet", "conv6-1", "", true, 7));
    s.push_back(make_softmax(10, 2));
    n.n_out = 3;
  }
}

struct Shape {
  int h = 0, w = 0, c = 0;
  long long numel() const { return (long long)h * w * c; }
};

// output geometry of one step (TF SAME: out = ceil(in / s), pad_before = total / 2; VALID: floor((in - k) / s) + 1)
Shape step_shape(const Step& s, const Shape& in, int* pad_t, int* pad_l) {
  Shape o = in;
  *pad_t = *pad_l = 0;
  auto dim = [&](int size, int k, int stride, int* before) {
    if (s.same) {
      const int out = (size + stride - 1) / stride;
      const int total = std::max((out - 1) * stride + k - size, 0);
      *before = total / 2;
      return out;
    }
    *before = 0;
    return (size - k) / stride + 1;
  };
  switch (s.kind) {
    case Step::CONV:
      if (in.c != s.cin) throw Error(HFR_ERR_INVALID, "mtcnn: channel mismatch");
      o.h = dim(in.h, s.kh, 1, pad_t);
      o.w = dim(in.w, s.kw, 1, pad_l);
      o.c = s.cout;
      break;
    case Step::POOL:
      o.h = dim(in.h, s.pool_k, s.pool_s, pad_t);
      o.w = dim(in.w, s.pool_k, s.pool_s, pad_l);
      break;
    case Step::FC:
      if (in.numel() != s.cin) throw Error(HFR_ERR_INVALID, "mtcnn: the dense layer expects " + std::to_string(s.cin) +
                                                                 " inputs, the map has " + std::to_string(in.numel()));
      o.h = o.w = 1;
      o.c = s.cout;
      break;
    case Step::SOFTMAX:
      break;
  }
  if (o.h <= 0 || o.w <= 0) throw Error(HFR_ERR_INVALID, "mtcnn: input too small for this network");
  return o;
}

template <typename F>
int guarded_mt(F&& f) {
  try {
    f();
    return HFR_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {   // graphdef: malformed file
    set_last_error(e.what());
    return HFR_ERR_FORMAT;
  }
}

}  // namespace

extern "C" {

int hfr_mtcnn_load(const char* path, int device, hfr_mtcnn** out) {
  return guarded_mt([&] {
    if (!path || !out) throw Error(HFR_ERR_INVALID, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) throw Error(HFR_ERR_IO, std::string("cannot open '") + path + "'");
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + n);
    fclose(f);
    Graph g;
    parse_graphdef(data.data(), data.size(), &g);
    use_device(device);
    std::unique_ptr<hfr_mtcnn> m(new hfr_mtcnn());
    m->device = device;
    build(m.get(), g);
    *out = m.release();
  });
}

int hfr_mtcnn_out_shape(const hfr_mtcnn* m, int net, int h, int w, int slot, int* oh, int* ow, int* oc) {
  return guarded_mt([&] {
    if (!m || net < 0 || net > 2 || !oh || !ow || !oc) throw Error(HFR_ERR_INVALID, "bad argument");
    const Net& N = m->nets[net];
    std::vector<Shape> shp(N.steps.size());
    for (size_t i = 0; i < N.steps.size(); ++i) {
      int pt, pl;
      const Shape in = N.steps[i].src < 0 ? Shape{h, w, 3} : shp[(size_t)N.steps[i].src];
      shp[i] = step_shape(N.steps[i], in, &pt, &pl);
      if (N.steps[i].out_slot == slot) {
        *oh = shp[i].h; *ow = shp[i].w; *oc = shp[i].c;
        return;
      }
    }
    throw Error(HFR_ERR_INVALID, "no such output slot");
  });
}

int hfr_mtcnn_run(hfr_mtcnn* m, int net, const float* x, int n, int h, int w, float* out0, float* out1, float* out2,
                  void* stream) {
  return guarded_mt([&] {
    if (!m || net < 0 || net > 2 || !x || n <= 0 || h <= 0 || w <= 0) throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(m->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Net& N = m->nets[net];
    float* outs[3] = {out0, out1, out2};
    for (int i = 0; i < N.n_out; ++i)
      if (!outs[i]) throw Error(HFR_ERR_INVALID, "missing output buffer");
    std::vector<Shape> shp(N.steps.size());
    std::vector<float*> buf(N.steps.size(), nullptr);
    std::vector<float*> temps;
    for (size_t i = 0; i < N.steps.size(); ++i) {
      const Step& st = N.steps[i];
      const Shape in = st.src < 0 ? Shape{h, w, 3} : shp[(size_t)st.src];
      const float* src = st.src < 0 ? x : buf[(size_t)st.src];
      int pt, pl;
      shp[i] = step_shape(st, in, &pt, &pl);
      const long long numel = (long long)n * shp[i].numel();
      if (st.out_slot >= 0) {
        buf[i] = outs[st.out_slot];
      } else {
        cuda_check(cudaMallocAsync((void**)&buf[i], (size_t)numel * 4, s), "cudaMallocAsync(mtcnn activation)");
        temps.push_back(buf[i]);
      }
      switch (st.kind) {
        case Step::CONV: {
          const long long npix = (long long)n * shp[i].h * shp[i].w;
          dim3 grid((unsigned)((npix + 127) / 128), (unsigned)((st.cout + 15) / 16));
          conv_small_kernel<<<grid, 128, (size_t)st.kh * st.kw * st.cin * 16 * 4, s>>>(
              src, st.w.p, st.b.p, st.alpha.p, buf[i], n, in.h, in.w, in.c, shp[i].h, shp[i].w, st.cout, st.kh, st.kw, pt, pl);
          break;
        }
        case Step::POOL:
          pool_small_kernel<<<blocks_for(numel, 256), 256, 0, s>>>(src, buf[i], n, in.h, in.w, in.c, shp[i].h, shp[i].w,
                                                                 st.pool_k, st.pool_s, pt, pl);
          break;
        case Step::FC:
          fc_small_kernel<<<blocks_for(numel, 256), 256, 0, s>>>(src, st.w.p, st.b.p, st.alpha.p, buf[i], n, st.cin, st.cout);
          break;
        case Step::SOFTMAX:
          softmax_small_kernel<<<blocks_for((long long)n * in.h * in.w, 256), 256, 0, s>>>(src, buf[i], (long long)n * in.h * in.w,
                                                                                         in.c);
          break;
      }
      cuda_check(cudaGetLastError(), "launch mtcnn kernel");
      count_launch();
    }
    for (float* p : temps) cuda_check(cudaFreeAsync(p, s), "cudaFreeAsync");
  });
}

void hfr_mtcnn_free(hfr_mtcnn* m) { delete m; }

}  // extern "C"
