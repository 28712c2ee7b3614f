// C ABI (include/hfr.h): model handle (weights, activation arena, CUDA-graph cache), 1-NN handle, single-op entries.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <vector>

#include "../../include/hfr.h"
#include "graph.h"
#include "launch.h"

using namespace hfr;

static thread_local std::string g_err;
namespace hfr {
void set_last_error(const std::string& m) { g_err = m; }
}

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return HFR_OK;
  } catch (const Error& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    const std::string m = e.what();
    if (m.find("not found in graph") != std::string::npos) return HFR_ERR_NOT_FOUND;
    if (m.rfind("graphdef:", 0) == 0 || m.rfind("hdf5:", 0) == 0) return HFR_ERR_FORMAT;
    if (m.rfind("compile:", 0) == 0) return HFR_ERR_UNSUPPORTED;
    return HFR_ERR_INVALID;
  }
}

// ------------------------------------------------------------------------------------------------ helpers
static uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float f32_to_tf32(float f) {  // round to nearest even on the 13 dropped mantissa bits
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return f;
  u += 0x0FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

// Stem weights for the tensor-core path (see stem_s2d_kernel): [ka*kb taps][cout][16] bf16.  The colour flip, the
// input scale and the mean subtraction of the reference's pre-processing are folded in: data channels carry scale*W for
// the raw channel order, channels 12..14 carry -sum(mean*W) split into three bf16 terms (fp32-accurate constant).
struct StemTcGeom { int ka, kb, pt2, pl2; };
// lo_sweep: append a second [taps][cout][16] block holding the bf16 residual of every data weight (w = hi + lo carries
// 16 mantissa bits: the experimental tf32-mode stem issues both sweeps into one accumulator); its constant channels are 0.
static std::vector<uint16_t> stem_tc_weights(const std::vector<float>& w, int kh, int kw, int cout, int pad_t, int pad_l,
                                             int flip, float scale, const float mean[3], StemTcGeom* g,
                                             bool lo_sweep = false) {
  const int shy = pad_t & 1, shx = pad_l & 1;
  g->pt2 = pad_t + shy;
  g->pl2 = pad_l + shx;
  g->ka = (kh + shy + 1) / 2;
  g->kb = (kw + shx + 1) / 2;
  const size_t sweep = (size_t)g->ka * g->kb * cout * 16;
  std::vector<uint16_t> out(sweep * (lo_sweep ? 2 : 1), 0);
  for (int a = 0; a < g->ka; ++a)
    for (int b = 0; b < g->kb; ++b)
      for (int co = 0; co < cout; ++co) {
        uint16_t* dst = &out[(((size_t)a * g->kb + b) * cout + co) * 16];
        double cst = 0.0;
        for (int dy = 0; dy < 2; ++dy)
          for (int dx = 0; dx < 2; ++dx) {
            const int r = 2 * a + dy - shy, q = 2 * b + dx - shx;
            if (r < 0 || r >= kh || q < 0 || q >= kw) continue;
            for (int j = 0; j < 3; ++j) {
              const int c = flip ? 2 - j : j;
              const float ww = w[(((size_t)r * kw + q) * 3 + c) * cout + co];
              const uint16_t hi_w = f32_to_bf16(scale * ww);
              dst[(dy * 2 + dx) * 3 + j] = hi_w;
              if (lo_sweep) {
                uint32_t u = (uint32_t)hi_w << 16;
                float hf;
                memcpy(&hf, &u, 4);
                dst[sweep + (dy * 2 + dx) * 3 + j] = f32_to_bf16(scale * ww - hf);
              }
              cst -= (double)mean[c] * ww;
            }
          }
        auto bf2f = [](uint16_t h) {
          uint32_t u = (uint32_t)h << 16;
          float f;
          memcpy(&f, &u, 4);
          return f;
        };
        const uint16_t hi = f32_to_bf16((float)cst);
        const uint16_t mid = f32_to_bf16((float)(cst - bf2f(hi)));
        const uint16_t lo = f32_to_bf16((float)(cst - bf2f(hi) - bf2f(mid)));
        dst[12] = hi;
        dst[13] = mid;
        dst[14] = lo;
      }
  return out;
}

// [taps][cout][16] (STEM16 layout) -> window layout [(tap*2 + plane)][64 rows][8] (rows >= cout are zero)
static std::vector<uint16_t> stem_window_layout(const std::vector<uint16_t>& w2, int taps, int cout) {
  std::vector<uint16_t> out((size_t)taps * 2 * 64 * 8, 0);
  for (int t = 0; t < taps; ++t)
    for (int co = 0; co < cout; ++co)
      for (int c = 0; c < 16; ++c)
        out[(((size_t)t * 2 + c / 8) * 64 + co) * 8 + c % 8] = w2[((size_t)t * cout + co) * 16 + c];
  return out;
}
// conv weights [cout][taps][cin] fp32 -> window layout [(tap*planes + plane)][64 rows][8] bf16
static std::vector<uint16_t> conv_window_layout(const float* w, int cout, int taps, int cin) {
  const int planes = cin / 8;
  std::vector<uint16_t> out((size_t)taps * planes * 64 * 8, 0);
  for (int co = 0; co < cout; ++co)
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < cin; ++c)
        out[(((size_t)t * planes + c / 8) * 64 + co) * 8 + c % 8] = f32_to_bf16(w[((size_t)co * taps + t) * cin + c]);
  return out;
}
static bool window_enabled() { return getenv("HFR_NO_WINDOW") == nullptr; }

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cuda_check(cudaMalloc(&p, n), "cudaMalloc");
    bytes = n;
  }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

static void* upload(const void* host, size_t bytes) {
  void* d = nullptr;
  cuda_check(cudaMalloc(&d, bytes ? bytes : 4), "cudaMalloc(weights)");
  if (bytes) cuda_check(cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice), "cudaMemcpy(weights)");
  return d;
}

// ------------------------------------------------------------------------------------------------ model
struct LayerDev {
  void* w = nullptr;      // kernel-ready weights
  void* w_win = nullptr;  // conv_window_kernel layout, when the layer qualifies
  float* bias = nullptr;
  // K-concatenation candidate (kcat_candidate): [cout][K_prev + K] = the previous layer's weights and this layer's side
  // by side, and the sum of the two biases
  void* w_cat = nullptr;
  float* bias_cat = nullptr;
};

struct GraphKey {
  const void* x;
  int in_dtype, batch, flags;
  std::vector<void*> outs;
  bool operator<(const GraphKey& o) const {
    if (x != o.x) return x < o.x;
    if (in_dtype != o.in_dtype) return in_dtype < o.in_dtype;
    if (batch != o.batch) return batch < o.batch;
    if (flags != o.flags) return flags < o.flags;
    return outs < o.outs;
  }
};

struct hfr_model {
  Plan plan;
  int device = -1, precision = HFR_BF16;
  std::vector<LayerDev> dev;
  // activation arena: per-image byte offsets of every value (multiplied by the batch at run time)
  std::vector<size_t> val_off, val_bytes;
  std::vector<int> val_release;   // layer after which a value's arena slot is free again (-1: never / not materialised)
  size_t per_image_bytes = 0;
  bool keep_all = false;
  bool stem_force_direct = getenv("HFR_STEM_DIRECT") != nullptr;  // debugging: CUDA-core stem in every mode
  int sub_batch = getenv("HFR_SUB_BATCH") ? atoi(getenv("HFR_SUB_BATCH")) : 0;
  bool stem_window = window_enabled();
  // tf32 mode's stem on the tensor cores (measured: ResNet-50 31.2k -> 40.1k img/s); HFR_TF32_TC_STEM=0: CUDA-core stem
  bool tf32_tc_stem = !(getenv("HFR_TF32_TC_STEM") && getenv("HFR_TF32_TC_STEM")[0] == '0');
  DevBuf arena;
  int last_batch = 0;
  std::map<GraphKey, cudaGraphExec_t> graphs;
  // per-layer timing (eager mode only): event pairs around every layer launch
  bool timing = false;
  std::vector<cudaEvent_t> ev;           // 2 per layer, re-used every step
  std::vector<double> layer_ms;
  int timed_steps = 0;
  // tensor-core stem: staged space-to-depth input + per-preprocessing-flags weights
  DevBuf stem_scratch;
  std::map<int, std::pair<void*, StemTcGeom>> stem_w2;
  // host-buffer path staging
  DevBuf stage_in;
  std::vector<std::unique_ptr<DevBuf>> stage_out;
  // hfr_model_submit_host / hfr_model_wait_host: kHostSlots batches in flight - the H2D copy of slot i+1 and the D2H copy
  // of slot i-1 run on their own streams while slot i computes
  static constexpr int kHostSlots = 4;
  struct HostSlot {
    DevBuf in;
    std::vector<std::unique_ptr<DevBuf>> out;
    cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr;
    bool busy = false;
  };
  HostSlot slots[kHostSlots];
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  void init_host_pipeline() {
    if (copy_in) return;
    cuda_check(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking), "cudaStreamCreate(copy in)");
    cuda_check(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking), "cudaStreamCreate(copy out)");
    for (HostSlot& h : slots) {
      cuda_check(cudaEventCreateWithFlags(&h.in_done, cudaEventDisableTiming), "cudaEventCreate");
      cuda_check(cudaEventCreateWithFlags(&h.comp_done, cudaEventDisableTiming), "cudaEventCreate");
      cuda_check(cudaEventCreateWithFlags(&h.out_done, cudaEventDisableTiming), "cudaEventCreate");
    }
  }
  void free_host_pipeline() {
    if (!copy_in) return;
    for (HostSlot& h : slots) {
      cudaEventDestroy(h.in_done);
      cudaEventDestroy(h.comp_done);
      cudaEventDestroy(h.out_done);
    }
    cudaStreamDestroy(copy_in);
    cudaStreamDestroy(copy_out);
    copy_in = copy_out = nullptr;
  }

  ~hfr_model() {
    free_host_pipeline();
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    for (auto e : ev) cudaEventDestroy(e);
    for (auto& kv : stem_w2) cudaFree(kv.second.first);
    for (auto& d : dev) {
      if (d.w) cudaFree(d.w);
      if (d.w_win) cudaFree(d.w_win);
      if (d.bias) cudaFree(d.bias);
    }
  }

  size_t value_image_bytes(int v) const {
    const ValueInfo& vi = plan.values[(size_t)v];
    const size_t n = (size_t)vi.H * vi.W * vi.C;
    const size_t b = vi.is_vector ? n * 4 : n * elt_size(precision);
    return (b + 255) / 256 * 256;
  }

  // Strided 1x1 convolutions: the plan states them as gather (L_SUBSAMPLE) + GEMM.  On the tensor-core path the GEMM's A
  // operand can be fetched by an im2col tensor map with a traversal stride instead, straight from the un-gathered tensor:
  // gather_of[li] >= 0 names the L_SUBSAMPLE layer a pointwise layer li bypasses; a gather whose consumers all bypass it
  // is skipped.  Off when every activation has to exist (keep_all) or in fp32 (SIMT GEMM has no im2col operand).
  std::vector<int> gather_of;
  std::vector<char> skip_layer;
  void plan_gather_bypass() {
    const size_t n = plan.layers.size();
    gather_of.assign(n, -1);
    skip_layer.assign(n, 0);
    if (keep_all || precision == HFR_FP32 || getenv("HFR_NO_STRIDED_GEMM")) return;
    const int bk = 128 / (int)elt_size(precision);
    for (size_t si = 0; si < n; ++si) {
      const Layer& S = plan.layers[si];
      if (S.kind != L_SUBSAMPLE || S.cin % bk) continue;
      bool all = true, any = false;
      for (size_t li = 0; li < n; ++li) {
        const Layer& L = plan.layers[li];
        if (L.in2 == S.out) all = false;
        if (L.in != S.out) continue;
        if (L.kind == L_PW && li > si) any = true; else all = false;
      }
      for (int o : plan.outputs) if (o == S.out) all = false;
      if (!all || !any) continue;
      skip_layer[si] = 1;
      for (size_t li = si + 1; li < n; ++li)
        if (plan.layers[li].in == S.out) gather_of[li] = (int)si;
    }
  }

  // First block of a ResNet stage:  y = ReLU(increase(x_mid) + projection(x_in)), stated by the plan as a 1x1 convolution
  // without activation ('increase_bn') whose output is the residual input of the next 1x1 convolution (the projection
  // shortcut).  Both are linear, so they are ONE GEMM over the concatenated reduction dimension,
  // [x_mid | x_in] [W_inc | W_proj]^T + (b_inc + b_proj): the 'increase' tensor is never written or re-read (411 MB each
  // way in stage 2 at batch 256) and a launch disappears; the sum is accumulated in fp32 instead of being rounded to
  // the storage type in between.  kcat_of[j] = i: layer j absorbs layer i.  Off when every activation has to exist.
  std::vector<int> kcat_of;
  std::vector<char> kcat_skip;
  bool kcat_candidate(size_t i) const {   // pattern only (what upload_weights prepares), independent of the run mode
    if (i + 1 >= plan.layers.size() || precision == HFR_FP32) return false;
    const Layer& A = plan.layers[i];
    const Layer& B = plan.layers[i + 1];
    if (A.kind != L_PW || B.kind != L_PW || A.act != A_NONE || A.in2 >= 0 || B.in2 != A.out || B.in == A.out) return false;
    if (A.cout != B.cout || A.Ho != B.Ho || A.Wo != B.Wo || A.bias.empty() != B.bias.empty()) return false;
    const int bk = 128 / (int)elt_size(precision);
    if (A.cin % bk || B.cin % bk) return false;
    for (size_t li = 0; li < plan.layers.size(); ++li) {      // nobody else reads the intermediate
      if (li == i + 1) continue;
      if (plan.layers[li].in == A.out || plan.layers[li].in2 == A.out) return false;
    }
    for (int o : plan.outputs) if (o == A.out) return false;
    return true;
  }
  void plan_kcat() {
    const size_t n = plan.layers.size();
    kcat_of.assign(n, -1);
    kcat_skip.assign(n, 0);
    const char* e = getenv("HFR_KCAT");
    if (keep_all || (e && atoi(e) == 0)) return;
    // (a host-only handle has no device weights; HFR_PLAN_ASSUME_KCAT=1 lets it plan as the device handle would, so that
    // the arena layout of the fused plan can be checked without a GPU: tests/test_arena_cpu.py)
    const bool assume = device < 0 && getenv("HFR_PLAN_ASSUME_KCAT") != nullptr;
    for (size_t i = 0; i + 1 < n; ++i) {
      if (!kcat_candidate(i) || gather_of[i] >= 0) continue;
      if (!assume && (dev.size() != n || dev[i + 1].w_cat == nullptr)) continue;
      kcat_of[i + 1] = (int)i;
      kcat_skip[i] = 1;
    }
  }

  // The dense tail (a hidden Dense layer on the pooled vector followed by the output heads) runs as one launch of
  // dense_heads_kernel when it has that shape: layers [head_first, head_first + head_count) of the plan.
  int head_first = -1, head_count = 0;
  void plan_heads() {
    head_first = -1;
    head_count = 0;
    if (getenv("HFR_NO_FUSED_HEADS")) return;
    const int n = (int)plan.layers.size();
    for (int i = 0; i < n; ++i) {
      const Layer& H = plan.layers[(size_t)i];
      if (H.kind != L_FC || H.in2 >= 0 || (H.act != A_NONE && H.act != A_RELU)) continue;
      int cnt = 1, widths[4] = {0, 0, 0, 0};
      while (i + cnt < n && cnt <= 4) {
        const Layer& L = plan.layers[(size_t)(i + cnt)];
        if (L.kind != L_FC || L.in != H.out || L.in2 >= 0) break;
        widths[cnt - 1] = L.cout;
        ++cnt;
      }
      if (i + cnt != n && plan.layers[(size_t)(i + cnt)].kind == L_FC) continue;   // a longer / different dense chain
      if (cnt < 2 || !dense_heads_supported(H.cin, H.cout, cnt - 1, widths)) continue;
      bool later_use = false;   // nothing after the group may read the hidden vector through another kind of layer
      for (int j = i + cnt; j < n; ++j)
        if (plan.layers[(size_t)j].in == H.out || plan.layers[(size_t)j].in2 == H.out) later_use = true;
      if (later_use) continue;
      head_first = i;
      head_count = cnt;
      return;
    }
  }

  void plan_arena() {
    plan_heads();
    const int nv = (int)plan.values.size();
    val_off.assign((size_t)nv, 0);
    val_bytes.assign((size_t)nv, 0);
    plan_gather_bypass();
    plan_kcat();
    // a bypassing layer reads the gather's INPUT: that value lives until its last bypassing reader
    std::vector<int> last_use((size_t)nv);
    for (int v = 0; v < nv; ++v) last_use[(size_t)v] = plan.values[(size_t)v].last_use;
    for (size_t li = 0; li < plan.layers.size(); ++li) {
      if (gather_of[li] >= 0) {
        const int src = plan.layers[(size_t)gather_of[li]].in;
        if (src > 0 && last_use[(size_t)src] < (int)li) last_use[(size_t)src] = (int)li;
      }
      if (kcat_of[li] >= 0) {   // the absorbing layer reads the absorbed layer's input
        const int src = plan.layers[(size_t)kcat_of[li]].in;
        if (src > 0 && last_use[(size_t)src] < (int)li) last_use[(size_t)src] = (int)li;
      }
    }
    struct Block { size_t off, size; };
    std::vector<Block> free_list;
    size_t top = 0;
    auto alloc = [&](size_t sz) {
      for (size_t i = 0; i < free_list.size(); ++i) {
        if (free_list[i].size >= sz) {
          size_t off = free_list[i].off;
          free_list[i].off += sz;
          free_list[i].size -= sz;
          if (!free_list[i].size) free_list.erase(free_list.begin() + (long)i);
          return off;
        }
      }
      size_t off = top;
      top += sz;
      return off;
    };
    auto release = [&](size_t off, size_t sz) {
      free_list.push_back({off, sz});
      std::sort(free_list.begin(), free_list.end(), [](const Block& a, const Block& b) { return a.off < b.off; });
      for (size_t i = 0; i + 1 < free_list.size();) {
        if (free_list[i].off + free_list[i].size == free_list[i + 1].off) {
          free_list[i].size += free_list[i + 1].size;
          free_list.erase(free_list.begin() + (long)i + 1);
        } else {
          ++i;
        }
      }
    };
    // Two 1x1 convolutions that may run as ONE gemm_pair_kernel launch (decided per call: gemm_pair_eligible): the second
    // layer's output is written while other CTAs still read the first layer's inputs, so those inputs must outlive the
    // second layer's allocation - otherwise first-fit could hand the second output the slot the first layer's residual
    // (or input) has just left.  Every input of layer li stays live through layer li + 1.
    for (size_t li = 0; li + 1 < plan.layers.size(); ++li) {
      const Layer& A = plan.layers[li];
      const Layer& B = plan.layers[li + 1];
      if (precision == HFR_FP32 || A.kind != L_PW || B.kind != L_PW || B.in != A.out || B.in2 >= 0) continue;
      if (gather_of[li] >= 0 || gather_of[li + 1] >= 0 || kcat_skip[li]) continue;
      int ins[3] = {A.in, kcat_of[li] >= 0 ? -1 : A.in2, kcat_of[li] >= 0 ? plan.layers[(size_t)kcat_of[li]].in : -1};
      for (int v : ins)
        if (v > 0 && last_use[(size_t)v] < (int)li + 1) last_use[(size_t)v] = (int)li + 1;
    }
    // The fused dense tail (dense_heads_kernel: hidden layer + every head in one launch) writes the heads' outputs while
    // other CTAs still read the pooled vector: it stays live through the last layer of the group.
    if (head_first >= 0 && head_count > 1) {
      const int v = plan.layers[(size_t)head_first].in;
      const int last = head_first + head_count - 1;
      if (v > 0 && last_use[(size_t)v] < last) last_use[(size_t)v] = last;
    }
    val_release.assign((size_t)nv, -1);
    std::vector<char> live((size_t)nv, 0);
    for (int li = 0; li < (int)plan.layers.size(); ++li) {
      const Layer& L = plan.layers[(size_t)li];
      if (!(skip_layer[(size_t)li] || kcat_skip[(size_t)li])) {   // (those layers' outputs are never materialised)
        val_bytes[(size_t)L.out] = value_image_bytes(L.out);
        val_off[(size_t)L.out] = alloc(val_bytes[(size_t)L.out]);
        live[(size_t)L.out] = 1;
      }
      if (keep_all) continue;
      // every value whose last reader is this layer leaves the arena now (after this layer's output was placed)
      for (int v = 1; v < nv; ++v) {
        if (!live[(size_t)v]) continue;
        if (last_use[(size_t)v] == li || (v == L.out && last_use[(size_t)v] < 0)) {
          release(val_off[(size_t)v], val_bytes[(size_t)v]);
          live[(size_t)v] = 0;
          val_release[(size_t)v] = li;
        }
      }
    }
    per_image_bytes = top;
  }

  void upload_weights() {
    // does any consumer of value v run on the tensor cores?  (tf32 mode: producers round their output to tf32)
    dev.resize(plan.layers.size());
    for (size_t i = 0; i < plan.layers.size(); ++i) {
      const Layer& L = plan.layers[i];
      LayerDev& d = dev[i];
      if (!L.bias.empty()) d.bias = (float*)upload(L.bias.data(), L.bias.size() * 4);
      if (L.w.empty()) continue;
      const bool gemm_operand = (L.kind == L_PW || L.kind == L_CONV);
      if (L.kind == L_CONV && precision == HFR_BF16 && getenv("HFR_NO_WINDOW_CONV") == nullptr && L.stride == 1 && L.dil == 1 &&
          (L.cout == 64 || L.cout == 32) && conv_window_fits(L.cin, L.kh, L.kw)) {
        std::vector<uint16_t> hw = conv_window_layout(L.w.data(), L.cout, L.kh * L.kw, L.cin);
        d.w_win = upload(hw.data(), hw.size() * 2);
      }
      if (gemm_operand && precision == HFR_BF16) {
        std::vector<uint16_t> h(L.w.size());
        for (size_t j = 0; j < h.size(); ++j) h[j] = f32_to_bf16(L.w[j]);
        d.w = upload(h.data(), h.size() * 2);
      } else if (gemm_operand && precision == HFR_TF32) {
        std::vector<float> h(L.w.size());
        for (size_t j = 0; j < h.size(); ++j) h[j] = f32_to_tf32(L.w[j]);
        d.w = upload(h.data(), h.size() * 4);
      } else {
        d.w = upload(L.w.data(), L.w.size() * 4);
      }
    }
    for (size_t i = 0; i + 1 < plan.layers.size(); ++i) {
      if (!kcat_candidate(i)) continue;
      const Layer& A = plan.layers[i];
      const Layer& B = plan.layers[i + 1];
      const size_t N = (size_t)A.cout, Ka = (size_t)A.cin, Kb = (size_t)B.cin;
      if (A.w.size() != N * Ka || B.w.size() != N * Kb) continue;
      std::vector<float> cat(N * (Ka + Kb));
      for (size_t n = 0; n < N; ++n) {
        std::copy(A.w.begin() + (long)(n * Ka), A.w.begin() + (long)((n + 1) * Ka), cat.begin() + (long)(n * (Ka + Kb)));
        std::copy(B.w.begin() + (long)(n * Kb), B.w.begin() + (long)((n + 1) * Kb), cat.begin() + (long)(n * (Ka + Kb) + Ka));
      }
      LayerDev& d = dev[i + 1];
      if (precision == HFR_BF16) {
        std::vector<uint16_t> h(cat.size());
        for (size_t j = 0; j < h.size(); ++j) h[j] = f32_to_bf16(cat[j]);
        d.w_cat = upload(h.data(), h.size() * 2);
      } else {
        for (float& v : cat) v = f32_to_tf32(v);
        d.w_cat = upload(cat.data(), cat.size() * 4);
      }
      if (!A.bias.empty()) {
        std::vector<float> bsum(N);
        for (size_t n = 0; n < N; ++n) bsum[n] = A.bias[n] + B.bias[n];
        d.bias_cat = (float*)upload(bsum.data(), N * 4);
      }
    }
    plan_arena();   // the K-concatenation plan depends on the uploaded weights
  }

  bool feeds_tensor_core(int v) const {
    for (const Layer& L : plan.layers)
      if ((L.in == v) && (L.kind == L_PW || L.kind == L_CONV || L.kind == L_SUBSAMPLE)) return true;
    return false;
  }

  void* val_ptr(int v, int batch) const { return (char*)arena.p + val_off[(size_t)v] * (size_t)batch; }

  // Every cached graph bakes device pointers (arena, stem staging buffer) into its kernel arguments: whenever one of
  // those buffers is reallocated the cache goes, after everything that may still be replaying it has drained.
  void drop_graphs() {
    if (graphs.empty()) return;
    cuda_check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    graphs.clear();
  }

  void ensure_arena(int batch) {
    const size_t need = per_image_bytes * (size_t)batch;
    if (need > arena.bytes) {
      drop_graphs();
      arena.ensure(need);
    }
  }

  // exact (unpadded) bytes of one image of value v
  size_t value_exact_bytes(int v) const {
    const ValueInfo& vi = plan.values[(size_t)v];
    return (size_t)vi.H * vi.W * vi.C * (vi.is_vector ? 4 : elt_size(precision));
  }

  void run_layers(const void* x, int in_dtype, int batch, int flags, void* const* outs, cudaStream_t s) {
    // L2-sized sub-batching: every tensor is batch-major, so a slice of the batch is a pointer offset.  Running the
    // whole layer list on one slice at a time keeps producer->consumer tensors L2-resident.
    int sub = sub_batch > 0 && !timing ? sub_batch : batch;
    if (sub > batch) sub = batch;
    for (int b0 = 0; b0 < batch; b0 += sub) run_slice(x, in_dtype, batch, b0, std::min(sub, batch - b0), flags, s);
    finish_outputs(batch, flags, outs, s);
  }

  void run_slice(const void* x_all, int in_dtype, int total, int img0, int batch, int flags, cudaStream_t s) {
    const int prec = precision;
    const int rt = (prec == HFR_TF32);
    const size_t in_img_bytes = (size_t)plan.in_h * plan.in_w * plan.in_c * (in_dtype == HFR_IN_U8 ? 1 : 4);
    const void* x = (const char*)x_all + (size_t)img0 * in_img_bytes;
    std::vector<char> launched(plan.layers.size(), 1);
    auto vptr = [&](int v) -> void* { return (char*)val_ptr(v, total) + (size_t)img0 * value_exact_bytes(v); };
    for (size_t i = 0; i < plan.layers.size(); ++i) {
      const Layer& L = plan.layers[i];
      const LayerDev& d = dev[i];
      const void* in = L.in == 0 ? x : vptr(L.in);
      void* out = vptr(L.out);
      const int act = L.act;  // A_NONE/A_RELU/A_RELU6 share values with the kernels' ACT_* codes
      const int round_out = rt && feeds_tensor_core(L.out);
      if (timing) cuda_check(cudaEventRecord(ev[2 * i], s), "cudaEventRecord");
      launched[i] = 1;
      if ((int)i == head_first) {   // hidden Dense + its heads: one launch, booked on the hidden layer
        HeadsArgs h;
        memset(&h, 0, sizeof(h));
        h.x = (const float*)in; h.w1 = (const float*)d.w; h.b1 = d.bias; h.hidden = (float*)out;
        h.B = batch; h.K = L.cin; h.n1 = L.cout; h.act1 = act; h.n_heads = head_count - 1;
        for (int j = 1; j < head_count; ++j) {
          const Layer& P = plan.layers[i + (size_t)j];
          h.w[j - 1] = (const float*)dev[i + (size_t)j].w; h.b[j - 1] = dev[i + (size_t)j].bias;
          h.y[j - 1] = (float*)vptr(P.out); h.n[j - 1] = P.cout; h.act[j - 1] = P.act;
        }
        launch_dense_heads(h, s);
        if (timing) cuda_check(cudaEventRecord(ev[2 * i + 1], s), "cudaEventRecord");
        for (int j = 1; j < head_count; ++j) {
          launched[i + (size_t)j] = 0;
          if (timing) {
            cuda_check(cudaEventRecord(ev[2 * (i + (size_t)j)], s), "cudaEventRecord");
            cuda_check(cudaEventRecord(ev[2 * (i + (size_t)j) + 1], s), "cudaEventRecord");
          }
        }
        i += (size_t)head_count - 1;
        continue;
      }
      switch (L.kind) {
        case L_STEM: {
          StemArgs a;
          memset(&a, 0, sizeof(a));
          a.x = in; a.in_u8 = (in_dtype == HFR_IN_U8); a.w = (const float*)d.w; a.bias = d.bias; a.y = out;
          a.B = batch; a.H = L.H; a.W = L.W; a.Ho = L.Ho; a.Wo = L.Wo; a.kh = L.kh; a.kw = L.kw; a.stride = L.stride;
          a.pad_t = L.pad_t; a.pad_l = L.pad_l; a.cout = L.cout;
          a.scale = 1.f;
          if (a.in_u8) {
            a.flip = (flags & HFR_FLAG_BGR) ? 1 : 0;
            if (flags & HFR_FLAG_MEAN_IMAGENET) { a.mean[0] = 103.939f; a.mean[1] = 116.779f; a.mean[2] = 123.68f; }
            if (flags & HFR_FLAG_MEAN_VGGFACE2) { a.mean[0] = 91.4953f; a.mean[1] = 103.8827f; a.mean[2] = 131.0912f; }
            if (flags & HFR_FLAG_SCALE_PM1) { a.scale = 1.f / 127.5f; a.mean[0] = a.mean[1] = a.mean[2] = 1.f; }
          }
          a.act = act; a.round_tf32 = round_out;
          // the tf32 mode's stem runs on the tensor cores as well: fp32 output, weights as hi + lo bf16 sweeps into one
          // accumulator (16 mantissa bits; the uint8 image is exact in bf16), window kernel only
          const bool tc32 = prec == HFR_TF32 && stem_window && tf32_tc_stem;
          const bool tc = (prec == HFR_BF16 || tc32) && a.in_u8 && L.stride == 2 && L.dil == 1 && (L.H % 2 == 0) &&
                          (L.W % 2 == 0) && (L.cout == 32 || L.cout == 64) && !stem_force_direct;
          if (!tc) {
            launch_stem(a, prec, s);
            break;
          }
          const int key = flags & (HFR_FLAG_BGR | HFR_FLAG_MEAN_IMAGENET | HFR_FLAG_MEAN_VGGFACE2 | HFR_FLAG_SCALE_PM1);
          auto it = stem_w2.find(key);
          if (it == stem_w2.end()) {
            StemTcGeom g;
            std::vector<uint16_t> h =
                stem_tc_weights(L.w, L.kh, L.kw, L.cout, L.pad_t, L.pad_l, a.flip, a.scale, a.mean, &g, tc32);
            if (stem_window) h = stem_window_layout(h, g.ka * g.kb * (tc32 ? 2 : 1), L.cout);
            it = stem_w2.emplace(key, std::make_pair(upload(h.data(), h.size() * 2), g)).first;
          }
          const StemTcGeom& g = it->second.second;
          StemTcArgs t;
          t.x = (const uint8_t*)in; t.w2 = it->second.first; t.bias = d.bias; t.y = out;
          t.B = batch; t.H = L.H; t.W = L.W; t.Ho = L.Ho; t.Wo = L.Wo; t.cout = L.cout;
          t.ka = g.ka; t.kb = g.kb; t.pt2 = g.pt2; t.pl2 = g.pl2; t.act = act;
          t.use_window = stem_window;
          if (tc32) {
            t.out_f32 = 1;
            t.passes = 2;
            t.round_tf32 = round_out;
          }
          const size_t need = (size_t)batch * (L.Ho + g.ka - 1) * (L.Wo + g.kb - 1) * 32;
          if (need > stem_scratch.bytes) {
            // never reached under stream capture: forward() runs every new (batch, flags) eagerly once before capturing
            cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");  // previous users of the old buffer
            drop_graphs();  // graphs captured at smaller batches hold the old pointer
            stem_scratch.ensure(need);
          }
          t.scratch = stem_scratch.p;
          launch_stem_tc(t, device, s);
          break;
        }
        case L_DW: {
          DwArgs a;
          a.x = in; a.w = (const float*)d.w; a.bias = d.bias; a.y = out;
          a.B = batch; a.H = L.H; a.W = L.W; a.C = L.cin; a.Ho = L.Ho; a.Wo = L.Wo; a.stride = L.stride;
          a.pad_t = L.pad_t; a.pad_l = L.pad_l; a.act = act; a.round_tf32 = round_out;
          launch_dw(a, prec, s);
          break;
        }
        case L_PW: {
          if (kcat_skip[i]) { launched[i] = 0; break; }   // absorbed by the next layer's GEMM (plan_kcat)
          // K-concatenation: this GEMM also reduces over the absorbed layer's input with its weights, no residual
          const bool kcat = kcat_of[i] >= 0;
          const Layer* KA = kcat ? &plan.layers[(size_t)kcat_of[i]] : nullptr;
          const void* w_use = kcat ? d.w_cat : d.w;
          const float* bias_use = kcat ? d.bias_cat : d.bias;
          const void* res_use = kcat ? nullptr : (L.in2 >= 0 ? vptr(L.in2) : nullptr);
          const void* a0_use = kcat ? (KA->in == 0 ? x : vptr(KA->in)) : nullptr;
          const int k0_use = kcat ? KA->cin : 0;
          if (gather_of[i] >= 0) {  // strided 1x1: im2col-gathered A operand, no materialised subsample
            const Layer& S = plan.layers[(size_t)gather_of[i]];
            ConvArgs a;
            a.x = S.in == 0 ? x : vptr(S.in); a.w = w_use; a.bias = bias_use;
            a.residual = res_use; a.a0 = a0_use; a.K0 = k0_use;
            a.y = out; a.B = batch; a.H = S.H; a.W = S.W; a.cin = L.cin; a.Ho = L.Ho; a.Wo = L.Wo; a.cout = L.cout;
            a.kh = 1; a.kw = 1; a.stride = S.stride; a.pad_t = 0; a.pad_l = 0; a.dil = 1;
            a.act = act; a.round_tf32 = round_out;
            launch_conv(a, prec, device, s);
            break;
          }
          GemmArgs a;
          a.a = in; a.b = w_use; a.bias = bias_use;
          a.residual = res_use; a.a0 = a0_use; a.K0 = k0_use;
          a.y = out; a.M = (int64_t)batch * L.Ho * L.Wo; a.N = L.cout; a.K = L.cin;
          a.act = act; a.round_tf32 = round_out;
          // the next layer is a plain 1x1 convolution over this layer's output (the seam between two bottleneck blocks):
          // both GEMMs in one launch, the second fed from the first's staged output chunks (gemm_pair.cuh)
          if (i + 1 < plan.layers.size()) {
            const Layer& N2 = plan.layers[i + 1];
            if (N2.kind == L_PW && N2.in == L.out && N2.in2 < 0 && gather_of[i + 1] < 0) {
              GemmArgs b;
              b.a = out; b.b = dev[i + 1].w; b.bias = dev[i + 1].bias; b.residual = nullptr;
              b.y = vptr(N2.out); b.M = a.M; b.N = N2.cout; b.K = N2.cin;
              b.act = N2.act; b.round_tf32 = rt && feeds_tensor_core(N2.out);
              if (gemm_pair_eligible(a, b, prec, device)) {
                launch_gemm_pair(a, b, prec, device, s);
                if (timing) {
                  cuda_check(cudaEventRecord(ev[2 * i + 1], s), "cudaEventRecord");
                  cuda_check(cudaEventRecord(ev[2 * i + 2], s), "cudaEventRecord");
                  cuda_check(cudaEventRecord(ev[2 * i + 3], s), "cudaEventRecord");
                }
                launched[i + 1] = 0;
                ++i;
                continue;
              }
            }
          }
          launch_gemm(a, prec, device, s);
          break;
        }
        case L_CONV: {
          if (d.w_win != nullptr && L.in2 < 0) {
            WinArgs w;
            w.x = in; w.w = d.w_win; w.bias = d.bias; w.y = out; w.B = batch; w.H = L.H; w.W = L.W; w.cin = L.cin;
            w.Ho = L.Ho; w.Wo = L.Wo; w.cout = L.cout; w.kh = L.kh; w.kw = L.kw; w.pad_t = L.pad_t; w.pad_l = L.pad_l;
            w.act = act; w.plane_major = 0;
            launch_conv_window(w, device, s);
            break;
          }
          ConvArgs a;
          a.x = in; a.w = d.w; a.bias = d.bias;
          a.residual = L.in2 >= 0 ? vptr(L.in2) : nullptr;
          a.y = out; a.B = batch; a.H = L.H; a.W = L.W; a.cin = L.cin; a.Ho = L.Ho; a.Wo = L.Wo; a.cout = L.cout;
          a.kh = L.kh; a.kw = L.kw; a.stride = L.stride; a.pad_t = L.pad_t; a.pad_l = L.pad_l; a.dil = L.dil;
          a.act = act; a.round_tf32 = round_out;
          launch_conv(a, prec, device, s);
          break;
        }
        case L_MAXPOOL: {
          PoolArgs a;
          a.x = in; a.y = out; a.B = batch; a.H = L.H; a.W = L.W; a.C = L.cin; a.Ho = L.Ho; a.Wo = L.Wo; a.k = L.kh;
          a.stride = L.stride; a.pad_t = L.pad_t; a.pad_l = L.pad_l; a.explicit_zero = L.explicit_zero_pad;
          launch_maxpool(a, prec, s);
          break;
        }
        case L_SUBSAMPLE:
          if (skip_layer[i]) { launched[i] = 0; break; }
          launch_subsample(in, out, batch, L.H, L.W, L.cin, L.Ho, L.Wo, L.stride, prec, s);
          break;
        case L_GAP:
          launch_gap(in, (float*)out, batch, L.H * L.W, L.cin, prec, s);
          break;
        case L_FC:
          launch_fc((const float*)in, (const float*)d.w, d.bias, (float*)out, batch, L.cin, L.cout, act, s);
          break;
        default:
          throw Error(HFR_ERR_UNSUPPORTED, "internal: unknown layer kind");
      }
      if (timing) cuda_check(cudaEventRecord(ev[2 * i + 1], s), "cudaEventRecord");
    }
    if (timing) {
      // read back the previous step's events lazily would need double buffering; steps under timing are few, so sync
      cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize(timing)");
      for (size_t i = 0; i < plan.layers.size(); ++i) {
        float ms = 0.f;
        cuda_check(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]), "cudaEventElapsedTime");
        if (launched[i]) layer_ms[i] += ms;
        else layer_ms[i] = -1.0;   // nothing was launched for this layer (bypassed gather / fused into a neighbour)
      }
      ++timed_steps;
    }
  }

  void finish_outputs(int batch, int flags, void* const* outs, cudaStream_t s) {
    const int prec = precision;
    for (size_t o = 0; o < plan.outputs.size(); ++o) {
      const int v = plan.outputs[o];
      const ValueInfo& vi = plan.values[(size_t)v];
      const int64_t dim = (int64_t)vi.H * vi.W * vi.C;
      if (!vi.is_vector) {
        launch_cast_to_f32(val_ptr(v, batch), (float*)outs[o], dim * batch, prec, s);
        if (o == 0 && (flags & HFR_FLAG_L2NORM)) launch_l2norm((float*)outs[o], (float*)outs[o], batch, (int)dim, s);
      } else if (o == 0 && (flags & HFR_FLAG_L2NORM)) {
        launch_l2norm((const float*)val_ptr(v, batch), (float*)outs[o], batch, (int)dim, s);
      } else {
        cuda_check(cudaMemcpyAsync(outs[o], val_ptr(v, batch), (size_t)dim * batch * 4, cudaMemcpyDeviceToDevice, s),
                   "cudaMemcpyAsync(output)");
      }
    }
  }

  void forward(const void* x, int in_dtype, int batch, int flags, void* const* outs, cudaStream_t s) {
    if (device < 0) throw Error(HFR_ERR_STATE, "model was loaded host-only (device < 0); forward needs a GPU handle");
    if (batch <= 0) throw Error(HFR_ERR_INVALID, "batch must be positive");
    if (in_dtype != HFR_IN_F32 && in_dtype != HFR_IN_U8) throw Error(HFR_ERR_INVALID, "bad input dtype");
    use_device(device);
    ensure_arena(batch);
    last_batch = batch;
    if (!(flags & HFR_FLAG_CUDA_GRAPH) || s == nullptr || timing) {
      run_layers(x, in_dtype, batch, flags, outs, s);
      return;
    }
    GraphKey key{x, in_dtype, batch, flags, std::vector<void*>(outs, outs + plan.outputs.size())};
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      if (graphs.size() >= 64) drop_graphs();  // (batches submitted through hfr_model_submit_host may still replay them)
      run_layers(x, in_dtype, batch, flags, outs, s);  // eager first pass: configures kernels, validates arguments
      cudaGraph_t g = nullptr;
      cuda_check(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
      try {
        run_layers(x, in_dtype, batch, flags, outs, s);
      } catch (...) {
        cudaStreamEndCapture(s, &g);
        if (g) cudaGraphDestroy(g);
        throw;
      }
      cuda_check(cudaStreamEndCapture(s, &g), "cudaStreamEndCapture");
      cudaGraphExec_t ge = nullptr;
      cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
      cudaGraphDestroy(g);
      cuda_check(e, "cudaGraphInstantiate");
      graphs[key] = ge;
      return;  // the eager pass already produced this step's result
    }
    cuda_check(cudaGraphLaunch(it->second, s), "cudaGraphLaunch");
  }
};

static std::vector<uint8_t> read_file(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(HFR_ERR_IO, std::string("cannot open '") + path + "'");
  f.seekg(0, std::ios::end);
  std::streamoff n = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> data((size_t)n);
  if (n && !f.read((char*)data.data(), n)) throw Error(HFR_ERR_IO, std::string("cannot read '") + path + "'");
  return data;
}

extern "C" {

const char* hfr_last_error(void) { return g_err.c_str(); }
int hfr_version(void) { return 100; }
int64_t hfr_launch_count(void) { return launch_count(); }

int hfr_model_load(const char* path, const char* input_name, const char* output_names_csv, const char* phase_name,
                   float phase_value, int input_hw, int device, int precision, hfr_model** out) {
  return guarded([&] {
    if (!path || !out) throw Error(HFR_ERR_INVALID, "null argument");
    if (precision < HFR_FP32 || precision > HFR_BF16) throw Error(HFR_ERR_INVALID, "bad precision");
    std::vector<uint8_t> data = read_file(path);
    std::unique_ptr<hfr_model> m(new hfr_model());
    m->precision = precision;
    m->device = device;
    static const uint8_t kHdf5Magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    CompileOptions opt;
    if (output_names_csv) {
      std::stringstream ss(output_names_csv);
      std::string tok;
      while (std::getline(ss, tok, ',')) {
        while (!tok.empty() && tok.front() == ' ') tok.erase(tok.begin());
        while (!tok.empty() && tok.back() == ' ') tok.pop_back();
        if (!tok.empty()) opt.output_names.push_back(tok);
      }
    }
    if (phase_name && *phase_name) opt.phase_name = phase_name;
    opt.phase_value = phase_value;
    opt.override_hw = input_hw;
    if (data.size() >= 8 && memcmp(data.data(), kHdf5Magic, 8) == 0) {
      opt.phase_name.clear();  // a Keras weight file has no learning-phase conditionals
      m->plan = compile_keras_mobilenet_h5(data.data(), data.size(), input_hw, opt);
    } else {
      if (!input_name || !*input_name || opt.output_names.empty())
        throw Error(HFR_ERR_INVALID, "input/output tensor names are required");
      Graph g;
      parse_graphdef(data.data(), data.size(), &g);
      opt.input_name = input_name;
      m->plan = compile_graph(g, opt);
    }
    m->plan_arena();
    if (device >= 0) {
      use_device(device);
      m->upload_weights();
    }
    *out = m.release();
  });
}

int hfr_model_info(const hfr_model* m, int* in_h, int* in_w, int* in_c, int* n_outputs, int* out_dims) {
  return guarded([&] {
    if (!m) throw Error(HFR_ERR_INVALID, "null model");
    if (in_h) *in_h = m->plan.in_h;
    if (in_w) *in_w = m->plan.in_w;
    if (in_c) *in_c = m->plan.in_c;
    if (n_outputs) *n_outputs = (int)m->plan.outputs.size();
    if (out_dims)
      for (size_t i = 0; i < m->plan.outputs.size(); ++i) {
        const ValueInfo& v = m->plan.values[(size_t)m->plan.outputs[i]];
        out_dims[i] = v.H * v.W * v.C;
      }
  });
}

int64_t hfr_model_plan_json(const hfr_model* m, char* buf, int64_t buf_len) {
  if (!m) return HFR_ERR_INVALID;
  std::string js = m->plan.to_json();
  // arena layout per value: [value id, producer layer, byte offset per image, bytes per image, layer after which it is free]
  std::string arena = ",\"arena\":[";
  bool first = true;
  for (size_t v = 1; v < m->plan.values.size() && v < m->val_bytes.size(); ++v) {
    if (m->val_bytes[v] == 0) continue;
    arena += std::string(first ? "" : ",") + "[" + std::to_string(v) + "," + std::to_string(m->plan.values[v].producer) + "," +
             std::to_string(m->val_off[v]) + "," + std::to_string(m->val_bytes[v]) + "," +
             std::to_string(v < m->val_release.size() ? m->val_release[v] : -1) + "]";
    first = false;
  }
  arena += "]";
  js.insert(js.size() - 1, ",\"arena_bytes_per_image\":" + std::to_string(m->per_image_bytes) +
                               ",\"precision\":" + std::to_string(m->precision) + arena);
  if (buf && buf_len > (int64_t)js.size()) memcpy(buf, js.c_str(), js.size() + 1);
  return (int64_t)js.size() + 1;
}

int64_t hfr_model_layer_weights(const hfr_model* m, int layer, float* w, int64_t w_cap, float* bias, int64_t b_cap) {
  if (!m || layer < 0 || layer >= (int)m->plan.layers.size()) return HFR_ERR_INVALID;
  const Layer& L = m->plan.layers[(size_t)layer];
  if (w && w_cap >= (int64_t)L.w.size()) memcpy(w, L.w.data(), L.w.size() * 4);
  if (bias && b_cap >= (int64_t)L.bias.size()) memcpy(bias, L.bias.data(), L.bias.size() * 4);
  return (int64_t)L.w.size();
}

int hfr_model_forward(hfr_model* m, const void* x, int in_dtype, int batch, int flags, void* const* outs, void* stream) {
  return guarded([&] {
    if (!m || !x || !outs) throw Error(HFR_ERR_INVALID, "null argument");
    m->forward(x, in_dtype, batch, flags, outs, (cudaStream_t)stream);
  });
}

int hfr_model_forward_host(hfr_model* m, const void* x_host, int in_dtype, int batch, int flags, void* const* outs_host,
                           void* stream) {
  return guarded([&] {
    if (!m || !x_host || !outs_host) throw Error(HFR_ERR_INVALID, "null argument");
    if (m->device < 0) throw Error(HFR_ERR_STATE, "model was loaded host-only");
    if (batch <= 0) throw Error(HFR_ERR_INVALID, "batch must be positive");
    use_device(m->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t in_bytes =
        (size_t)batch * m->plan.in_h * m->plan.in_w * m->plan.in_c * (in_dtype == HFR_IN_U8 ? 1 : 4);
    m->stage_in.ensure(in_bytes);
    const size_t no = m->plan.outputs.size();
    if (m->stage_out.size() < no)
      for (size_t i = m->stage_out.size(); i < no; ++i) m->stage_out.emplace_back(new DevBuf());
    std::vector<void*> douts(no);
    std::vector<size_t> obytes(no);
    for (size_t i = 0; i < no; ++i) {
      const ValueInfo& v = m->plan.values[(size_t)m->plan.outputs[i]];
      obytes[i] = (size_t)batch * v.H * v.W * v.C * 4;
      m->stage_out[i]->ensure(obytes[i]);
      douts[i] = m->stage_out[i]->p;
    }
    cuda_check(cudaMemcpyAsync(m->stage_in.p, x_host, in_bytes, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync(H2D)");
    m->forward(m->stage_in.p, in_dtype, batch, flags, douts.data(), s);
    for (size_t i = 0; i < no; ++i)
      cuda_check(cudaMemcpyAsync(outs_host[i], douts[i], obytes[i], cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync(D2H)");
    cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  });
}

int hfr_model_submit_host(hfr_model* m, int slot, const void* x_host, int in_dtype, int batch, int flags,
                          void* const* outs_host, void* stream) {
  return guarded([&] {
    if (!m || !x_host || !outs_host) throw Error(HFR_ERR_INVALID, "null argument");
    if (m->device < 0) throw Error(HFR_ERR_STATE, "model was loaded host-only");
    if (batch <= 0) throw Error(HFR_ERR_INVALID, "batch must be positive");
    if (slot < 0 || slot >= hfr_model::kHostSlots) throw Error(HFR_ERR_INVALID, "slot out of range");
    use_device(m->device);
    m->init_host_pipeline();
    hfr_model::HostSlot& h = m->slots[slot];
    if (h.busy) throw Error(HFR_ERR_STATE, "slot still in flight: call hfr_model_wait_host(m, slot) first");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t in_bytes =
        (size_t)batch * m->plan.in_h * m->plan.in_w * m->plan.in_c * (in_dtype == HFR_IN_U8 ? 1 : 4);
    h.in.ensure(in_bytes);
    const size_t no = m->plan.outputs.size();
    for (size_t i = h.out.size(); i < no; ++i) h.out.emplace_back(new DevBuf());
    std::vector<void*> douts(no);
    std::vector<size_t> obytes(no);
    for (size_t i = 0; i < no; ++i) {
      const ValueInfo& v = m->plan.values[(size_t)m->plan.outputs[i]];
      obytes[i] = (size_t)batch * v.H * v.W * v.C * 4;
      h.out[i]->ensure(obytes[i]);
      douts[i] = h.out[i]->p;
    }
    // the slot's previous use has been waited for (busy == false), so its buffers are free on every stream
    cuda_check(cudaMemcpyAsync(h.in.p, x_host, in_bytes, cudaMemcpyHostToDevice, m->copy_in), "cudaMemcpyAsync(H2D)");
    cuda_check(cudaEventRecord(h.in_done, m->copy_in), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(s, h.in_done, 0), "cudaStreamWaitEvent");
    m->forward(h.in.p, in_dtype, batch, flags, douts.data(), s);
    cuda_check(cudaEventRecord(h.comp_done, s), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(m->copy_out, h.comp_done, 0), "cudaStreamWaitEvent");
    for (size_t i = 0; i < no; ++i)
      cuda_check(cudaMemcpyAsync(outs_host[i], douts[i], obytes[i], cudaMemcpyDeviceToHost, m->copy_out),
                 "cudaMemcpyAsync(D2H)");
    cuda_check(cudaEventRecord(h.out_done, m->copy_out), "cudaEventRecord");
    h.busy = true;
  });
}

int hfr_model_wait_host(hfr_model* m, int slot) {
  return guarded([&] {
    if (!m) throw Error(HFR_ERR_INVALID, "null model");
    if (slot < 0 || slot >= hfr_model::kHostSlots) throw Error(HFR_ERR_INVALID, "slot out of range");
    hfr_model::HostSlot& h = m->slots[slot];
    if (!h.busy) return;
    cuda_check(cudaEventSynchronize(h.out_done), "cudaEventSynchronize");
    h.busy = false;
  });
}

int hfr_model_set_keep_activations(hfr_model* m, int keep) {
  return guarded([&] {
    if (!m) throw Error(HFR_ERR_INVALID, "null model");
    m->keep_all = keep != 0;
    m->drop_graphs();
    m->plan_arena();
  });
}

int hfr_model_set_layer_timing(hfr_model* m, int enable) {
  return guarded([&] {
    if (!m || m->device < 0) throw Error(HFR_ERR_INVALID, "needs a GPU model handle");
    use_device(m->device);
    if (enable && m->ev.empty()) {
      m->ev.resize(2 * m->plan.layers.size());
      for (auto& e : m->ev) cuda_check(cudaEventCreate(&e), "cudaEventCreate");
    }
    m->timing = enable != 0;
    m->layer_ms.assign(m->plan.layers.size(), 0.0);
    m->timed_steps = 0;
  });
}

int hfr_model_get_layer_times(const hfr_model* m, double* ms_per_layer, int* steps) {
  return guarded([&] {
    if (!m || !ms_per_layer) throw Error(HFR_ERR_INVALID, "null argument");
    for (size_t i = 0; i < m->layer_ms.size(); ++i) ms_per_layer[i] = m->layer_ms[i];
    if (steps) *steps = m->timed_steps;
  });
}

int64_t hfr_model_debug_layer(hfr_model* m, int layer_index, int batch, float* dst, void* stream) {
  int64_t n = 0;
  int rc = guarded([&] {
    if (!m || !dst) throw Error(HFR_ERR_INVALID, "null argument");
    if (!m->keep_all) throw Error(HFR_ERR_STATE, "call hfr_model_set_keep_activations(m, 1) before forward");
    if (layer_index < 0 || layer_index >= (int)m->plan.layers.size()) throw Error(HFR_ERR_INVALID, "bad layer index");
    if (batch != m->last_batch) throw Error(HFR_ERR_STATE, "batch does not match the last forward");
    const int v = m->plan.layers[(size_t)layer_index].out;
    const ValueInfo& vi = m->plan.values[(size_t)v];
    n = (int64_t)vi.H * vi.W * vi.C;
    use_device(m->device);
    if (vi.is_vector)
      cuda_check(cudaMemcpyAsync(dst, m->val_ptr(v, batch), (size_t)n * batch * 4, cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream), "cudaMemcpyAsync");
    else
      launch_cast_to_f32(m->val_ptr(v, batch), dst, n * batch, m->precision, (cudaStream_t)stream);
  });
  return rc < 0 ? rc : n;
}

void hfr_model_free(hfr_model* m) { delete m; }

int hfr_age_gender_post(const float* age_probs, int batch, int n, float* age_out, int device, void* stream) {
  return guarded([&] {
    if (!age_probs || !age_out || batch <= 0 || n < 2) throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(device);
    launch_age_post(age_probs, age_out, batch, n, (cudaStream_t)stream);
  });
}

int hfr_crop_resize_u8(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* boxes, int n,
                       uint8_t* out, int out_h, int out_w, int device, void* stream) {
  return guarded([&] {
    if (!frames || !boxes || !out || n < 0 || n_frames <= 0 || frame_h <= 0 || frame_w <= 0 || out_h <= 0 || out_w <= 0)
      throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(device);
    launch_crop_resize(frames, frame_h, frame_w, (const int*)boxes, n, out, out_h, out_w, (cudaStream_t)stream);
  });
}

int hfr_resize_pil_u8(const uint8_t* images, const int64_t* desc_host, int n, uint8_t* out, int out_h, int out_w,
                      int device, void* stream) {
  return guarded([&] {
    if (!images || !desc_host || !out || n < 0 || out_h <= 0 || out_w <= 0) throw Error(HFR_ERR_INVALID, "bad argument");
    if (n == 0) return;
    use_device(device);
    cudaStream_t s = (cudaStream_t)stream;
    int taps_v = 1, taps_h = 1;
    for (int i = 0; i < n; ++i) {
      const int64_t h = desc_host[4 * i + 1], w = desc_host[4 * i + 2], pitch = desc_host[4 * i + 3];
      if (desc_host[4 * i] < 0 || h <= 0 || w <= 0 || pitch < 3 * w) throw Error(HFR_ERR_INVALID, "bad image descriptor");
      for (int axis = 0; axis < 2; ++axis) {
        const int64_t in_size = axis ? w : h;
        const int out_size = axis ? out_w : out_h;
        if (in_size == out_size) continue;  // pass skipped: one identity tap
        const double scale = (double)in_size / out_size;
        const int taps = (int)std::ceil(scale < 1.0 ? 1.0 : scale) * 2 + 1;
        if (taps > 64) throw Error(HFR_ERR_UNSUPPORTED, "PIL resize: more than 31x reduction is not supported");
        int& t = axis ? taps_h : taps_v;
        if (taps > t) t = taps;
      }
    }
    long long* d_desc = nullptr;
    int* d_tab = nullptr;
    cuda_check(cudaMallocAsync((void**)&d_desc, (size_t)n * 4 * sizeof(long long), s), "cudaMallocAsync(desc)");
    cuda_check(cudaMallocAsync((void**)&d_tab, resize_pil_table_ints(n, out_h, out_w, taps_h, taps_v) * sizeof(int), s),
               "cudaMallocAsync(coefficient tables)");
    static_assert(sizeof(long long) == sizeof(int64_t), "descriptor width");
    // desc_host may be pageable: the copy is staged by the runtime before this call returns
    cuda_check(cudaMemcpyAsync(d_desc, desc_host, (size_t)n * 4 * sizeof(long long), cudaMemcpyHostToDevice, s),
               "cudaMemcpyAsync(desc)");
    launch_resize_pil(images, d_desc, d_tab, n, out, out_h, out_w, taps_h, taps_v, s);
    cuda_check(cudaFreeAsync(d_tab, s), "cudaFreeAsync(tables)");
    cuda_check(cudaFreeAsync(d_desc, s), "cudaFreeAsync(desc)");
  });
}

int hfr_pairwise_dist(const float* x, int64_t n, const float* y, int64_t m, int dim, const float* year_x,
                      const float* born_x, const float* year_y, const float* born_y, float age_weight, float* out,
                      int device, void* stream) {
  return guarded([&] {
    if (!x || !out || n < 0 || m < 0 || dim <= 0) throw Error(HFR_ERR_INVALID, "bad argument");
    if (!y) {
      y = x;
      if (m != n) throw Error(HFR_ERR_INVALID, "y == NULL means y = x: m must equal n");
      year_y = year_x;
      born_y = born_x;
    }
    const int given = (year_x != nullptr) + (born_x != nullptr) + (year_y != nullptr) + (born_y != nullptr);
    if (given != 0 && given != 4) throw Error(HFR_ERR_INVALID, "the age penalty needs all of year/born for both sides");
    use_device(device);
    launch_pairwise_dist(x, y, n, m, dim, year_x, born_x, year_y, born_y, age_weight, out, device, (cudaStream_t)stream);
  });
}

int hfr_debug_gemm_tile_choice(int64_t m, int n, int k, int conv_taps, int sms, int* ctas, int* block_n) {
  return guarded([&] {
    if (!ctas || !block_n || m <= 0 || n <= 0 || k <= 0 || conv_taps < 0 || sms <= 0) throw Error(HFR_ERR_INVALID, "bad argument");
    gemm_tile_choice(m, n, k, conv_taps, sms, ctas, block_n);
  });
}

int hfr_debug_gemm_pair_config(int64_t m, int k0, int k1, int n1, int n2, int has_residual, int precision, int* eligible,
                               int* nbuf, int* pf, int* na, int* stages, int* smem_bytes) {
  return guarded([&] {
    if (!eligible || !nbuf || !pf || !na || !stages || !smem_bytes || m <= 0 || k1 <= 0 || n1 <= 0 || n2 <= 0 || k0 < 0)
      throw Error(HFR_ERR_INVALID, "bad argument");
    // the launcher's own predicate on stand-in operands (only their identity matters: second.a == first.y)
    static char x0, xa, xy, xz, xr;
    GemmArgs a, b;
    a.a = &xa; a.b = &xa; a.bias = nullptr; a.residual = has_residual ? &xr : nullptr; a.y = &xy; a.M = m; a.N = n1; a.K = k1;
    a.act = 1; a.round_tf32 = 0; a.a0 = k0 ? &x0 : nullptr; a.K0 = k0;
    b.a = &xy; b.b = &xa; b.bias = nullptr; b.residual = nullptr; b.y = &xz; b.M = m; b.N = n2; b.K = n1; b.act = 1;
    b.round_tf32 = 0;
    *eligible = gemm_pair_eligible(a, b, precision, 0) ? 1 : 0;
    *nbuf = *pf = *na = *stages = *smem_bytes = 0;
    if (*eligible) {
      const int bk = precision == HFR_BF16 ? 64 : 32;
      gemm_pair_config((k0 + k1 + bk - 1) / bk, nbuf, pf, na, stages, smem_bytes);
    }
  });
}

int hfr_debug_knn_plan(int64_t nq, int64_t n, int* splits, int* n_blocks_per_unit) {
  return guarded([&] {
    if (!splits || !n_blocks_per_unit || nq <= 0 || n <= 0) throw Error(HFR_ERR_INVALID, "bad argument");
    knn_plan(nq, n, splits, n_blocks_per_unit);
  });
}

int hfr_l2_normalize(const float* x, float* y, int64_t n, int dim, int device, void* stream) {
  return guarded([&] {
    if (!x || !y || n < 0 || dim <= 0) throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(device);
    launch_l2norm(x, y, n, dim, (cudaStream_t)stream);
  });
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ 1-NN
struct hfr_knn {
  int device = 0, dim = 0, precision = HFR_BF16;
  const float* gallery = nullptr;  // borrowed fp32 rows
  int64_t n_local = 0, row_offset = 0;
  DevBuf g_lowp, g_norm, g_max2, q_lowp, part_score, part_idx;
  DevBuf unc_list, counters, locks;  // certification: queries for the exact pass, their count, per-query merge locks
  DevBuf h_q, h_out;                 // host-path staging
  int64_t last_nq = 0;
  int last_records = 0;              // candidate records per query of the last call (debugging)
  cudaStream_t last_stream = nullptr;
};

extern "C" {

int hfr_knn_create(int device, int dim, int precision, hfr_knn** out) {
  return guarded([&] {
    if (!out) throw Error(HFR_ERR_INVALID, "null argument");
    if (precision != HFR_TF32 && precision != HFR_BF16) throw Error(HFR_ERR_INVALID, "1-NN precision must be tf32 or bf16");
    if (dim <= 0 || dim % 8) throw Error(HFR_ERR_INVALID, "1-NN dimension must be a positive multiple of 8");
    use_device(device);
    hfr_knn* k = new hfr_knn();
    k->device = device;
    k->dim = dim;
    k->precision = precision;
    *out = k;
  });
}

int hfr_knn_set_gallery(hfr_knn* k, const float* gallery, int64_t n_local, int64_t global_row_offset, void* stream) {
  return guarded([&] {
    if (!k || !gallery || n_local <= 0) throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(k->device);
    k->gallery = gallery;
    k->n_local = n_local;
    k->row_offset = global_row_offset;
    k->g_norm.ensure((size_t)n_local * 4);
    k->g_max2.ensure(4);
    void* lowp = nullptr;
    if (k->precision == HFR_BF16) {
      k->g_lowp.ensure((size_t)n_local * k->dim * 2);
      lowp = k->g_lowp.p;
    }
    cuda_check(cudaMemsetAsync(k->g_max2.p, 0, 4, (cudaStream_t)stream), "cudaMemsetAsync");
    launch_rows_prep(gallery, lowp, (float*)k->g_norm.p, (float*)k->g_max2.p, n_local, k->dim, (cudaStream_t)stream);
  });
}

static int knn_query_impl(hfr_knn* k, const float* queries, int64_t nq, int n_neighbors, hfr_neighbor* out, void* stream,
                          int partial) {
  return guarded([&] {
    if (!k || !queries || !out || nq < 0) throw Error(HFR_ERR_INVALID, "bad argument");
    if (n_neighbors < 1 || n_neighbors > 4) throw Error(HFR_ERR_UNSUPPORTED, "k-NN on the GPU path supports 1 <= n_neighbors <= 4");
    if (!k->gallery) throw Error(HFR_ERR_STATE, "hfr_knn_query before hfr_knn_set_gallery (NotFittedError)");
    if (nq == 0) return;
    if (nq >= (1ll << 31)) throw Error(HFR_ERR_INVALID, "too many queries for one call");
    use_device(k->device);
    cudaStream_t s = (cudaStream_t)stream;
    int splits, per;
    knn_plan(nq, k->n_local, &splits, &per);
    const int cand = n_neighbors == 1 ? 2 : 4;  // candidates per (query, gallery split, epilogue warpgroup)
    k->part_score.ensure((size_t)nq * splits * 2 * cand * 4);
    k->part_idx.ensure((size_t)nq * splits * 2 * cand * 4);
    k->unc_list.ensure((size_t)nq * 4);
    k->locks.ensure((size_t)nq * 4);
    k->counters.ensure(16);
    cuda_check(cudaMemsetAsync(k->counters.p, 0, 16, s), "cudaMemsetAsync");
    cuda_check(cudaMemsetAsync(k->locks.p, 0, (size_t)nq * 4, s), "cudaMemsetAsync");
    KnnGemmArgs a;
    a.nq = nq; a.n = k->n_local; a.d = k->dim; a.splits = splits; a.n_blocks_per_unit = per;
    a.cand = cand;
    a.gnorm = (const float*)k->g_norm.p;
    a.part_score = (float*)k->part_score.p;
    a.part_idx = (int*)k->part_idx.p;
    if (k->precision == HFR_BF16) {
      k->q_lowp.ensure((size_t)nq * k->dim * 2);
      launch_rows_prep(queries, k->q_lowp.p, nullptr, nullptr, nq, k->dim, s);
      a.q = k->q_lowp.p;
      a.g = k->g_lowp.p;
    } else {
      a.q = queries;
      a.g = k->gallery;
    }
    launch_knn_gemm(a, k->precision, k->device, s);
    KnnFinalizeArgs f;
    f.q = queries; f.g = k->gallery; f.part_score = a.part_score; f.part_idx = a.part_idx; f.splits = splits;
    f.cand = cand; f.nq = nq; f.n = k->n_local; f.d = k->dim; f.row_offset = k->row_offset; f.k = n_neighbors;
    f.precision = k->precision; f.gmax2 = (const float*)k->g_max2.p; f.out = out;
    f.unc_list = (int*)k->unc_list.p; f.counters = (int*)k->counters.p; f.locks = (int*)k->locks.p;
    f.partial = partial;
    launch_knn_finalize(f, k->device, s);
    k->last_nq = nq;
    k->last_records = splits * 2 * cand;
    k->last_stream = s;
  });
}

int hfr_knn_query(hfr_knn* k, const float* queries, int64_t nq, int n_neighbors, hfr_neighbor* out, void* stream) {
  return knn_query_impl(k, queries, nq, n_neighbors, out, stream, 0);
}

int hfr_knn_query_partial(hfr_knn* k, const float* queries, int64_t nq, int n_neighbors, hfr_neighbor* out, void* stream) {
  return knn_query_impl(k, queries, nq, n_neighbors, out, stream, 1);
}

int hfr_knn_merge_certify(const hfr_neighbor* parts, int n_parts, int64_t nq, int n_neighbors, hfr_neighbor* out,
                          int32_t* unc_list, int32_t* unc_count, int device, void* stream) {
  return guarded([&] {
    if (!parts || !out || !unc_list || !unc_count || n_parts <= 0 || nq < 0 || n_neighbors < 1 || n_neighbors > 4)
      throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(device);
    launch_knn_merge_certify(parts, n_parts, nq, n_neighbors, out, unc_list, unc_count, (cudaStream_t)stream);
  });
}

int hfr_knn_query_exact(hfr_knn* k, const float* queries, int64_t nq, int n_neighbors, const int32_t* unc_list,
                        const int32_t* unc_count, hfr_neighbor* out, void* stream) {
  return guarded([&] {
    if (!k || !queries || !out || !unc_list || !unc_count || nq <= 0 || n_neighbors < 1 || n_neighbors > 4)
      throw Error(HFR_ERR_INVALID, "bad argument");
    if (!k->gallery) throw Error(HFR_ERR_STATE, "hfr_knn_query_exact before hfr_knn_set_gallery (NotFittedError)");
    use_device(k->device);
    cudaStream_t s = (cudaStream_t)stream;
    k->locks.ensure((size_t)nq * 4);
    cuda_check(cudaMemsetAsync(k->locks.p, 0, (size_t)nq * 4, s), "cudaMemsetAsync");
    launch_knn_exact_listed(queries, (const float*)k->gallery, k->n_local, k->dim, k->row_offset, n_neighbors, unc_list,
                            unc_count, (int*)k->locks.p, out, k->device, s);
  });
}

int hfr_knn_merge_listed(const hfr_neighbor* parts, int n_parts, int64_t nq, int n_neighbors, const int32_t* unc_list,
                         const int32_t* unc_count, hfr_neighbor* out, int device, void* stream) {
  return guarded([&] {
    if (!parts || !out || !unc_list || !unc_count || n_parts <= 0 || nq < 0 || n_neighbors < 1 || n_neighbors > 4)
      throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(device);
    launch_knn_merge_listed(parts, n_parts, nq, n_neighbors, unc_list, unc_count, out, (cudaStream_t)stream);
  });
}

int hfr_knn_query_host(hfr_knn* k, const float* queries_host, int64_t nq, int n_neighbors, hfr_neighbor* out_host,
                       void* stream) {
  int rc = guarded([&] {
    if (!k || !queries_host || !out_host || nq <= 0 || n_neighbors < 1 || n_neighbors > 4)
      throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(k->device);
    k->h_q.ensure((size_t)nq * k->dim * 4);
    k->h_out.ensure((size_t)nq * n_neighbors * sizeof(hfr_neighbor));
    cuda_check(cudaMemcpyAsync(k->h_q.p, queries_host, (size_t)nq * k->dim * 4, cudaMemcpyHostToDevice,
                               (cudaStream_t)stream), "cudaMemcpyAsync(H2D)");
  });
  if (rc) return rc;
  rc = knn_query_impl(k, (const float*)k->h_q.p, nq, n_neighbors, (hfr_neighbor*)k->h_out.p, stream, 0);
  if (rc) return rc;
  return guarded([&] {
    cudaStream_t s = (cudaStream_t)stream;
    cuda_check(cudaMemcpyAsync(out_host, k->h_out.p, (size_t)nq * n_neighbors * sizeof(hfr_neighbor),
                               cudaMemcpyDeviceToHost, s), "D2H");
    cuda_check(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  });
}

int hfr_knn_merge(const hfr_neighbor* parts, int n_parts, int64_t nq, int n_neighbors, hfr_neighbor* out, int device,
                  void* stream) {
  return guarded([&] {
    if (!parts || !out || n_parts <= 0 || nq < 0 || n_neighbors < 1 || n_neighbors > 4)
      throw Error(HFR_ERR_INVALID, "bad argument");
    use_device(device);
    launch_knn_merge(parts, n_parts, nq, n_neighbors, out, (cudaStream_t)stream);
  });
}

int hfr_knn_stats(hfr_knn* k, int64_t* certified, int64_t* rescored) {
  return guarded([&] {
    if (!k) throw Error(HFR_ERR_INVALID, "null handle");
    int64_t unc = 0;
    if (k->last_nq > 0) {
      use_device(k->device);
      int c = 0;
      cuda_check(cudaStreamSynchronize(k->last_stream), "cudaStreamSynchronize");
      cuda_check(cudaMemcpy(&c, k->counters.p, 4, cudaMemcpyDeviceToHost), "cudaMemcpy(counters)");
      unc = c;
    }
    if (certified) *certified = k->last_nq - unc;
    if (rescored) *rescored = unc;
  });
}

int64_t hfr_knn_debug_candidates(hfr_knn* k, float* score_host, int32_t* index_host, int64_t nq) {
  int64_t rec = 0;
  int rc = guarded([&] {
    if (!k) throw Error(HFR_ERR_INVALID, "null handle");
    rec = k->last_records;
    if (!score_host && !index_host) return;
    if (nq != k->last_nq) throw Error(HFR_ERR_STATE, "nq does not match the last query");
    use_device(k->device);
    cuda_check(cudaStreamSynchronize(k->last_stream), "cudaStreamSynchronize");
    const size_t bytes = (size_t)nq * rec * 4;
    if (score_host) cuda_check(cudaMemcpy(score_host, k->part_score.p, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy");
    if (index_host) cuda_check(cudaMemcpy(index_host, k->part_idx.p, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy");
  });
  return rc < 0 ? rc : rec;
}

void hfr_knn_free(hfr_knn* k) { delete k; }

// ------------------------------------------------------------------------------------------------ single operators
int hfr_op_dwconv3x3(const void* x, const float* w9c, const float* bias, void* y, int batch, int h, int w, int c,
                     int stride, int pad_t, int pad_l, int ho, int wo, int act, int dtype, int device, void* stream) {
  return guarded([&] {
    use_device(device);
    DwArgs a;
    a.x = x; a.w = w9c; a.bias = bias; a.y = y; a.B = batch; a.H = h; a.W = w; a.C = c; a.Ho = ho; a.Wo = wo;
    a.stride = stride; a.pad_t = pad_t; a.pad_l = pad_l; a.act = act; a.round_tf32 = 0;
    launch_dw(a, dtype, (cudaStream_t)stream);
  });
}

int hfr_op_gemm_bias_act(const void* a_, const void* b, const float* bias, const void* residual, void* y, int64_t m,
                         int n, int k, int act, int dtype, int device, void* stream) {
  return guarded([&] {
    use_device(device);
    if (dtype == HFR_FP32 && residual == nullptr && k % 4 == 0) {
      // fp32 operands: the tensor cores at fp32-level accuracy (3xTF32) - the PCA projection of the classifier list
      launch_gemm_x3((const float*)a_, (const float*)b, bias, (float*)y, m, n, k, n, act, device, (cudaStream_t)stream);
      return;
    }
    GemmArgs a;
    a.a = a_; a.b = b; a.bias = bias; a.residual = residual; a.y = y; a.M = m; a.N = n; a.K = k; a.act = act;
    a.round_tf32 = 0;
    launch_gemm(a, dtype, device, (cudaStream_t)stream);
  });
}

int hfr_op_gemm_pair(const void* a0, int k0, const void* a_, int k1, const void* w1, const float* bias1, const void* residual,
                     void* y, int64_t m, int n1, int act1, const void* w2, const float* bias2, void* z, int n2, int act2,
                     int dtype, int device, void* stream) {
  return guarded([&] {
    use_device(device);
    if (!a_ || !w1 || !y) throw Error(HFR_ERR_INVALID, "null argument");
    GemmArgs a;
    a.a = a_; a.b = w1; a.bias = bias1; a.residual = residual; a.y = y; a.M = m; a.N = n1; a.K = k1; a.act = act1;
    a.round_tf32 = (dtype == HFR_TF32 && z != nullptr);   // y feeds the second tf32 GEMM
    a.a0 = a0; a.K0 = a0 ? k0 : 0;
    if (!z) {
      launch_gemm(a, dtype, device, (cudaStream_t)stream);
      return;
    }
    GemmArgs b;
    b.a = y; b.b = w2; b.bias = bias2; b.residual = nullptr; b.y = z; b.M = m; b.N = n2; b.K = n1; b.act = act2;
    b.round_tf32 = 0;
    if (!gemm_pair_eligible(a, b, dtype, device)) throw Error(HFR_ERR_UNSUPPORTED, "gemm pair: shapes not eligible");
    launch_gemm_pair(a, b, dtype, device, (cudaStream_t)stream);
  });
}

int hfr_op_stem_conv(const void* x, int in_dtype, const float* w, const float* bias, void* y, int batch, int h, int w_,
                     int kh, int kw, int stride, int pad_t, int pad_l, int ho, int wo, int cout, int flags, int act,
                     int dtype, int device, void* stream) {
  return guarded([&] {
    use_device(device);
    StemArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.in_u8 = (in_dtype == HFR_IN_U8); a.w = w; a.bias = bias; a.y = y; a.B = batch; a.H = h; a.W = w_;
    a.Ho = ho; a.Wo = wo; a.kh = kh; a.kw = kw; a.stride = stride; a.pad_t = pad_t; a.pad_l = pad_l; a.cout = cout;
    a.scale = 1.f;
    if (a.in_u8) {
      a.flip = (flags & HFR_FLAG_BGR) ? 1 : 0;
      if (flags & HFR_FLAG_MEAN_IMAGENET) { a.mean[0] = 103.939f; a.mean[1] = 116.779f; a.mean[2] = 123.68f; }
      if (flags & HFR_FLAG_MEAN_VGGFACE2) { a.mean[0] = 91.4953f; a.mean[1] = 103.8827f; a.mean[2] = 131.0912f; }
      if (flags & HFR_FLAG_SCALE_PM1) { a.scale = 1.f / 127.5f; a.mean[0] = a.mean[1] = a.mean[2] = 1.f; }
    }
    a.act = act;
    launch_stem(a, dtype, (cudaStream_t)stream);
  });
}

int hfr_op_stem_conv_tc(const void* x_u8, const float* w_host, const float* bias, void* y, int batch, int h, int w_,
                        int kh, int kw, int pad_t, int pad_l, int ho, int wo, int cout, int flags, int act, int device,
                        void* stream) {
  return guarded([&] {
    use_device(device);
    int flip = (flags & HFR_FLAG_BGR) ? 1 : 0;
    float scale = 1.f, mean[3] = {0.f, 0.f, 0.f};
    if (flags & HFR_FLAG_MEAN_IMAGENET) { mean[0] = 103.939f; mean[1] = 116.779f; mean[2] = 123.68f; }
    if (flags & HFR_FLAG_MEAN_VGGFACE2) { mean[0] = 91.4953f; mean[1] = 103.8827f; mean[2] = 131.0912f; }
    if (flags & HFR_FLAG_SCALE_PM1) { scale = 1.f / 127.5f; mean[0] = mean[1] = mean[2] = 1.f; }
    if ((h % 2) || (w_ % 2) || (cout != 32 && cout != 64)) throw Error(HFR_ERR_UNSUPPORTED, "tensor-core stem: even H/W, cout 32|64");
    std::vector<float> wv(w_host, w_host + (size_t)kh * kw * 3 * cout);
    StemTcGeom g;
    std::vector<uint16_t> w2 = stem_tc_weights(wv, kh, kw, cout, pad_t, pad_l, flip, scale, mean, &g);
    DevBuf scratch;
    scratch.ensure((size_t)batch * (ho + g.ka - 1) * (wo + g.kb - 1) * 32);
    const bool win = window_enabled();
    if (win) w2 = stem_window_layout(w2, g.ka * g.kb, cout);
    void* w2d = upload(w2.data(), w2.size() * 2);
    StemTcArgs t;
    t.use_window = win;
    t.x = (const uint8_t*)x_u8; t.scratch = scratch.p; t.w2 = w2d; t.bias = bias; t.y = y;
    t.B = batch; t.H = h; t.W = w_; t.Ho = ho; t.Wo = wo; t.cout = cout;
    t.ka = g.ka; t.kb = g.kb; t.pt2 = g.pt2; t.pl2 = g.pl2; t.act = act;
    try {
      launch_stem_tc(t, device, (cudaStream_t)stream);
      cuda_check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
    } catch (...) {
      cudaFree(w2d);
      throw;
    }
    cudaFree(w2d);
  });
}

int hfr_op_conv2d_window(const void* x, const float* w_host, const float* bias, void* y, int batch, int h, int w_,
                         int cin, int kh, int kw, int pad_t, int pad_l, int ho, int wo, int cout, int act, int device,
                         void* stream) {
  return guarded([&] {
    use_device(device);
    std::vector<uint16_t> hw = conv_window_layout(w_host, cout, kh * kw, cin);
    void* wd = upload(hw.data(), hw.size() * 2);
    WinArgs a;
    a.x = x; a.w = wd; a.bias = bias; a.y = y; a.B = batch; a.H = h; a.W = w_; a.cin = cin; a.Ho = ho; a.Wo = wo;
    a.cout = cout; a.kh = kh; a.kw = kw; a.pad_t = pad_t; a.pad_l = pad_l; a.act = act; a.plane_major = 0;
    try {
      launch_conv_window(a, device, (cudaStream_t)stream);
      cuda_check(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize");
    } catch (...) {
      cudaFree(wd);
      throw;
    }
    cudaFree(wd);
  });
}

int hfr_op_conv2d(const void* x, const void* w, const float* bias, const void* residual, void* y, int batch, int h,
                  int w_, int cin, int kh, int kw, int stride, int pad_t, int pad_l, int ho, int wo, int cout, int act,
                  int dtype, int device, void* stream) {
  return guarded([&] {
    use_device(device);
    ConvArgs a;
    a.x = x; a.w = w; a.bias = bias; a.residual = residual; a.y = y; a.B = batch; a.H = h; a.W = w_; a.cin = cin;
    a.Ho = ho; a.Wo = wo; a.cout = cout; a.kh = kh; a.kw = kw; a.stride = stride; a.pad_t = pad_t; a.pad_l = pad_l;
    a.dil = 1; a.act = act; a.round_tf32 = 0;
    launch_conv(a, dtype, device, (cudaStream_t)stream);
  });
}

int hfr_op_maxpool(const void* x, void* y, int batch, int h, int w, int c, int k, int stride, int pad_t, int pad_l,
                   int ho, int wo, int explicit_zero, int dtype, int device, void* stream) {
  return guarded([&] {
    use_device(device);
    PoolArgs a;
    a.x = x; a.y = y; a.B = batch; a.H = h; a.W = w; a.C = c; a.Ho = ho; a.Wo = wo; a.k = k; a.stride = stride;
    a.pad_t = pad_t; a.pad_l = pad_l; a.explicit_zero = explicit_zero;
    launch_maxpool(a, dtype, (cudaStream_t)stream);
  });
}

}  // extern "C"
