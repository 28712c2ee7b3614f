// Graph compiler: frozen TF GraphDef -> fused layer plan.
//
//   phase 1  name-resolved, memoised lowering from the requested outputs (nodes are NOT in topological order in the
//            reference's files): constants are folded on the host (Dequantize MIN_FIRST, BN arithmetic, Rsqrt ...),
//            Switch/Merge with a constant predicate are resolved by dead-value propagation, Identity/Reshape/Dropout
//            disappear.  What remains is a small SSA list of primitive ops over activation values.
//   phase 2  conv/depthwise/matmul ops absorb their single-consumer tails:  -> scale -> shift -> (+ residual)
//            -> relu | relu,min(6),max(0) | softmax | sigmoid.  BN scales may be negative: they are folded into the
//            weights *before* the activation only, never moved across it.
//   phase 3  layers are ordered topologically and value lifetimes recorded for the activation arena.
#include <algorithm>
#include <climits>
#include <cmath>
#include <functional>
#include <map>
#include <sstream>

#include "graph.h"

namespace hfr {
namespace {

[[noreturn]] void fail(const std::string& msg) { throw std::runtime_error("compile: " + msg); }

using TensorP = std::shared_ptr<HTensor>;

struct Sym {
  enum K { DEAD, CONST, ACT } k = DEAD;
  TensorP c;
  int vid = -1;
};

struct IOp {
  enum Kind { CONV, DWCONV, MATMUL, SCALE, SHIFT, RELU, RELU6, MINC, MAXC, ADD2, MAXPOOL, GAP, SOFTMAX, SIGMOID } kind;
  std::string name;
  int in0 = -1, in1 = -1, out = -1;
  TensorP c;        // weights / per-channel vector
  float scalar = 0.f;
  int kh = 1, kw = 1, stride = 1, dil = 1;
  int pad_t = 0, pad_b = 0, pad_l = 0, pad_r = 0;
  bool explicit_zero = false;
};

struct Val {
  int H = 0, W = 0, C = 0;
  bool vec = false;
};

struct PadInfo {
  int src;
  int t, b, l, r;
};

void split_ref(const std::string& ref, std::string* name, int* port) {
  size_t colon = ref.rfind(':');
  if (colon != std::string::npos && colon + 1 < ref.size() &&
      std::all_of(ref.begin() + colon + 1, ref.end(), [](char ch) { return ch >= '0' && ch <= '9'; })) {
    *name = ref.substr(0, colon);
    *port = std::stoi(ref.substr(colon + 1));
  } else {
    *name = ref;
    *port = 0;
  }
}

// numpy-style broadcasting binary op on host tensors
TensorP broadcast_binary(const HTensor& a, const HTensor& b, const std::function<float(float, float)>& fn) {
  size_t rank = std::max(a.shape.size(), b.shape.size());
  std::vector<int64_t> sa(rank, 1), sb(rank, 1), so(rank, 1);
  std::copy(a.shape.begin(), a.shape.end(), sa.begin() + (rank - a.shape.size()));
  std::copy(b.shape.begin(), b.shape.end(), sb.begin() + (rank - b.shape.size()));
  for (size_t i = 0; i < rank; ++i) {
    if (sa[i] != sb[i] && sa[i] != 1 && sb[i] != 1) fail("constant folding: incompatible broadcast shapes");
    so[i] = std::max(sa[i], sb[i]);
  }
  auto out = std::make_shared<HTensor>();
  out->dtype = a.dtype;
  out->shape = so;
  int64_t n = out->numel();
  out->f.resize((size_t)n);
  std::vector<int64_t> idx(rank, 0);
  for (int64_t lin = 0; lin < n; ++lin) {
    int64_t ia = 0, ib = 0;
    for (size_t d = 0; d < rank; ++d) {
      ia = ia * sa[d] + (sa[d] == 1 ? 0 : idx[d]);
      ib = ib * sb[d] + (sb[d] == 1 ? 0 : idx[d]);
    }
    out->f[(size_t)lin] = fn(a.f[(size_t)ia], b.f[(size_t)ib]);
    for (int d = (int)rank - 1; d >= 0; --d) {
      if (++idx[d] < so[d]) break;
      idx[d] = 0;
    }
  }
  return out;
}

// TF Dequantize, mode MIN_FIRST, T=quint8, evaluated in fp32 like TF's Eigen kernel:
//   s = (max - min) / 255 ;  w = round(min / s) * s + q * s      
TensorP dequantize_min_first(const HTensor& q, float mn, float mx) {
  auto out = std::make_shared<HTensor>();
  out->dtype = 1;
  out->shape = q.shape;
  out->f.resize(q.f.size());
  const float s = (mx - mn) / 255.0f;
  const float off = std::nearbyintf(mn / s) * s;
  for (size_t i = 0; i < q.f.size(); ++i) out->f[i] = q.f[i] * s + off;
  return out;
}

void same_pad(int size, int k, int s, int d, int* out, int* before, int* after) {
  *out = (size + s - 1) / s;
  int eff = (k - 1) * d + 1;
  int total = std::max((*out - 1) * s + eff - size, 0);
  *before = total / 2;
  *after = total - total / 2;
}

struct Lowerer {
  const Graph& g;
  const CompileOptions& opt;
  std::unordered_map<std::string, std::vector<Sym>> memo;
  std::vector<IOp> ops;
  std::vector<Val> vals;
  std::unordered_map<int, PadInfo> pads;  // virtual values produced by Pad
  std::string input_node, phase_node;
  int depth = 0;

  Lowerer(const Graph& g_, const CompileOptions& o) : g(g_), opt(o) {
    int port;
    split_ref(opt.input_name, &input_node, &port);
    if (!opt.phase_name.empty()) split_ref(opt.phase_name, &phase_node, &port);
  }

  int new_val(const Val& v) {
    vals.push_back(v);
    return (int)vals.size() - 1;
  }

  Sym get(const std::string& ref) {
    std::string name;
    int port;
    split_ref(ref, &name, &port);
    auto it = memo.find(name);
    if (it == memo.end()) {
      const GNode* n = g.find(name);
      if (!n) fail("node '" + name + "' not found in graph");
      if (++depth > 4000) fail("graph too deep (cycle?)");
      std::vector<Sym> r = lower(*n);
      --depth;
      it = memo.emplace(name, std::move(r)).first;
    }
    if (port >= (int)it->second.size()) fail("node '" + name + "' has no output port " + std::to_string(port));
    return it->second[port];
  }

  static Sym constant(TensorP t) {
    Sym s;
    s.k = Sym::CONST;
    s.c = std::move(t);
    return s;
  }
  static Sym act(int vid) {
    Sym s;
    s.k = Sym::ACT;
    s.vid = vid;
    return s;
  }

  std::vector<Sym> data_inputs(const GNode& n) {
    std::vector<Sym> r;
    for (auto& in : n.inputs)
      if (!in.empty() && in[0] != '^') r.push_back(get(in));
    return r;
  }

  // per-channel view of a constant operand combined with an activation of C channels
  TensorP channel_vector(const HTensor& c, int C, const std::string& where) {
    auto out = std::make_shared<HTensor>();
    out->shape = {C};
    if (c.numel() == 1) {
      out->f.assign((size_t)C, c.f[0]);
    } else if (c.numel() == C && !c.shape.empty() && c.shape.back() == C) {
      out->f = c.f;
    } else {
      fail(where + ": constant operand is neither a scalar nor a per-channel vector");
    }
    return out;
  }

  int resolve_pad(int vid, IOp* op) {
    auto it = pads.find(vid);
    if (it == pads.end()) return vid;
    op->pad_t += it->second.t;
    op->pad_b += it->second.b;
    op->pad_l += it->second.l;
    op->pad_r += it->second.r;
    op->explicit_zero = true;
    return it->second.src;
  }

  Sym emit_unary(IOp::Kind kind, const GNode& n, const Sym& x) {
    if (pads.count(x.vid)) fail(n.name + ": Pad must feed a convolution or pooling op");
    IOp op;
    op.kind = kind;
    op.name = n.name;
    op.in0 = x.vid;
    op.out = new_val(vals[x.vid]);
    ops.push_back(op);
    return act(op.out);
  }

  Sym emit_affine(IOp::Kind kind, const GNode& n, const Sym& x, const HTensor& c, float mul) {
    if (pads.count(x.vid)) fail(n.name + ": Pad must feed a convolution or pooling op");
    IOp op;
    op.kind = kind;
    op.name = n.name;
    op.in0 = x.vid;
    op.c = channel_vector(c, vals[x.vid].C, n.name);
    if (mul != 1.f)
      for (auto& v : op.c->f) v *= mul;
    op.out = new_val(vals[x.vid]);
    ops.push_back(op);
    return act(op.out);
  }

  std::vector<Sym> lower(const GNode& n) {
    const std::string& op = n.op;
    // ---- sources
    if (n.name == input_node) {
      if (!vals.empty()) fail("internal: the input placeholder must be lowered first");
      Val v;
      const AttrVal* sh = n.attr("shape");
      if (sh && sh->shape.size() == 4) {
        v.H = (int)sh->shape[1];
        v.W = (int)sh->shape[2];
        v.C = (int)sh->shape[3];
      }
      if (opt.override_hw > 0) v.H = v.W = opt.override_hw;
      if (v.H <= 0 || v.W <= 0) v.H = v.W = 160;  // facerec_test.py:66-67 fallback for unknown shapes
      if (v.C <= 0) v.C = 3;
      if (v.C != 3) fail("input placeholder must have 3 channels");
      vals.push_back(v);
      return {act(0)};
    }
    if (!phase_node.empty() && n.name == phase_node) {
      auto t = std::make_shared<HTensor>();
      t->dtype = 10;
      t->f = {opt.phase_value};
      return {constant(t)};
    }
    if (op == "Const") {
      const AttrVal* v = n.attr("value");
      if (!v || v->kind != AttrVal::TENSOR) fail("Const '" + n.name + "' without a value");
      return {constant(std::make_shared<HTensor>(v->tensor))};
    }
    if (op == "Placeholder") fail("placeholder '" + n.name + "' is neither the input nor the learning-phase tensor");

    // ---- control flow (Keras learning-phase conditionals around BN)
    if (op == "Merge") {
      for (auto& in : n.inputs) {
        if (in.empty() || in[0] == '^') continue;
        Sym s = get(in);
        if (s.k != Sym::DEAD) return {s, Sym()};
      }
      return {Sym(), Sym()};
    }
    std::vector<Sym> a = data_inputs(n);
    if (op == "Switch") {
      if (a.size() != 2) fail("Switch '" + n.name + "' expects 2 inputs");
      if (a[0].k == Sym::DEAD || a[1].k == Sym::DEAD) return {Sym(), Sym()};
      if (a[1].k != Sym::CONST) fail("Switch '" + n.name + "': predicate is not constant (pass learning_phase_tensor)");
      bool pred = a[1].c->f[0] != 0.f;
      return pred ? std::vector<Sym>{Sym(), a[0]} : std::vector<Sym>{a[0], Sym()};
    }
    for (auto& s : a)
      if (s.k == Sym::DEAD) return {Sym(), Sym(), Sym(), Sym(), Sym(), Sym()};

    auto all_const = [&]() { return std::all_of(a.begin(), a.end(), [](const Sym& s) { return s.k == Sym::CONST; }); };

    if (op == "Identity" || op == "StopGradient" || op == "PlaceholderWithDefault" || op == "Dropout") return {a.at(0)};

    if (op == "Dequantize") {
      const AttrVal* mode = n.attr("mode");
      if (!mode || mode->s != "MIN_FIRST") fail("Dequantize '" + n.name + "': only mode=MIN_FIRST is supported");
      if (!all_const() || a.size() != 3) fail("Dequantize '" + n.name + "': non-constant input");
      return {constant(dequantize_min_first(*a[0].c, a[1].c->f[0], a[2].c->f[0]))};
    }

    // ---- element-wise arithmetic: fold if constant, otherwise scale/shift of an activation
    if (op == "Add" || op == "AddV2" || op == "BiasAdd" || op == "Sub" || op == "Mul" || op == "RealDiv" ||
        op == "Maximum" || op == "Minimum") {
      if (a.size() != 2) fail(op + " '" + n.name + "' expects 2 inputs");
      if (all_const()) {
        std::function<float(float, float)> fn;
        if (op == "Sub") fn = [](float x, float y) { return x - y; };
        else if (op == "Mul") fn = [](float x, float y) { return x * y; };
        else if (op == "RealDiv") fn = [](float x, float y) { return x / y; };
        else if (op == "Maximum") fn = [](float x, float y) { return std::max(x, y); };
        else if (op == "Minimum") fn = [](float x, float y) { return std::min(x, y); };
        else fn = [](float x, float y) { return x + y; };
        return {constant(broadcast_binary(*a[0].c, *a[1].c, fn))};
      }
      if (a[0].k == Sym::ACT && a[1].k == Sym::ACT) {
        if (op != "Add" && op != "AddV2") fail(op + " '" + n.name + "' of two activations is not supported");
        const Val &v0 = vals[a[0].vid], &v1 = vals[a[1].vid];
        if (v0.H != v1.H || v0.W != v1.W || v0.C != v1.C) fail("residual add '" + n.name + "': shape mismatch");
        IOp o;
        o.kind = IOp::ADD2;
        o.name = n.name;
        o.in0 = a[0].vid;
        o.in1 = a[1].vid;
        o.out = new_val(v0);
        ops.push_back(o);
        return {act(o.out)};
      }
      const bool act_first = a[0].k == Sym::ACT;
      const Sym& x = act_first ? a[0] : a[1];
      const HTensor& c = act_first ? *a[1].c : *a[0].c;
      if (op == "Add" || op == "AddV2" || op == "BiasAdd") return {emit_affine(IOp::SHIFT, n, x, c, 1.f)};
      if (op == "Mul") return {emit_affine(IOp::SCALE, n, x, c, 1.f)};
      if (op == "Sub") {
        if (!act_first) fail("Sub '" + n.name + "': constant - activation is not supported");
        return {emit_affine(IOp::SHIFT, n, x, c, -1.f)};
      }
      if (op == "RealDiv") {
        if (!act_first) fail("RealDiv '" + n.name + "': constant / activation is not supported");
        auto inv = std::make_shared<HTensor>(c);
        for (auto& v : inv->f) v = 1.f / v;
        return {emit_affine(IOp::SCALE, n, x, *inv, 1.f)};
      }
      if (c.numel() != 1) fail(op + " '" + n.name + "': only scalar bounds are supported");
      Sym r = emit_unary(op == "Maximum" ? IOp::MAXC : IOp::MINC, n, x);
      ops.back().scalar = c.f[0];
      return {r};
    }
    if (op == "Rsqrt" || op == "Sqrt" || op == "Neg") {
      if (!all_const()) fail(op + " '" + n.name + "' on an activation is not supported");
      auto t = std::make_shared<HTensor>(*a[0].c);
      for (auto& v : t->f) v = op == "Rsqrt" ? 1.f / std::sqrt(v) : op == "Sqrt" ? std::sqrt(v) : -v;
      return {constant(t)};
    }
    if (op == "Reshape" || op == "Squeeze" || op == "ExpandDims" || op == "Flatten") {
      if (a[0].k == Sym::CONST) {
        auto t = std::make_shared<HTensor>(*a[0].c);
        if (op == "Reshape" && a.size() > 1) {
          t->shape.clear();
          int64_t known = 1, neg = -1;
          for (size_t i = 0; i < a[1].c->f.size(); ++i) {
            int64_t d = (int64_t)a[1].c->f[i];
            t->shape.push_back(d);
            if (d < 0) neg = (int64_t)i; else known *= d;
          }
          if (neg >= 0) t->shape[(size_t)neg] = (int64_t)t->f.size() / std::max<int64_t>(known, 1);
        }
        return {constant(t)};
      }
      // activation: only layout-preserving reshapes of pooled vectors ([B,C] <-> [B,1,1,C]) occur on this path
      const Val& v = vals[a[0].vid];
      if (!(v.vec || (v.H == 1 && v.W == 1))) fail(op + " '" + n.name + "': only reshapes of pooled [B,C] features are supported");
      return {a[0]};
    }
    if (op == "Relu") return {emit_unary(IOp::RELU, n, a.at(0))};
    if (op == "Relu6") return {emit_unary(IOp::RELU6, n, a.at(0))};
    if (op == "Softmax") return {emit_unary(IOp::SOFTMAX, n, a.at(0))};
    if (op == "Sigmoid") return {emit_unary(IOp::SIGMOID, n, a.at(0))};

    if (op == "FusedBatchNorm" || op == "FusedBatchNormV2" || op == "FusedBatchNormV3") {
      if (a.size() != 5 || a[0].k != Sym::ACT) fail("FusedBatchNorm '" + n.name + "': unexpected inputs");
      const AttrVal* tr = n.attr("is_training");
      if (tr && tr->b) fail("FusedBatchNorm '" + n.name + "' is in training mode");
      const AttrVal* ea = n.attr("epsilon");
      const float eps = ea ? ea->f : 1e-3f;
      const int C = vals[a[0].vid].C;
      for (int i = 1; i < 5; ++i)   // untrusted file: the four parameter vectors must be constants of C elements
        if (a[(size_t)i].k != Sym::CONST || !a[(size_t)i].c || (int64_t)a[(size_t)i].c->f.size() < C)
          fail("FusedBatchNorm '" + n.name + "': scale / offset / mean / variance must be constant vectors of " +
               std::to_string(C) + " elements");
      HTensor sc, sh;
      sc.shape = sh.shape = {C};
      sc.f.resize((size_t)C);
      sh.f.resize((size_t)C);
      for (int i = 0; i < C; ++i) {
        const float s = a[1].c->f[(size_t)i] / std::sqrt(a[4].c->f[(size_t)i] + eps);
        sc.f[(size_t)i] = s;
        sh.f[(size_t)i] = a[2].c->f[(size_t)i] - a[3].c->f[(size_t)i] * s;
      }
      Sym s1 = emit_affine(IOp::SCALE, n, a[0], sc, 1.f);
      Sym s2 = emit_affine(IOp::SHIFT, n, s1, sh, 1.f);
      return {s2, Sym(), Sym(), Sym(), Sym(), Sym()};
    }

    if (op == "Pad") {
      if (a[0].k != Sym::ACT || a[1].k != Sym::CONST || a[1].c->f.size() != 8) fail("Pad '" + n.name + "': unsupported");
      const auto& p = a[1].c->f;
      if (p[0] != 0 || p[1] != 0 || p[6] != 0 || p[7] != 0) fail("Pad '" + n.name + "': only spatial padding is supported");
      if (pads.count(a[0].vid)) fail("Pad of Pad");
      Val v = vals[a[0].vid];
      PadInfo pi{a[0].vid, (int)p[2], (int)p[3], (int)p[4], (int)p[5]};
      v.H += pi.t + pi.b;
      v.W += pi.l + pi.r;
      int vid = new_val(v);
      pads[vid] = pi;
      return {act(vid)};
    }

    if (op == "Conv2D" || op == "DepthwiseConv2dNative") {
      if (a.size() != 2 || a[0].k != Sym::ACT || a[1].k != Sym::CONST) fail(op + " '" + n.name + "': expects (activation, constant kernel)");
      const AttrVal* df = n.attr("data_format");
      if (df && df->s != "NHWC") fail(op + " '" + n.name + "': only NHWC is supported");
      const HTensor& w = *a[1].c;
      if (w.shape.size() != 4) fail(op + " '" + n.name + "': kernel must be rank 4");
      if ((int64_t)w.f.size() != w.numel() || w.numel() <= 0) fail(op + " '" + n.name + "': kernel data does not match its shape");
      IOp o;
      o.kind = op == "Conv2D" ? IOp::CONV : IOp::DWCONV;
      o.name = n.name;
      o.c = a[1].c;
      o.kh = (int)w.shape[0];
      o.kw = (int)w.shape[1];
      const AttrVal* st = n.attr("strides");
      if (!st || st->shape.size() != 4 || st->shape[1] != st->shape[2]) fail(op + " '" + n.name + "': bad strides");
      o.stride = (int)st->shape[1];
      const AttrVal* dl = n.attr("dilations");
      o.dil = (dl && dl->shape.size() == 4) ? (int)dl->shape[1] : 1;
      const Val vin = vals[a[0].vid];  // (padded dims when fed by Pad)
      o.in0 = resolve_pad(a[0].vid, &o);
      if ((int)w.shape[2] != vin.C) fail(op + " '" + n.name + "': kernel input channels do not match the activation");
      if (o.kind == IOp::DWCONV && w.shape[3] != 1) fail("DepthwiseConv2dNative '" + n.name + "': channel multiplier must be 1");
      const AttrVal* pd = n.attr("padding");
      Val vo;
      vo.C = o.kind == IOp::CONV ? (int)w.shape[3] : vin.C;
      if (pd && pd->s == "SAME") {
        int bt, at, bl, al;
        same_pad(vin.H, o.kh, o.stride, o.dil, &vo.H, &bt, &at);
        same_pad(vin.W, o.kw, o.stride, o.dil, &vo.W, &bl, &al);
        o.pad_t += bt; o.pad_b += at; o.pad_l += bl; o.pad_r += al;
      } else {
        vo.H = (vin.H - ((o.kh - 1) * o.dil + 1)) / o.stride + 1;
        vo.W = (vin.W - ((o.kw - 1) * o.dil + 1)) / o.stride + 1;
      }
      if (vo.H <= 0 || vo.W <= 0) fail(op + " '" + n.name + "': empty output");
      o.out = new_val(vo);
      ops.push_back(o);
      return {act(o.out)};
    }

    if (op == "MaxPool" || op == "AvgPool" || op == "Mean") {
      if (a[0].k != Sym::ACT) fail(op + " '" + n.name + "': expects an activation");
      const Val vin = vals[a[0].vid];
      IOp o;
      o.name = n.name;
      if (op == "Mean") {
        if (a.size() != 2 || a[1].k != Sym::CONST) fail("Mean '" + n.name + "': axes must be constant");
        std::vector<int> axes;
        for (float f : a[1].c->f) axes.push_back((int)f);
        std::sort(axes.begin(), axes.end());
        if (axes != std::vector<int>{1, 2}) fail("Mean '" + n.name + "': only reduction over axes [1,2] is supported");
        if (pads.count(a[0].vid)) fail("Mean of Pad");
        o.kind = IOp::GAP;
        o.in0 = a[0].vid;
        Val vo;
        vo.C = vin.C;
        vo.vec = true;
        vo.H = vo.W = 1;
        o.out = new_val(vo);
        ops.push_back(o);
        return {act(o.out)};
      }
      const AttrVal* ks = n.attr("ksize");
      const AttrVal* st = n.attr("strides");
      const AttrVal* pd = n.attr("padding");
      if (!ks || ks->shape.size() != 4 || !st || st->shape.size() != 4) fail(op + " '" + n.name + "': bad ksize/strides");
      o.kh = (int)ks->shape[1];
      o.kw = (int)ks->shape[2];
      o.stride = (int)st->shape[1];
      o.in0 = resolve_pad(a[0].vid, &o);
      Val vo;
      vo.C = vin.C;
      const bool same = pd && pd->s == "SAME";
      if (same) {
        int bt, at, bl, al;
        same_pad(vin.H, o.kh, o.stride, 1, &vo.H, &bt, &at);
        same_pad(vin.W, o.kw, o.stride, 1, &vo.W, &bl, &al);
        o.pad_t += bt; o.pad_b += at; o.pad_l += bl; o.pad_r += al;
      } else {
        vo.H = (vin.H - o.kh) / o.stride + 1;
        vo.W = (vin.W - o.kw) / o.stride + 1;
      }
      if (op == "AvgPool") {
        if (!(o.kh == vin.H && o.kw == vin.W && vo.H == 1 && vo.W == 1) || o.explicit_zero)
          fail("AvgPool '" + n.name + "': only the global (full-map) average pool of the embedding nets is supported");
        o.kind = IOp::GAP;
        vo.vec = true;
      } else {
        o.kind = IOp::MAXPOOL;
      }
      o.out = new_val(vo);
      ops.push_back(o);
      return {act(o.out)};
    }

    if (op == "MatMul") {
      if (a.size() != 2 || a[0].k != Sym::ACT || a[1].k != Sym::CONST) fail("MatMul '" + n.name + "': expects (activation, constant)");
      const Val vin = vals[a[0].vid];
      if (!(vin.vec || (vin.H == 1 && vin.W == 1))) fail("MatMul '" + n.name + "': input must be a pooled feature vector");
      const AttrVal* ta = n.attr("transpose_a");
      const AttrVal* tb = n.attr("transpose_b");
      if (ta && ta->b) fail("MatMul '" + n.name + "': transpose_a is not supported");
      auto w = std::make_shared<HTensor>(*a[1].c);
      if (w->shape.size() != 2) fail("MatMul '" + n.name + "': weight must be rank 2");
      if ((int64_t)w->f.size() != w->numel() || w->numel() <= 0) fail("MatMul '" + n.name + "': weight data does not match its shape");
      if (tb && tb->b) {
        HTensor t = *w;
        int64_t R = t.shape[0], Cc = t.shape[1];
        w->shape = {Cc, R};
        for (int64_t i = 0; i < R; ++i)
          for (int64_t j = 0; j < Cc; ++j) w->f[(size_t)(j * R + i)] = t.f[(size_t)(i * Cc + j)];
      }
      if (w->shape[0] != vin.C) fail("MatMul '" + n.name + "': inner dimensions do not match");
      IOp o;
      o.kind = IOp::MATMUL;
      o.name = n.name;
      o.in0 = a[0].vid;
      o.c = w;
      Val vo;
      vo.vec = true;
      vo.H = vo.W = 1;
      vo.C = (int)w->shape[1];
      o.out = new_val(vo);
      ops.push_back(o);
      return {act(o.out)};
    }
    fail("unsupported op '" + op + "' (node '" + n.name + "')");
  }
};

}  // namespace

Plan compile_graph(const Graph& g, const CompileOptions& opt) {
  if (opt.input_name.empty() || opt.output_names.empty()) fail("input and output tensor names are required");
  Lowerer lw(g, opt);
  {
    std::string in_name;
    int port;
    split_ref(opt.input_name, &in_name, &port);
    const GNode* in = g.find(in_name);
    if (!in) fail("input tensor '" + opt.input_name + "' not found in graph");  // KeyError in the reference
    if (in->op != "Placeholder") fail("input tensor '" + opt.input_name + "' is not a Placeholder");
    lw.get(opt.input_name);  // value 0
  }
  std::vector<int> out_vals;
  for (auto& o : opt.output_names) {
    Sym s = lw.get(o);
    if (s.k != Sym::ACT) fail("output '" + o + "' does not depend on the input");
    if (lw.pads.count(s.vid)) fail("output '" + o + "' is a Pad");
    out_vals.push_back(s.vid);
  }
  auto& ops = lw.ops;
  auto& vals = lw.vals;
  const int nv = (int)vals.size();
  std::vector<std::vector<int>> uses((size_t)nv);
  std::vector<char> is_out((size_t)nv, 0);
  for (int v : out_vals) is_out[(size_t)v] = 1;
  for (int i = 0; i < (int)ops.size(); ++i) {
    uses[(size_t)ops[i].in0].push_back(i);
    if (ops[i].in1 >= 0) uses[(size_t)ops[i].in1].push_back(i);
  }

  // ---- phase 2: fusion
  std::vector<char> absorbed(ops.size(), 0);
  std::vector<Layer> layers;
  for (int i = 0; i < (int)ops.size(); ++i) {
    const IOp& o = ops[i];
    if (absorbed[i]) continue;
    Layer L;
    L.name = o.name;
    L.in = o.in0;
    L.kh = o.kh; L.kw = o.kw; L.stride = o.stride; L.dil = o.dil;
    L.pad_t = o.pad_t; L.pad_l = o.pad_l; L.pad_b = o.pad_b; L.pad_r = o.pad_r;
    L.H = vals[(size_t)o.in0].H; L.W = vals[(size_t)o.in0].W;
    L.cin = vals[(size_t)o.in0].C;
    L.Ho = vals[(size_t)o.out].H; L.Wo = vals[(size_t)o.out].W;
    L.cout = vals[(size_t)o.out].C;
    L.explicit_zero_pad = o.explicit_zero;
    int cur = o.out;
    if (o.kind == IOp::CONV || o.kind == IOp::DWCONV || o.kind == IOp::MATMUL) {
      const HTensor& w = *o.c;
      if (o.kind == IOp::CONV) {
        const bool stem = (o.in0 == 0);
        if (stem) {
          L.kind = L_STEM;
          L.w = w.f;  // [kh][kw][cin][cout] as stored
        } else {
          L.kind = (o.kh == 1 && o.kw == 1) ? L_PW : L_CONV;
          const int taps = o.kh * o.kw;
          L.w.resize(w.f.size());
          for (int t = 0; t < taps; ++t)
            for (int ci = 0; ci < L.cin; ++ci)
              for (int co = 0; co < L.cout; ++co)
                L.w[((size_t)co * taps + t) * L.cin + ci] = w.f[((size_t)t * L.cin + ci) * L.cout + co];
        }
      } else if (o.kind == IOp::DWCONV) {
        if (o.kh != 3 || o.kw != 3 || (o.stride != 1 && o.stride != 2) || o.dil != 1)
          fail("depthwise '" + o.name + "': only 3x3, stride 1|2 is supported");
        L.kind = L_DW;
        L.w = w.f;  // [3][3][C][1] == [9][C]
      } else {
        L.kind = L_FC;
        L.w = w.f;  // [K][N]
      }
      bool has_act = false, has_res = false;
      auto scale_out_channel = [&](int co, float s) {
        if (L.kind == L_STEM) {
          const int rows = o.kh * o.kw * L.cin;
          for (int r = 0; r < rows; ++r) L.w[(size_t)r * L.cout + co] *= s;
        } else if (L.kind == L_DW) {
          for (int t = 0; t < 9; ++t) L.w[(size_t)t * L.cout + co] *= s;
        } else if (L.kind == L_FC) {
          for (int k = 0; k < L.cin; ++k) L.w[(size_t)k * L.cout + co] *= s;
        } else {
          const size_t K = (size_t)o.kh * o.kw * L.cin;
          for (size_t k = 0; k < K; ++k) L.w[(size_t)co * K + k] *= s;
        }
      };
      while (uses[(size_t)cur].size() == 1 && !is_out[(size_t)cur]) {
        const int ni = uses[(size_t)cur][0];
        if (absorbed[(size_t)ni]) break;  // e.g. the residual add already belongs to the other branch's convolution
        const IOp& nx = ops[(size_t)ni];
        bool took = true;
        if (nx.kind == IOp::SCALE && !has_act && !has_res) {
          for (int co = 0; co < L.cout; ++co) {
            scale_out_channel(co, nx.c->f[(size_t)co]);
            if (!L.bias.empty()) L.bias[(size_t)co] *= nx.c->f[(size_t)co];
          }
        } else if (nx.kind == IOp::SHIFT && !has_act && !has_res) {
          if (L.bias.empty()) L.bias.assign((size_t)L.cout, 0.f);
          for (int co = 0; co < L.cout; ++co) L.bias[(size_t)co] += nx.c->f[(size_t)co];
        } else if (nx.kind == IOp::ADD2 && !has_act && !has_res && (L.kind == L_PW || L.kind == L_CONV)) {
          L.in2 = nx.in0 == cur ? nx.in1 : nx.in0;
          has_res = true;
        } else if (nx.kind == IOp::RELU && !has_act) {
          L.act = A_RELU;
          has_act = true;
        } else if (nx.kind == IOp::RELU6 && !has_act) {
          L.act = A_RELU6;
          has_act = true;
        } else if (nx.kind == IOp::MINC && L.act == A_RELU && nx.scalar == 6.f) {
          L.act = A_RELU6;  // Keras relu(max_value=6): Relu -> Minimum(6) -> Maximum(0)
        } else if (nx.kind == IOp::MAXC && (L.act == A_RELU || L.act == A_RELU6) && nx.scalar == 0.f) {
          // already non-negative
        } else if (nx.kind == IOp::SOFTMAX && !has_act && L.kind == L_FC) {
          L.act = A_SOFTMAX;
          has_act = true;
        } else if (nx.kind == IOp::SIGMOID && !has_act && L.kind == L_FC) {
          L.act = A_SIGMOID;
          has_act = true;
        } else {
          took = false;
        }
        if (!took) break;
        absorbed[(size_t)ni] = 1;
        cur = nx.out;
        L.name = nx.name;
      }
    } else if (o.kind == IOp::MAXPOOL) {
      L.kind = L_MAXPOOL;
    } else if (o.kind == IOp::GAP) {
      L.kind = L_GAP;
    } else {
      fail("op '" + o.name + "' could not be fused into a producing convolution/dense layer (unsupported op order)");
    }
    L.out = cur;
    layers.push_back(std::move(L));
  }

  // ---- phase 3: topological order from the outputs, drop dead layers, renumber values, lifetimes
  std::unordered_map<int, int> producer;  // value -> layer index (pre-sort)
  for (int i = 0; i < (int)layers.size(); ++i) producer[layers[(size_t)i].out] = i;
  std::vector<int> order;
  std::vector<char> state(layers.size(), 0);
  std::function<void(int)> visit = [&](int v) {
    if (v == 0) return;
    auto it = producer.find(v);
    if (it == producer.end()) fail("internal: value without producer");
    int li = it->second;
    if (state[(size_t)li] == 2) return;
    if (state[(size_t)li] == 1) fail("cycle in layer graph");
    state[(size_t)li] = 1;
    visit(layers[(size_t)li].in);
    if (layers[(size_t)li].in2 >= 0) visit(layers[(size_t)li].in2);
    state[(size_t)li] = 2;
    order.push_back(li);
  };
  for (int v : out_vals) visit(v);

  Plan plan;
  plan.in_h = vals[0].H;
  plan.in_w = vals[0].W;
  plan.in_c = vals[0].C;
  std::unordered_map<int, int> remap;
  remap[0] = 0;
  ValueInfo vi0;
  vi0.H = vals[0].H; vi0.W = vals[0].W; vi0.C = vals[0].C;
  plan.values.push_back(vi0);
  std::map<std::pair<int, int>, int> subsampled;  // (plan value, stride) -> subsampled plan value
  for (int li : order) {
    Layer L = layers[(size_t)li];
    if (L.kind == L_PW && L.stride > 1) {
      // strided 1x1 convolution = spatial gather + GEMM; blocks that share the input (reduce / proj) share the gather
      const int src = remap.at(L.in);
      auto key = std::make_pair(src, L.stride);
      auto it = subsampled.find(key);
      if (it == subsampled.end()) {
        Layer S;
        S.kind = L_SUBSAMPLE;
        S.name = L.name + "/subsample";
        S.in = src;
        S.stride = L.stride;
        S.H = L.H; S.W = L.W; S.Ho = L.Ho; S.Wo = L.Wo;
        S.cin = S.cout = L.cin;
        ValueInfo sv;
        sv.H = L.Ho; sv.W = L.Wo; sv.C = L.cin;
        sv.producer = (int)plan.layers.size();
        plan.values.push_back(sv);
        S.out = (int)plan.values.size() - 1;
        plan.layers.push_back(S);
        it = subsampled.emplace(key, S.out).first;
      }
      const int tmp_id = -2 - it->second;  // negative marker: already a plan value id
      L.stride = 1;
      L.H = L.Ho;
      L.W = L.Wo;
      L.pad_t = L.pad_b = L.pad_l = L.pad_r = 0;
      L.in = tmp_id;
    }
    const Val& vo = vals[(size_t)L.out];
    ValueInfo vi;
    vi.H = vo.H; vi.W = vo.W; vi.C = vo.C;
    vi.is_vector = vo.vec || L.kind == L_GAP || L.kind == L_FC;
    vi.producer = (int)plan.layers.size();
    plan.values.push_back(vi);
    remap[L.out] = (int)plan.values.size() - 1;
    L.in = L.in <= -2 ? -2 - L.in : remap.at(L.in);
    if (L.in2 >= 0) L.in2 = remap.at(L.in2);
    L.out = remap.at(L.out);
    plan.layers.push_back(std::move(L));
  }
  for (int i = 0; i < (int)plan.layers.size(); ++i) {
    const Layer& L = plan.layers[(size_t)i];
    plan.values[(size_t)L.in].last_use = i;
    if (L.in2 >= 0) plan.values[(size_t)L.in2].last_use = i;
  }
  for (size_t i = 0; i < out_vals.size(); ++i) {
    int v = remap.at(out_vals[i]);
    plan.values[(size_t)v].last_use = INT_MAX;
    plan.outputs.push_back(v);
    plan.output_names.push_back(opt.output_names[i]);
  }
  // sanity: the kernels' structural requirements
  for (const Layer& L : plan.layers) {
    if (L.kind == L_STEM && (L.cout % 32 != 0)) fail("stem '" + L.name + "': output channels must be a multiple of 32");
    if (L.kind == L_PW && (L.pad_t || L.pad_l || L.pad_b || L.pad_r)) fail("1x1 convolution '" + L.name + "' with padding");
    if ((L.kind == L_DW || L.kind == L_PW || L.kind == L_CONV) && (L.cin % 32 != 0 || L.cout % 32 != 0))
      fail("layer '" + L.name + "': channel counts must be multiples of 32");
    if (L.kind == L_FC && !plan.values[(size_t)L.in].is_vector) fail("dense '" + L.name + "' needs a pooled input");
  }
  return plan;
}

std::string Plan::to_json() const {
  std::ostringstream os;
  os.precision(9);
  os << "{\"in_h\":" << in_h << ",\"in_w\":" << in_w << ",\"in_c\":" << in_c << ",\"outputs\":[";
  for (size_t i = 0; i < outputs.size(); ++i) {
    const ValueInfo& v = values[(size_t)outputs[i]];
    os << (i ? "," : "") << "{\"name\":\"" << output_names[i] << "\",\"value\":" << outputs[i]
       << ",\"dim\":" << (long long)v.H * v.W * v.C << "}";
  }
  os << "],\"layers\":[";
  static const char* kinds[] = {"stem", "dw", "pw", "conv", "maxpool", "gap", "fc", "subsample"};
  static const char* acts[] = {"none", "relu", "relu6", "sigmoid", "softmax"};
  for (size_t i = 0; i < layers.size(); ++i) {
    const Layer& L = layers[i];
    double ws = 0, bs = 0, wa = 0;
    for (float f : L.w) { ws += f; wa += std::fabs(f); }
    for (float f : L.bias) bs += f;
    os << (i ? "," : "") << "{\"kind\":\"" << kinds[L.kind] << "\",\"name\":\"" << L.name << "\",\"in\":" << L.in
       << ",\"in2\":" << L.in2 << ",\"out\":" << L.out << ",\"k\":[" << L.kh << "," << L.kw << "],\"stride\":" << L.stride
       << ",\"pad\":[" << L.pad_t << "," << L.pad_b << "," << L.pad_l << "," << L.pad_r << "],\"cin\":" << L.cin
       << ",\"cout\":" << L.cout << ",\"hw_in\":[" << L.H << "," << L.W << "],\"hw_out\":[" << L.Ho << "," << L.Wo
       << "],\"act\":\"" << acts[L.act] << "\",\"has_bias\":" << (L.bias.empty() ? "false" : "true")
       << ",\"explicit_zero_pad\":" << (L.explicit_zero_pad ? "true" : "false") << ",\"w_sum\":" << ws
       << ",\"w_abs_sum\":" << wa << ",\"b_sum\":" << bs << "}";
  }
  os << "]}";
  return os.str();
}

}  // namespace hfr
