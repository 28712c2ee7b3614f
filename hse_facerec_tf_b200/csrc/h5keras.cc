// Keras HDF5 weight files without libhdf5: a minimal reader for the HDF5 structures h5py/Keras 2.x write
// (superblock v0/v1, version-1 object headers, symbol-table groups = B-tree v1 + local heap + SNOD nodes, contiguous or
// compact little-endian float32 datasets), and the front end that turns `models/vgg2_mobilenet.h5`
// (facerec_test.py:326-334: MobileNet(include_top=False) + GlobalAveragePooling2D + Reshape((1,1,1024),'reshape_1');
// also the age/gender .hdf5 of age_gender_train.py:89-100 when the head layers are present) into the same Graph the
// .pb path compiles, so folding / fusion / planning are shared.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>

#include "graph.h"

namespace hfr {
namespace {

[[noreturn]] void h5fail(const std::string& m) { throw std::runtime_error("hdf5: " + m); }

struct H5Dataset {
  std::vector<int64_t> dims;
  std::vector<float> data;
};

class H5File {
 public:
  H5File(const uint8_t* d, size_t n) : d_(d), n_(n) {}

  // every float32 dataset of the file, keyed by its absolute path ("/model_weights/conv1/conv1/kernel:0")
  std::map<std::string, H5Dataset> datasets() {
    static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (n_ < 96 || memcmp(d_, sig, 8) != 0) h5fail("bad signature");
    const int ver = d_[8];
    if (ver > 1) h5fail("superblock version " + std::to_string(ver) + " (libver='latest' files) is not supported");
    if (d_[13] != 8 || d_[14] != 8) h5fail("only 8-byte offsets/lengths are supported");
    size_t pos = 24 + (ver == 1 ? 4 : 0);
    base_ = u64(pos);
    pos += 32;  // base, free-space, eof, driver info
    // root symbol table entry: link name offset, object header address, cache type, reserved, scratch
    const uint64_t root_hdr = u64(pos + 8);
    std::map<std::string, H5Dataset> out;
    walk(root_hdr, "", out, 0);
    return out;
  }

 private:
  const uint8_t* d_;
  size_t n_;
  uint64_t base_ = 0;

  void need(uint64_t off, uint64_t len) const {
    if (off > n_ || len > n_ - off) h5fail("truncated file / bad address");
  }
  uint64_t u64(uint64_t off) const { need(off, 8); uint64_t v; memcpy(&v, d_ + off, 8); return v; }
  uint32_t u32(uint64_t off) const { need(off, 4); uint32_t v; memcpy(&v, d_ + off, 4); return v; }
  uint16_t u16(uint64_t off) const { need(off, 2); uint16_t v; memcpy(&v, d_ + off, 2); return v; }
  uint8_t u8(uint64_t off) const { need(off, 1); return d_[off]; }

  struct Msg { int type; uint64_t off; uint32_t size; };

  // version-1 object header (with continuation blocks)
  std::vector<Msg> messages(uint64_t addr) const {
    addr += base_;
    if (u8(addr) != 1) {
      need(addr, 4);
      if (memcmp(d_ + addr, "OHDR", 4) == 0) h5fail("version-2 object headers (libver='latest') are not supported");
      h5fail("unsupported object header version");
    }
    int nmsg = u16(addr + 2);
    uint64_t chunk_size = u32(addr + 8);
    std::vector<Msg> out;
    std::vector<std::pair<uint64_t, uint64_t>> chunks{{addr + 16, chunk_size}};
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
      uint64_t p = chunks[ci].first, end = p + chunks[ci].second;
      need(p, chunks[ci].second);
      while (p + 8 <= end && (int)out.size() < nmsg) {
        Msg m{(int)u16(p), p + 8, u16(p + 2)};
        need(m.off, m.size);
        if (m.type == 0x0010) chunks.push_back({u64(m.off) + base_, u64(m.off + 8)});
        out.push_back(m);
        p += 8 + m.size;
      }
    }
    return out;
  }

  std::string heap_string(uint64_t heap_addr, uint64_t off) const {
    heap_addr += base_;
    need(heap_addr, 32);
    if (memcmp(d_ + heap_addr, "HEAP", 4) != 0) h5fail("bad local heap signature");
    const uint64_t seg_size = u64(heap_addr + 8), seg = u64(heap_addr + 24) + base_;
    if (off >= seg_size) h5fail("bad link name offset");
    need(seg, seg_size);
    const char* s = (const char*)d_ + seg + off;
    return std::string(s, strnlen(s, (size_t)(seg_size - off)));
  }

  // group B-tree (type 0) -> symbol table nodes -> (name, object header address)
  void btree(uint64_t addr, uint64_t heap, std::vector<std::pair<std::string, uint64_t>>& links, int depth) const {
    if (depth > 32) h5fail("B-tree too deep");
    addr += base_;
    need(addr, 24);
    if (memcmp(d_ + addr, "TREE", 4) == 0) {
      if (u8(addr + 4) != 0) h5fail("unexpected B-tree node type");
      const int level = u8(addr + 5), used = u16(addr + 6);
      uint64_t p = addr + 24;  // key0
      for (int i = 0; i < used; ++i) {
        const uint64_t child = u64(p + 8);
        if (level > 0) btree(child, heap, links, depth + 1);
        else snod(child, heap, links);
        p += 16;
      }
    } else {
      h5fail("bad B-tree signature");
    }
  }
  void snod(uint64_t addr, uint64_t heap, std::vector<std::pair<std::string, uint64_t>>& links) const {
    addr += base_;
    need(addr, 8);
    if (memcmp(d_ + addr, "SNOD", 4) != 0) h5fail("bad symbol table node signature");
    const int n = u16(addr + 6);
    for (int i = 0; i < n; ++i) {
      const uint64_t e = addr + 8 + (uint64_t)i * 40;
      links.push_back({heap_string(heap, u64(e)), u64(e + 8)});
    }
  }

  void walk(uint64_t hdr, const std::string& path, std::map<std::string, H5Dataset>& out, int depth) {
    if (depth > 16) h5fail("group nesting too deep");
    std::vector<Msg> msgs = messages(hdr);
    const Msg *stab = nullptr, *space = nullptr, *dtype = nullptr, *layout = nullptr, *filters = nullptr;
    for (const Msg& m : msgs) {
      if (m.type == 0x0011) stab = &m;
      if (m.type == 0x0001) space = &m;
      if (m.type == 0x0003) dtype = &m;
      if (m.type == 0x0008) layout = &m;
      if (m.type == 0x000B) filters = &m;
    }
    if (stab) {
      std::vector<std::pair<std::string, uint64_t>> links;
      btree(u64(stab->off), u64(stab->off + 8), links, 0);
      for (auto& l : links) walk(l.second, path + "/" + l.first, out, depth + 1);
      return;
    }
    if (!(space && dtype && layout)) return;  // not a dataset we understand (e.g. a committed datatype)
    // datatype: class 1 (floating point), 4 bytes, little endian
    const int cls = u8(dtype->off) & 0x0F;
    const uint32_t tsize = u32(dtype->off + 4);
    if (cls != 1 || tsize != 4 || (u8(dtype->off + 1) & 1)) return;  // only LE float32 datasets carry weights
    if (filters) h5fail("dataset '" + path + "' is filtered/compressed");
    H5Dataset ds;
    const int sver = u8(space->off), rank = u8(space->off + 1);
    uint64_t dp = space->off + (sver == 1 ? 8 : 4);
    int64_t numel = 1;
    for (int i = 0; i < rank; ++i) {
      ds.dims.push_back((int64_t)u64(dp + 8 * i));
      numel *= ds.dims.back();
    }
    const int lver = u8(layout->off);
    uint64_t data_addr = 0;
    const uint8_t* src = nullptr;
    if (lver == 3) {
      const int lclass = u8(layout->off + 1);
      if (lclass == 1) {
        data_addr = u64(layout->off + 2);
        const uint64_t sz = u64(layout->off + 10);
        if (sz < (uint64_t)numel * 4) h5fail("dataset '" + path + "': storage smaller than its shape");
      } else if (lclass == 0) {
        const uint32_t sz = u16(layout->off + 2);
        if (sz < (uint64_t)numel * 4) h5fail("dataset '" + path + "': compact storage smaller than its shape");
        src = d_ + layout->off + 4;
      } else {
        h5fail("dataset '" + path + "' is chunked (not supported; Keras writes weights contiguously)");
      }
    } else if (lver == 1 || lver == 2) {
      const int ldim = u8(layout->off + 1), lclass = u8(layout->off + 2);
      if (lclass != 1) h5fail("dataset '" + path + "': only contiguous storage is supported");
      (void)ldim;
      data_addr = u64(layout->off + 8);
    } else {
      h5fail("dataset '" + path + "': unsupported data layout version");
    }
    if (!src) {
      if (data_addr == ~0ull) {
        if (numel) h5fail("dataset '" + path + "' has no allocated storage");
      } else {
        need(data_addr + base_, (uint64_t)numel * 4);
        src = d_ + data_addr + base_;
      }
    }
    ds.data.resize((size_t)numel);
    if (numel) memcpy(ds.data.data(), src, (size_t)numel * 4);
    out[path] = std::move(ds);
  }
};

// ---------------------------------------------------------------------------------------------- Keras -> Graph
struct GraphBuilder {
  Graph g;
  GNode& add(const std::string& name, const std::string& op, std::vector<std::string> inputs = {}) {
    GNode n;
    n.name = name;
    n.op = op;
    n.inputs = std::move(inputs);
    g.nodes.push_back(std::move(n));
    return g.nodes.back();
  }
  void constant(const std::string& name, const std::vector<int64_t>& shape, const std::vector<float>& vals, int dtype = 1) {
    GNode& n = add(name, "Const");
    AttrVal v;
    v.kind = AttrVal::TENSOR;
    v.tensor.dtype = dtype;
    v.tensor.shape = shape;
    v.tensor.f = vals;
    n.attrs["value"] = std::move(v);
  }
  static AttrVal ints(std::vector<int64_t> v) {
    AttrVal a;
    a.kind = AttrVal::LIST;
    a.shape = std::move(v);
    return a;
  }
  static AttrVal str(const std::string& s) {
    AttrVal a;
    a.kind = AttrVal::S;
    a.s = s;
    return a;
  }
  void finish() {
    for (size_t i = 0; i < g.nodes.size(); ++i) g.index[g.nodes[i].name] = (int)i;
  }
};

}  // namespace

Plan compile_keras_mobilenet_h5(const uint8_t* data, size_t size, int input_hw, const CompileOptions& opt_in) {
  std::map<std::string, H5Dataset> ds = H5File(data, size).datasets();
  // Keras stores weight w of layer L at .../L/L/w:0 (full-model files add a leading /model_weights)
  auto find = [&](const std::string& layer, const std::string& w) -> const H5Dataset* {
    const std::string tail = "/" + layer + "/" + layer + "/" + w + ":0";
    const H5Dataset* hit = nullptr;
    for (auto& kv : ds) {
      const std::string& p = kv.first;
      if (p.size() >= tail.size() && p.compare(p.size() - tail.size(), tail.size(), tail) == 0 &&
          p.find("optimizer_weights") == std::string::npos)
        hit = &kv.second;
    }
    return hit;
  };
  auto must = [&](const std::string& layer, const std::string& w) -> const H5Dataset& {
    const H5Dataset* d = find(layer, w);
    if (!d) h5fail("weight '" + layer + "/" + w + ":0' not found (not a Keras MobileNet-v1 weight file?)");
    return *d;
  };

  GraphBuilder b;
  const int hw = input_hw > 0 ? input_hw : 192;  // facerec_test.py:325 sz=192
  {
    GNode& in = b.add("input_1", "Placeholder");
    AttrVal sh;
    sh.kind = AttrVal::SHAPE;
    sh.shape = {-1, hw, hw, 3};
    in.attrs["shape"] = sh;
  }
  auto bn_relu6 = [&](const std::string& layer, const std::string& x) -> std::string {
    const std::string bn = layer + "_bn";
    for (const char* w : {"gamma", "beta", "moving_mean", "moving_variance"}) {
      const H5Dataset& d = must(bn, w);
      b.constant(bn + "/" + w, d.dims, d.data);
    }
    GNode& n = b.add(bn + "/FusedBatchNorm", "FusedBatchNorm",
                     {x, bn + "/gamma", bn + "/beta", bn + "/moving_mean", bn + "/moving_variance"});
    AttrVal eps;
    eps.kind = AttrVal::F;
    eps.f = 1e-3f;  // Keras BatchNormalization default epsilon
    n.attrs["epsilon"] = eps;
    AttrVal tr;
    tr.kind = AttrVal::B;
    tr.b = false;
    n.attrs["is_training"] = tr;
    b.add(layer + "_relu/Relu6", "Relu6", {bn + "/FusedBatchNorm"});
    return layer + "_relu/Relu6";
  };
  auto conv = [&](const std::string& layer, const std::string& op, const std::string& wname, const std::string& x,
                  int stride) -> std::string {
    const H5Dataset& k = must(layer, wname);
    if (k.dims.size() != 4) h5fail("kernel of '" + layer + "' is not rank 4");
    b.constant(layer + "/" + wname, k.dims, k.data);
    GNode& n = b.add(layer + "/" + (op == "Conv2D" ? "convolution" : "depthwise"), op, {x, layer + "/" + wname});
    n.attrs["strides"] = GraphBuilder::ints({1, stride, stride, 1});
    n.attrs["padding"] = GraphBuilder::str("SAME");
    n.attrs["data_format"] = GraphBuilder::str("NHWC");
    return n.name;
  };
  static const int kStrides[13] = {1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1};
  std::string x = bn_relu6("conv1", conv("conv1", "Conv2D", "kernel", "input_1", 2));
  int n_blocks = 0;
  for (int i = 1; i <= 13; ++i) {
    const std::string dw = "conv_dw_" + std::to_string(i), pw = "conv_pw_" + std::to_string(i);
    if (!find(dw, "depthwise_kernel")) break;
    x = bn_relu6(dw, conv(dw, "DepthwiseConv2dNative", "depthwise_kernel", x, kStrides[i - 1]));
    x = bn_relu6(pw, conv(pw, "Conv2D", "kernel", x, 1));
    ++n_blocks;
  }
  if (n_blocks != 13) h5fail("expected 13 depthwise/pointwise blocks, found " + std::to_string(n_blocks));
  const int feat = (int)must("conv_pw_13", "kernel").dims[3];
  b.constant("global_pooling/Mean/reduction_indices", {2}, {1.f, 2.f}, 3);
  b.add("global_pooling/Mean", "Mean", {x, "global_pooling/Mean/reduction_indices"});
  b.constant("reshape_1/Reshape/shape", {4}, {-1.f, 1.f, 1.f, (float)feat}, 3);
  b.add("reshape_1/Reshape", "Reshape", {"global_pooling/Mean", "reshape_1/Reshape/shape"});
  // optional age/gender heads (age_gender_train.py:91-98)
  auto dense = [&](const std::string& layer, const std::string& xin, const std::string& act) {
    const H5Dataset& k = must(layer, "kernel");
    b.constant(layer + "/kernel", k.dims, k.data);
    b.add(layer + "/MatMul", "MatMul", {xin, layer + "/kernel"});
    std::string y = layer + "/MatMul";
    if (const H5Dataset* bias = find(layer, "bias")) {
      b.constant(layer + "/bias", bias->dims, bias->data);
      b.add(layer + "/BiasAdd", "BiasAdd", {y, layer + "/bias"});
      y = layer + "/BiasAdd";
    }
    b.add(layer + "/" + act, act, {y});
  };
  if (find("feats", "kernel")) {
    dense("feats", "global_pooling/Mean", "Relu");
    if (find("age_pred", "kernel")) dense("age_pred", "feats/Relu", "Softmax");
    if (find("gender_pred", "kernel")) dense("gender_pred", "feats/Relu", "Sigmoid");
  }
  b.finish();
  CompileOptions opt = opt_in;
  opt.input_name = "input_1:0";
  opt.override_hw = 0;
  if (opt.output_names.empty()) opt.output_names = {"reshape_1/Reshape:0"};
  return compile_graph(b.g, opt);
}

}  // namespace hfr
