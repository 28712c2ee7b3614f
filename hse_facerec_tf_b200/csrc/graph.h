// Host-side model front end: TensorFlow GraphDef (frozen .pb) reader and the graph compiler that lowers it to the
// fused layer plan the CUDA executor runs.  Replaces `load_graph` (facerec_test.py:41-48) /
// `FacialImageProcessing.load_graph_def` (facial_analysis.py:319-325) + tf.import_graph_def + the TF runtime's
// placement/constant folding.  Pure C++17, no CUDA, no protobuf library.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace hfr {

struct HTensor {
  int dtype = 1;                 // TF DataType enum (1 float, 3 int32, 9 int64, 10 bool, 12 quint8 ...)
  std::vector<int64_t> shape;
  std::vector<float> f;          // values converted to float (integers are exact up to 2^24; shapes/axes are tiny)
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

struct AttrVal {
  enum Kind { NONE, S, I, F, B, TYPE, SHAPE, TENSOR, LIST } kind = NONE;
  std::string s;
  int64_t i = 0;
  float f = 0.f;
  bool b = false;
  std::vector<int64_t> shape;    // SHAPE (-1 = unknown) ; LIST: list(i)
  bool unknown_rank = false;
  HTensor tensor;
};

struct GNode {
  std::string name, op;
  std::vector<std::string> inputs;
  std::map<std::string, AttrVal> attrs;
  const AttrVal* attr(const std::string& k) const {
    auto it = attrs.find(k);
    return it == attrs.end() ? nullptr : &it->second;
  }
};

struct Graph {
  std::vector<GNode> nodes;
  std::unordered_map<std::string, int> index;
  const GNode* find(const std::string& name) const {
    auto it = index.find(name);
    return it == index.end() ? nullptr : &nodes[it->second];
  }
};

// Throws std::runtime_error on malformed input.
void parse_graphdef(const uint8_t* data, size_t size, Graph* g);

// ------------------------------------------------------------------------------------------------------------------
// Fused layer plan
enum LayerKind {
  L_STEM = 0,     // direct KxK convolution over the 3-channel input (+ fused input pre-processing)
  L_DW = 1,       // depthwise 3x3
  L_PW = 2,       // 1x1 convolution = GEMM (stride > 1: spatial subsample first)
  L_CONV = 3,     // KxK convolution as implicit GEMM
  L_MAXPOOL = 4,
  L_GAP = 5,      // global average pool -> fp32 [B, C]
  L_FC = 6,       // dense head on fp32 vectors
  L_SUBSAMPLE = 7 // spatial stride-s gather in front of a strided 1x1 convolution
};
enum { A_NONE = 0, A_RELU = 1, A_RELU6 = 2, A_SIGMOID = 3, A_SOFTMAX = 4 };

struct Layer {
  int kind = 0;
  std::string name;       // name of the graph node that produces the layer's final value
  int in = -1, in2 = -1;  // value ids: main input, residual input (-1 = none)
  int out = -1;
  int kh = 1, kw = 1, stride = 1, dil = 1;
  int pad_t = 0, pad_l = 0, pad_b = 0, pad_r = 0;
  int cin = 0, cout = 0;
  int H = 0, W = 0, Ho = 0, Wo = 0;
  int act = A_NONE;
  bool explicit_zero_pad = false;  // max pool fed by an explicit Pad op
  std::vector<float> w;            // STEM [kh][kw][cin][cout]; DW [9][c]; PW/CONV [cout][kh*kw][cin]; FC [k][n]
  std::vector<float> bias;         // [cout] (empty = none)
};

struct ValueInfo {
  int H = 0, W = 0, C = 0;
  bool is_vector = false;  // fp32 [B, C] (after pooling) instead of an NHWC activation
  int producer = -1;       // layer index, -1 = graph input
  int last_use = -1;       // last layer index reading it (outputs: INT_MAX)
};

struct Plan {
  int in_h = 0, in_w = 0, in_c = 3;
  std::vector<Layer> layers;          // in execution order
  std::vector<ValueInfo> values;      // value 0 = network input
  std::vector<int> outputs;           // value ids, in the order requested
  std::vector<std::string> output_names;
  std::string to_json() const;
};

struct CompileOptions {
  std::string input_name;                 // e.g. "input_1:0"
  std::vector<std::string> output_names;  // e.g. {"global_pooling/Mean:0"}
  std::string phase_name;                 // learning-phase placeholder fed with a constant ("" = none)
  float phase_value = 0.f;
  int override_hw = 0;                    // run the (fully convolutional) graph at another input size; 0 = placeholder's
};

// Lowers the graph: constant folding (Dequantize MIN_FIRST, BN arithmetic), dead-branch elimination through
// Switch/Merge, then fusion of conv -> scale -> shift -> (residual add) -> activation chains.
Plan compile_graph(const Graph& g, const CompileOptions& opt);

// Keras HDF5 weights (models/vgg2_mobilenet.h5, facerec_test.py:326-334) -> the same plan.  opt.output_names may name
// reshape_1/Reshape:0 (default), global_pooling/Mean:0 and, when the head layers exist, feats/Relu:0,
// age_pred/Softmax:0, gender_pred/Sigmoid:0.
Plan compile_keras_mobilenet_h5(const uint8_t* data, size_t size, int input_hw, const CompileOptions& opt);

}  // namespace hfr
