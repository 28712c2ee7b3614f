"""Synthetic model files in the reference's on-disk formats, for the models whose weights are not redistributable /
not shipped (models/vgg2_mobilenet.pb, models/vgg2_resnet.pb - .MISSING_LARGE_BLOBS in the reference mount).

The files are real frozen TF-1.x GraphDefs (protobuf wire format written by hand, no TensorFlow needed) with the
node names the reference binds to (facerec_test.py:212-213):
  vgg2_mobilenet.pb  input_1:0 -> reshape_1/Reshape:0, un-folded Keras BatchNorm under the
                     conv1_bn/keras_learning_phase Switch/Merge conditionals
  vgg2_resnet.pb     input:0 -> pool5_7x7_s1:0, Caffe-style VGGFace2 resnet50_ft (stride on the 1x1 reduce,
                     7x7/2 conv pad 3, 3x3/2 ceil-mode max pool -> 56x56, FusedBatchNorm eps 1e-5)
Weights are seeded random, scaled per layer so that activations stay O(1) through the depth of the network.
"""
from __future__ import annotations

import os
import struct

import numpy as np

DT_FLOAT, DT_INT32, DT_BOOL = 1, 3, 10


# ------------------------------------------------------------------------------------------- protobuf wire writer
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fno, wt):
    return _varint((fno << 3) | wt)


def _ld(fno, payload: bytes) -> bytes:
    return _key(fno, 2) + _varint(len(payload)) + payload


def _vi(fno, v: int) -> bytes:
    return _key(fno, 0) + _varint(v)


def _shape(dims) -> bytes:
    return b"".join(_ld(2, _vi(1, int(d))) for d in dims)


def _tensor(arr: np.ndarray) -> bytes:
    if arr.dtype == np.float32:
        dt = DT_FLOAT
    elif arr.dtype == np.int32:
        dt = DT_INT32
    elif arr.dtype == np.bool_:
        dt = DT_BOOL
    else:
        raise TypeError(arr.dtype)
    body = _vi(1, dt) + _ld(2, _shape(arr.shape))
    if arr.ndim == 0 and dt == DT_FLOAT:
        body += _key(5, 5) + struct.pack("<f", float(arr))       # scalar as a single float_val, like TF writes them
    else:
        body += _ld(4, np.ascontiguousarray(arr).tobytes())
    return body


def attr_s(s: bytes):
    return _ld(2, s)


def attr_i(v: int):
    return _vi(3, v)


def attr_f(v: float):
    return _key(4, 5) + struct.pack("<f", v)


def attr_b(v: bool):
    return _vi(5, int(v))


def attr_type(t: int):
    return _vi(6, t)


def attr_shape(dims):
    return _ld(7, _shape(dims))


def attr_tensor(arr):
    return _ld(8, _tensor(arr))


def attr_ints(vals):
    return _ld(1, _ld(3, b"".join(_varint(int(v)) for v in vals)))


class GraphWriter:
    def __init__(self):
        self.nodes = []

    def node(self, name, op, inputs=(), **attrs):
        body = _ld(1, name.encode()) + _ld(2, op.encode())
        for i in inputs:
            body += _ld(3, i.encode())
        for k, v in attrs.items():
            body += _ld(5, _ld(1, k.encode()) + _ld(2, v))
        self.nodes.append(_ld(1, body))
        return name

    def const(self, name, arr):
        arr = np.asarray(arr)
        dt = {np.dtype("float32"): DT_FLOAT, np.dtype("int32"): DT_INT32, np.dtype("bool"): DT_BOOL}[arr.dtype]
        return self.node(name, "Const", dtype=attr_type(dt), value=attr_tensor(arr))

    def save(self, path):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            for n in self.nodes:
                f.write(n)
            f.write(_ld(4, _vi(1, 27)))  # versions { producer: 27 }
        os.replace(tmp, path)


def _conv(g, name, x, w, stride, padding):
    g.const(name + "/kernel", w.astype(np.float32))
    return g.node(name, "Conv2D", [x, name + "/kernel"], T=attr_type(DT_FLOAT), strides=attr_ints([1, stride, stride, 1]),
                  padding=attr_s(padding), data_format=attr_s(b"NHWC"), dilations=attr_ints([1, 1, 1, 1]),
                  use_cudnn_on_gpu=attr_b(True))


def _bn_params(rs, c, gamma_scale=1.0):
    gamma = (rs.uniform(0.6, 1.4, c) * gamma_scale).astype(np.float32)
    beta = rs.normal(0, 0.15, c).astype(np.float32)
    mean = rs.normal(0, 0.2, c).astype(np.float32)
    var = rs.uniform(0.5, 1.5, c).astype(np.float32)
    return gamma, beta, mean, var


# ------------------------------------------------------------------------------------------- ResNet-50 (Caffe style)
def write_resnet50_pb(path, seed=1234, input_hw=224):
    rs = np.random.RandomState(seed)
    g = GraphWriter()
    g.node("input", "Placeholder", dtype=attr_type(DT_FLOAT), shape=attr_shape([-1, input_hw, input_hw, 3]))

    def fused_bn(name, x, c, gamma_scale=1.0):
        gamma, beta, mean, var = _bn_params(rs, c, gamma_scale)
        for suffix, arr in (("gamma", gamma), ("beta", beta), ("moving_mean", mean), ("moving_variance", var)):
            g.const(f"{name}/{suffix}", arr)
        return g.node(name, "FusedBatchNorm", [x, f"{name}/gamma", f"{name}/beta", f"{name}/moving_mean",
                                               f"{name}/moving_variance"], T=attr_type(DT_FLOAT), epsilon=attr_f(1e-5),
                      is_training=attr_b(False), data_format=attr_s(b"NHWC"))

    def relu(name, x):
        return g.node(name, "Relu", [x], T=attr_type(DT_FLOAT))

    def he(kh, kw, cin, cout, gain=2.0):
        return (rs.standard_normal((kh, kw, cin, cout)) * np.sqrt(gain / (kh * kw * cin))).astype(np.float32)

    g.const("conv1_7x7_s2/pad/paddings", np.array([[0, 0], [3, 3], [3, 3], [0, 0]], np.int32))
    x = g.node("conv1_7x7_s2/pad", "Pad", ["input", "conv1_7x7_s2/pad/paddings"], T=attr_type(DT_FLOAT))
    w1 = he(7, 7, 3, 64) / 60.0   # the input is mean-subtracted pixels (+-128)
    x = _conv(g, "conv1_7x7_s2", x, w1, 2, b"VALID")
    x = relu("conv1_relu_7x7_s2", fused_bn("conv1_7x7_s2_bn", x, 64))
    # Caffe ceil-mode 3x3/2 pooling 112 -> 56: one pad row/column at the bottom/right
    g.const("pool1_3x3_s2/pad/paddings", np.array([[0, 0], [0, 1], [0, 1], [0, 0]], np.int32))
    x = g.node("pool1_3x3_s2/pad", "Pad", [x, "pool1_3x3_s2/pad/paddings"], T=attr_type(DT_FLOAT))
    x = g.node("pool1_3x3_s2", "MaxPool", [x], T=attr_type(DT_FLOAT), ksize=attr_ints([1, 3, 3, 1]),
               strides=attr_ints([1, 2, 2, 1]), padding=attr_s(b"VALID"), data_format=attr_s(b"NHWC"))
    cin = 64
    for stage, (blocks, mid) in zip((2, 3, 4, 5), ((3, 64), (4, 128), (6, 256), (3, 512))):
        cout = mid * 4
        for b in range(1, blocks + 1):
            p = f"conv{stage}_{b}"
            stride = 2 if (b == 1 and stage > 2) else 1
            if b == 1:
                sc = _conv(g, p + "_1x1_proj", x, he(1, 1, cin, cout, gain=1.0), stride, b"VALID")
                sc = fused_bn(p + "_1x1_proj_bn", sc, cout)
            else:
                sc = x
            y = _conv(g, p + "_1x1_reduce", x, he(1, 1, cin, mid), stride, b"VALID")
            y = relu(p + "_1x1_reduce_relu", fused_bn(p + "_1x1_reduce_bn", y, mid))
            y = _conv(g, p + "_3x3", y, he(3, 3, mid, mid), 1, b"SAME")
            y = relu(p + "_3x3_relu", fused_bn(p + "_3x3_bn", y, mid))
            y = _conv(g, p + "_1x1_increase", y, he(1, 1, mid, cout, gain=1.0), 1, b"VALID")
            y = fused_bn(p + "_1x1_increase_bn", y, cout, gamma_scale=0.5)
            x = relu(p + "_relu", g.node(p, "Add", [sc, y], T=attr_type(DT_FLOAT)))
            cin = cout
    g.node("pool5_7x7_s1", "AvgPool", [x], T=attr_type(DT_FLOAT), ksize=attr_ints([1, input_hw // 32, input_hw // 32, 1]),
           strides=attr_ints([1, 1, 1, 1]), padding=attr_s(b"VALID"), data_format=attr_s(b"NHWC"))
    g.save(path)
    return path


# ------------------------------------------------------------------------------------------- MobileNet-v1 (Keras 2.x)
MOBILENET_CFG = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1),
                 (512, 1), (1024, 2), (1024, 1)]


def write_mobilenet_pb(path, seed=1234, input_hw=192, learning_phase=True, width=1.0):
    """Keras MobileNet(include_top=False) + GlobalAveragePooling2D + Reshape((1,1,C),'reshape_1'), frozen WITHOUT the
    graph_transforms folding: BN is the non-fused chain add->Rsqrt->mul->mul_1/mul_2->sub->add_1 (eps 1e-3), each under
    cond/Switch...cond/Merge on the `conv1_bn/keras_learning_phase` placeholder when learning_phase=True."""
    rs = np.random.RandomState(seed)
    g = GraphWriter()
    g.node("input_1", "Placeholder", dtype=attr_type(DT_FLOAT), shape=attr_shape([-1, input_hw, input_hw, 3]))
    phase = "conv1_bn/keras_learning_phase"
    if learning_phase:
        g.node(phase, "Placeholder", dtype=attr_type(DT_BOOL), shape=attr_shape([]))

    def ch(c):
        return max(32, int(c * width) // 32 * 32)

    def bn(name, x, c):
        gamma, beta, mean, var = _bn_params(rs, c)
        for suffix, arr in (("gamma", gamma), ("beta", beta), ("moving_mean", mean), ("moving_variance", var)):
            g.const(f"{name}/{suffix}", arr)
            g.node(f"{name}/{suffix}/read", "Identity", [f"{name}/{suffix}"], T=attr_type(DT_FLOAT))
        f = attr_type(DT_FLOAT)

        def chain(prefix, xin, gm, bt, mn, vr):
            g.const(f"{prefix}/add/y", np.float32(1e-3))
            a = g.node(f"{prefix}/add", "Add", [vr, f"{prefix}/add/y"], T=f)
            r = g.node(f"{prefix}/Rsqrt", "Rsqrt", [a], T=f)
            m = g.node(f"{prefix}/mul", "Mul", [r, gm], T=f)
            m1 = g.node(f"{prefix}/mul_1", "Mul", [xin, m], T=f)
            m2 = g.node(f"{prefix}/mul_2", "Mul", [mn, m], T=f)
            s = g.node(f"{prefix}/sub", "Sub", [bt, m2], T=f)
            return g.node(f"{prefix}/add_1", "Add", [m1, s], T=f)

        reads = [f"{name}/{s}/read" for s in ("gamma", "beta", "moving_mean", "moving_variance")]
        if not learning_phase:
            return chain(f"{name}/batchnorm_1", x, *reads)
        c_ = f"{name}/cond"
        g.node(f"{c_}/Switch", "Switch", [phase, phase], T=attr_type(DT_BOOL))
        g.node(f"{c_}/switch_t", "Identity", [f"{c_}/Switch:1"], T=attr_type(DT_BOOL))
        g.node(f"{c_}/switch_f", "Identity", [f"{c_}/Switch"], T=attr_type(DT_BOOL))
        g.node(f"{c_}/pred_id", "Identity", [phase], T=attr_type(DT_BOOL))
        # training branch (dead at inference): stand-in for the moments path, fed from the :1 ports
        sx_t = g.node(f"{c_}/batchnorm/Switch", "Switch", [x, f"{c_}/pred_id"], T=f)
        g.const(f"{c_}/batchnorm/dead_scale", np.float32(123.0))
        tr = g.node(f"{c_}/batchnorm/add_1", "Mul", [sx_t + ":1", f"{c_}/batchnorm/dead_scale"], T=f)
        # inference branch: every tensor enters through Switch_k:0
        sw = []
        for k, src in enumerate([x] + reads):
            sw.append(g.node(f"{c_}/batchnorm_1/Switch_{k}", "Switch", [src, f"{c_}/pred_id"], T=f))
        inf = chain(f"{c_}/batchnorm_1", sw[0], sw[1], sw[2], sw[3], sw[4])
        return g.node(f"{c_}/Merge", "Merge", [inf, tr], T=f, N=attr_i(2))

    def relu6(name, x):
        f = attr_type(DT_FLOAT)
        r = g.node(f"{name}/Relu", "Relu", [x], T=f)
        g.const(f"{name}/Const", np.float32(6.0))
        g.const(f"{name}/Const_1", np.float32(0.0))
        m = g.node(f"{name}/clip_by_value/Minimum", "Minimum", [r, f"{name}/Const"], T=f)
        return g.node(f"{name}/clip_by_value", "Maximum", [m, f"{name}/Const_1"], T=f)

    c0 = ch(32)
    w = (rs.standard_normal((3, 3, 3, c0)) * np.sqrt(2.0 / 27) / 60.0).astype(np.float32)
    x = _conv(g, "conv1/convolution", "input_1", w, 2, b"SAME")
    x = relu6("conv1_relu", bn("conv1_bn", x, c0))
    cin = c0
    for i, (cout, stride) in enumerate(MOBILENET_CFG, start=1):
        cout = ch(cout)
        dw = (rs.standard_normal((3, 3, cin, 1)) * np.sqrt(2.0 / 9)).astype(np.float32)
        g.const(f"conv_dw_{i}/depthwise_kernel", dw)
        x = g.node(f"conv_dw_{i}/depthwise", "DepthwiseConv2dNative", [x, f"conv_dw_{i}/depthwise_kernel"],
                   T=attr_type(DT_FLOAT), strides=attr_ints([1, stride, stride, 1]), padding=attr_s(b"SAME"),
                   data_format=attr_s(b"NHWC"), dilations=attr_ints([1, 1, 1, 1]))
        x = relu6(f"conv_dw_{i}_relu", bn(f"conv_dw_{i}_bn", x, cin))
        pw = (rs.standard_normal((1, 1, cin, cout)) * np.sqrt(2.0 / cin)).astype(np.float32)
        x = _conv(g, f"conv_pw_{i}/convolution", x, pw, 1, b"SAME")
        x = relu6(f"conv_pw_{i}_relu", bn(f"conv_pw_{i}_bn", x, cout))
        cin = cout
    g.const("global_average_pooling2d_1/Mean/reduction_indices", np.array([1, 2], np.int32))
    x = g.node("global_average_pooling2d_1/Mean", "Mean", [x, "global_average_pooling2d_1/Mean/reduction_indices"],
               T=attr_type(DT_FLOAT), Tidx=attr_type(DT_INT32), keep_dims=attr_b(False))
    g.const("reshape_1/Reshape/shape", np.array([-1, 1, 1, cin], np.int32))
    g.node("reshape_1/Reshape", "Reshape", [x, "reshape_1/Reshape/shape"], T=attr_type(DT_FLOAT), Tshape=attr_type(DT_INT32))
    g.save(path)
    return path


def _cache_dir():
    d = os.environ.get("HFR_SYNTH_DIR", os.path.join("/tmp", "hfr_synth"))
    os.makedirs(d, exist_ok=True)
    return d


def ensure_resnet50_pb(seed=1234) -> str:
    p = os.path.join(_cache_dir(), f"vgg2_resnet_synth_{seed}.pb")
    if not os.path.exists(p):
        write_resnet50_pb(p, seed)
    return p


def ensure_mobilenet_pb(seed=1234, input_hw=192) -> str:
    p = os.path.join(_cache_dir(), f"vgg2_mobilenet_synth_{seed}_{input_hw}.pb")
    if not os.path.exists(p):
        write_mobilenet_pb(p, seed, input_hw)
    return p
