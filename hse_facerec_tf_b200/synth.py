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


# ------------------------------------------------------------------------------------------- HDF5 (Keras weight files)
class H5Writer:
    """Minimal HDF5 writer producing what h5py/Keras 2.x write with default settings: superblock v0, version-1 object
    headers, old-style groups (B-tree v1 + local heap + symbol table node), contiguous little-endian float32 datasets,
    fixed-length string array attributes.  Groups hold at most 2*K entries per symbol table node (K = 256 here), which
    is plenty for a Keras model."""
    UNDEF = 0xFFFFFFFFFFFFFFFF
    LEAF_K = 256

    def __init__(self):
        self.buf = bytearray(b"\0" * 96)      # superblock placeholder
        self.root = {"groups": {}, "datasets": {}, "attrs": {}}

    # -- tree construction -------------------------------------------------------------------------
    def group(self, path):
        node = self.root
        for part in [p for p in path.split("/") if p]:
            node = node["groups"].setdefault(part, {"groups": {}, "datasets": {}, "attrs": {}})
        return node

    def dataset(self, path, arr):
        parts = [p for p in path.split("/") if p]
        self.group("/".join(parts[:-1]))["datasets"][parts[-1]] = np.ascontiguousarray(arr, dtype="<f4")

    def attr_strings(self, group_path, name, strings):
        self.group(group_path)["attrs"][name] = [s.encode() if isinstance(s, str) else s for s in strings]

    # -- low level -----------------------------------------------------------------------------------
    def _alloc(self, data: bytes, align=8) -> int:
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    @staticmethod
    def _msg(mtype, body: bytes) -> bytes:
        body = body + b"\0" * (-len(body) % 8)
        return struct.pack("<HHB3x", mtype, len(body), 0) + body

    def _object_header(self, msgs) -> int:
        data = b"".join(msgs)
        hdr = struct.pack("<BxHII4x", 1, len(msgs), 1, len(data))
        return self._alloc(hdr + data)

    @staticmethod
    def _dataspace(dims) -> bytes:
        return struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)

    @staticmethod
    def _dtype_f32() -> bytes:
        return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dtype_str(n) -> bytes:
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, n)      # class 3 (string), null-terminated, ASCII

    def _attr(self, name: str, strings) -> bytes:
        n = max(len(s) for s in strings) + 1 if strings else 1
        nm = name.encode() + b"\0"
        dt, sp = self._dtype_str(n), self._dataspace([len(strings)])
        pad = lambda b: b + b"\0" * (-len(b) % 8)   # noqa: E731
        body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + pad(nm) + pad(dt) + pad(sp)
        body += b"".join(s.ljust(n, b"\0") for s in strings)
        return self._msg(0x000C, body)

    def _write_dataset(self, arr) -> int:
        addr = self._alloc(arr.tobytes()) if arr.size else self.UNDEF
        layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
        fill = struct.pack("<BBBB", 2, 2, 0, 0)   # fill value v2: alloc late, write never, undefined
        return self._object_header([self._msg(0x0001, self._dataspace(arr.shape)), self._msg(0x0003, self._dtype_f32()),
                                    self._msg(0x0005, fill), self._msg(0x0008, layout)])

    def _write_group(self, node) -> int:
        entries = {}
        for name, sub in node["groups"].items():
            entries[name] = self._write_group(sub)
        for name, arr in node["datasets"].items():
            entries[name] = self._write_dataset(arr)
        names = sorted(entries)             # symbol table entries are ordered by name
        if len(names) > 2 * self.LEAF_K:
            raise ValueError("too many entries in one group for this writer")
        heap_data = bytearray(b"\0" * 8)    # offset 0 = empty string
        offs = {}
        for nme in names:
            offs[nme] = len(heap_data)
            heap_data += nme.encode() + b"\0"
            heap_data += b"\0" * (-len(heap_data) % 8)
        seg = self._alloc(bytes(heap_data))
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), self.UNDEF, seg))
        snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
        for nme in names:
            snod += struct.pack("<QQII16x", offs[nme], entries[nme], 0, 0)
        snod += b"\0" * (40 * (2 * self.LEAF_K - len(names)))
        snod_addr = self._alloc(snod)
        last_key = offs[names[-1]] if names else 0
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, self.UNDEF, self.UNDEF)
        tree += struct.pack("<QQQ", 0, snod_addr, last_key)
        tree += b"\0" * (16 * (2 * 16 - 1))      # room for 2K children, K = 16
        tree_addr = self._alloc(tree)
        msgs = [self._msg(0x0011, struct.pack("<QQ", tree_addr, heap))]
        msgs += [self._attr(k, v) for k, v in node["attrs"].items()]
        node["_btree"], node["_heap"] = tree_addr, heap
        return self._object_header(msgs)

    def save(self, path):
        root_hdr = self._write_group(self.root)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, 16, 0)
        sb += struct.pack("<QQQQ", 0, self.UNDEF, len(self.buf), self.UNDEF)
        sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", self.root["_btree"], self.root["_heap"])
        assert len(sb) == 96
        self.buf[:96] = sb
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "wb") as f:
            f.write(self.buf)
        return path


def mobilenet_weights(seed=1234, width=1.0, heads=False):
    """Seeded Keras-named MobileNet-v1 parameters: {'conv1/kernel': ..., 'conv1_bn/gamma': ..., ...}."""
    rs = np.random.RandomState(seed)
    w = {}

    def ch(c):
        return max(32, int(c * width) // 32 * 32)

    def bn(name, c):
        g, b, m, v = _bn_params(rs, c)
        w[f"{name}/gamma"], w[f"{name}/beta"], w[f"{name}/moving_mean"], w[f"{name}/moving_variance"] = g, b, m, v

    c0 = ch(32)
    w["conv1/kernel"] = (rs.standard_normal((3, 3, 3, c0)) * np.sqrt(2.0 / 27) / 60.0).astype(np.float32)
    bn("conv1_bn", c0)
    cin = c0
    for i, (cout, _) in enumerate(MOBILENET_CFG, start=1):
        cout = ch(cout)
        w[f"conv_dw_{i}/depthwise_kernel"] = (rs.standard_normal((3, 3, cin, 1)) * np.sqrt(2.0 / 9)).astype(np.float32)
        bn(f"conv_dw_{i}_bn", cin)
        w[f"conv_pw_{i}/kernel"] = (rs.standard_normal((1, 1, cin, cout)) * np.sqrt(2.0 / cin)).astype(np.float32)
        bn(f"conv_pw_{i}_bn", cout)
        cin = cout
    if heads:
        for name, (k, n) in (("feats", (cin, 256)), ("age_pred", (256, 100)), ("gender_pred", (256, 1))):
            w[f"{name}/kernel"] = (rs.standard_normal((k, n)) * np.sqrt(1.0 / k)).astype(np.float32)
            w[f"{name}/bias"] = rs.normal(0, 0.1, n).astype(np.float32)
    return w


def write_keras_mobilenet_h5(path, weights=None, seed=1234, full_model=True):
    """models/vgg2_mobilenet.h5 stand-in.  full_model=True mimics model.save() (weights under /model_weights, as
    facerec_keras_train.py:122 writes them); False mimics save_weights()."""
    weights = weights or mobilenet_weights(seed)
    h5 = H5Writer()
    prefix = "model_weights/" if full_model else ""
    layers = []
    for key in weights:
        layer = key.split("/")[0]
        if layer not in layers:
            layers.append(layer)
    for key, arr in weights.items():
        layer, wname = key.split("/")
        h5.dataset(f"{prefix}{layer}/{layer}/{wname}:0", arr)
    h5.attr_strings(prefix.rstrip("/"), "layer_names", layers)
    h5.attr_strings(prefix.rstrip("/"), "backend", ["tensorflow"])
    h5.attr_strings(prefix.rstrip("/"), "keras_version", ["2.2.4"])
    for layer in layers:
        names = [f"{k.split('/')[0]}/{k.split('/')[1]}:0" for k in weights if k.split("/")[0] == layer]
        h5.attr_strings(f"{prefix}{layer}", "weight_names", names)
    return h5.save(path)


def write_mobilenet_pb_from_weights(path, weights, input_hw=192):
    """The same parameters as a frozen GraphDef with FusedBatchNorm (eps 1e-3) - the oracle-side twin of the .h5."""
    g = GraphWriter()
    f = attr_type(DT_FLOAT)
    g.node("input_1", "Placeholder", dtype=f, shape=attr_shape([-1, input_hw, input_hw, 3]))

    def bn_relu6(layer, x):
        for s in ("gamma", "beta", "moving_mean", "moving_variance"):
            g.const(f"{layer}_bn/{s}", weights[f"{layer}_bn/{s}"])
        y = g.node(f"{layer}_bn/FusedBatchNorm", "FusedBatchNorm",
                   [x] + [f"{layer}_bn/{s}" for s in ("gamma", "beta", "moving_mean", "moving_variance")], T=f,
                   epsilon=attr_f(1e-3), is_training=attr_b(False), data_format=attr_s(b"NHWC"))
        return g.node(f"{layer}_relu/Relu6", "Relu6", [y], T=f)

    x = bn_relu6("conv1", _conv(g, "conv1/convolution", "input_1", weights["conv1/kernel"], 2, b"SAME"))
    for i, (_, stride) in enumerate(MOBILENET_CFG, start=1):
        g.const(f"conv_dw_{i}/depthwise_kernel", weights[f"conv_dw_{i}/depthwise_kernel"])
        y = g.node(f"conv_dw_{i}/depthwise", "DepthwiseConv2dNative", [x, f"conv_dw_{i}/depthwise_kernel"], T=f,
                   strides=attr_ints([1, stride, stride, 1]), padding=attr_s(b"SAME"), data_format=attr_s(b"NHWC"),
                   dilations=attr_ints([1, 1, 1, 1]))
        x = bn_relu6(f"conv_dw_{i}", y)
        x = bn_relu6(f"conv_pw_{i}", _conv(g, f"conv_pw_{i}/convolution", x, weights[f"conv_pw_{i}/kernel"], 1, b"SAME"))
    g.const("global_pooling/Mean/reduction_indices", np.array([1, 2], np.int32))
    x = g.node("global_pooling/Mean", "Mean", [x, "global_pooling/Mean/reduction_indices"], T=f, Tidx=attr_type(DT_INT32),
               keep_dims=attr_b(False))
    cin = weights["conv_pw_13/kernel"].shape[3]
    g.const("reshape_1/Reshape/shape", np.array([-1, 1, 1, cin], np.int32))
    g.node("reshape_1/Reshape", "Reshape", [x, "reshape_1/Reshape/shape"], T=f, Tshape=attr_type(DT_INT32))
    if "feats/kernel" in weights:
        def dense(name, xin, act):
            g.const(f"{name}/kernel", weights[f"{name}/kernel"])
            g.const(f"{name}/bias", weights[f"{name}/bias"])
            m = g.node(f"{name}/MatMul", "MatMul", [xin, f"{name}/kernel"], T=f, transpose_a=attr_b(False),
                       transpose_b=attr_b(False))
            a = g.node(f"{name}/BiasAdd", "BiasAdd", [m, f"{name}/bias"], T=f)
            return g.node(f"{name}/{act}", act, [a], T=f)
        feats = dense("feats", x, "Relu")
        dense("age_pred", feats, "Softmax")
        dense("gender_pred", feats, "Sigmoid")
    g.save(path)
    return path
