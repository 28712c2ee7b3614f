"""ORACLE (test infrastructure, not product code): CPU evaluation of a frozen TF-1.x GraphDef.

Restates, on torch-CPU ops, what `tf.Session.run` computes for the reference's hot path:
  * facerec_test.py:114-122   TensorFlowInference.extract_features -> sess.run(output, {input: x, phase: v})
  * facial_analysis.py:83-130 load_age_gender / age_gender_fun     -> sess.run([age, gender, feat])
TensorFlow itself (an un-vendored third-party dependency of the reference, version unpinned, 1.x API)
is not installed in this image, so the op semantics follow TF's published kernels:
  Conv2D / DepthwiseConv2dNative NHWC with SAME/VALID padding (pad_before = pad_total // 2),
  Dequantize mode=MIN_FIRST (tensorflow/core/kernels/dequantize_op.cc), Keras-style ReLU6 spelled
  Relu->Minimum->Maximum, Mean, MatMul, BiasAdd, Softmax, Sigmoid, FusedBatchNorm (inference),
  MaxPool/AvgPool, Pad, and Switch/Merge with dead-branch propagation.

Parity status: UNPINNED BY THE REFERENCE (TensorFlow cannot run here; the reference has no golden
vectors for this path).  Pinned instead against the survey-time known-answer tables (SURVEY.md
section 4) that two independent interpreters reproduced - tests/test_oracle_kat.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .graphdef import DT_QUINT8, Node, load_graph


class _Dead:
    """Marker for the untaken output of a Switch."""


DEAD = _Dead()


def same_pads(size: int, k: int, s: int, d: int = 1):
    """TF SAME padding: returns (out, pad_before, pad_after)."""
    out = -(-size // s)
    eff = (k - 1) * d + 1
    total = max((out - 1) * s + eff - size, 0)
    return out, total // 2, total - total // 2


def dequantize_min_first(q: np.ndarray, mn: float, mx: float, fp64: bool) -> np.ndarray:
    """TF Dequantize, mode=MIN_FIRST, T=quint8: w = round(min/s)*s + q*s with s=(max-min)/255.

    fp64=False does the three operations in float32 (what TF's Eigen kernel does); fp64=True does them
    in double and rounds once - SURVEY.md section 4 bounds the difference at 2e-4 relative."""
    if fp64:
        s = (float(mx) - float(mn)) / 255.0
        return (np.round(float(mn) / s) * s + q.astype(np.float64) * s).astype(np.float32)
    mn32, mx32 = np.float32(mn), np.float32(mx)
    s = (mx32 - mn32) / np.float32(255.0)
    off = np.float32(np.round(mn32 / s)) * s
    return (q.astype(np.float32) * s + off).astype(np.float32)


class GraphOracle:
    def __init__(self, path_or_nodes, dtype=torch.float32, dequant_fp64=False):
        self.nodes: dict[str, Node] = load_graph(path_or_nodes) if isinstance(path_or_nodes, str) else path_or_nodes
        self.dtype = dtype
        self.dequant_fp64 = dequant_fp64
        self._const_cache: dict = {}

    # ---- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _split(ref: str):
        if ref.startswith("^"):
            return None, 0
        if ":" in ref:
            name, port = ref.rsplit(":", 1)
            return name, int(port)
        return ref, 0

    def placeholder_shape(self, name: str):
        nd = self.nodes[self._split(name)[0]]
        sh = nd.attrs.get("shape")
        return sh[1] if sh else None

    def run(self, outputs, feeds: dict):
        """outputs: list of tensor names ('a/b:0'); feeds: {tensor name: array}.  Returns list of np arrays."""
        feeds = {self._split(k)[0]: v for k, v in feeds.items()}
        memo: dict = {}
        res = []
        for o in outputs:
            name, port = self._split(o)
            v = self._eval(name, memo, feeds)
            v = v[port] if isinstance(v, tuple) else v
            res.append(v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
        return res

    def _t(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        return t.to(self.dtype) if t.is_floating_point() else t

    def _in(self, nd: Node, i: int, memo, feeds):
        name, port = self._split(nd.inputs[i])
        v = self._eval(name, memo, feeds)
        return v[port] if isinstance(v, tuple) else v

    # ---- evaluator -----------------------------------------------------------------------------
    def _eval(self, name: str, memo, feeds):
        if name in memo:
            return memo[name]
        nd = self.nodes[name]
        if name in feeds:
            v = feeds[name]
            out = self._t(np.asarray(v)) if not isinstance(v, torch.Tensor) else v.to(self.dtype)
            if nd.attrs.get("dtype") == ("type", 10):
                out = torch.tensor(bool(np.asarray(v)))
            memo[name] = out
            return out
        out = self._op(nd, memo, feeds)
        memo[name] = out
        return out

    def _conv(self, nd, x, w, depthwise):
        strides = nd.ints("strides")
        dil = nd.ints("dilations") or [1, 1, 1, 1]
        pad = nd.attrs["padding"]
        assert nd.attrs.get("data_format", b"NHWC") == b"NHWC"
        sh, sw = strides[1], strides[2]
        dh, dw_ = dil[1], dil[2]
        kh, kw, cin, cm = w.shape
        x = x.permute(0, 3, 1, 2)
        if pad == b"SAME":
            _, pt, pb = same_pads(x.shape[2], kh, sh, dh)
            _, pl, pr = same_pads(x.shape[3], kw, sw, dw_)
            x = F.pad(x, (pl, pr, pt, pb))
        if depthwise:
            wt = w.permute(2, 3, 0, 1).reshape(cin * cm, 1, kh, kw)
            y = F.conv2d(x, wt, stride=(sh, sw), dilation=(dh, dw_), groups=cin)
        else:
            wt = w.permute(3, 2, 0, 1)
            y = F.conv2d(x, wt, stride=(sh, sw), dilation=(dh, dw_))
        return y.permute(0, 2, 3, 1).contiguous()

    def _pool(self, nd, x, kind):
        ks = nd.ints("ksize")
        st = nd.ints("strides")
        pad = nd.attrs["padding"]
        x = x.permute(0, 3, 1, 2)
        kh, kw, sh, sw = ks[1], ks[2], st[1], st[2]
        if pad == b"SAME":
            _, pt, pb = same_pads(x.shape[2], kh, sh)
            _, pl, pr = same_pads(x.shape[3], kw, sw)
        else:
            pt = pb = pl = pr = 0
        if kind == "max":
            x = F.pad(x, (pl, pr, pt, pb), value=-math.inf)
            y = F.max_pool2d(x, (kh, kw), (sh, sw))
        else:
            # TF AvgPool SAME divides by the number of valid (unpadded) elements
            ones = torch.ones_like(x[:1, :1])
            xs = F.avg_pool2d(F.pad(x, (pl, pr, pt, pb)), (kh, kw), (sh, sw), divisor_override=1)
            cnt = F.avg_pool2d(F.pad(ones, (pl, pr, pt, pb)), (kh, kw), (sh, sw), divisor_override=1)
            y = xs / cnt
        return y.permute(0, 2, 3, 1).contiguous()

    def _op(self, nd: Node, memo, feeds):
        op = nd.op
        data_inputs = [i for i in nd.inputs if not i.startswith("^")]
        if op == "Merge":
            live = []
            for i in range(len(data_inputs)):
                v = self._in(nd, i, memo, feeds)
                if v is not DEAD:
                    live.append(v)
            return live[0] if live else DEAD
        args = [self._in(nd, i, memo, feeds) for i in range(len(data_inputs))]
        if op != "Switch" and any(a is DEAD for a in args):
            return DEAD
        if op == "Const":
            arr, dt = nd.tensor()
            return arr if dt == DT_QUINT8 else self._t(arr)
        if op in ("Identity", "StopGradient", "PlaceholderWithDefault"):
            return args[0]
        if op == "Placeholder":
            raise KeyError(f"placeholder {nd.name} not fed")
        if op == "Dequantize":
            assert nd.attrs.get("mode") == b"MIN_FIRST", nd.attrs.get("mode")
            q, mn, mx = args
            return self._t(dequantize_min_first(np.asarray(q), float(mn), float(mx), self.dequant_fp64))
        if op == "Switch":
            data, pred = args
            if data is DEAD:
                return (DEAD, DEAD)
            p = bool(pred)
            return (DEAD, data) if p else (data, DEAD)
        if op == "Conv2D":
            return self._conv(nd, args[0], args[1], False)
        if op == "DepthwiseConv2dNative":
            return self._conv(nd, args[0], args[1], True)
        if op in ("Add", "AddV2", "BiasAdd"):
            return args[0] + args[1]
        if op == "Sub":
            return args[0] - args[1]
        if op == "Mul":
            return args[0] * args[1]
        if op == "RealDiv":
            return args[0] / args[1]
        if op == "Rsqrt":
            return torch.rsqrt(args[0])
        if op == "Sqrt":
            return torch.sqrt(args[0])
        if op == "Neg":
            return -args[0]
        if op == "Exp":
            return torch.exp(args[0])
        if op in ("Max", "Sum"):       # reductions of the hand-spelled softmax in mtcnn.pb (Max -> Sub -> Exp -> Sum -> RealDiv)
            axes = [int(a) for a in np.asarray(args[1]).reshape(-1)]
            keep = bool(nd.attrs.get("keep_dims", False))
            return args[0].amax(dim=axes, keepdim=keep) if op == "Max" else args[0].sum(dim=axes, keepdim=keep)
        if op == "Relu":
            return torch.relu(args[0])
        if op == "Relu6":
            return torch.clamp(args[0], 0.0, 6.0)
        if op == "Minimum":
            return torch.minimum(args[0], args[1])
        if op == "Maximum":
            return torch.maximum(args[0], args[1])
        if op == "Mean":
            axes = [int(a) for a in np.asarray(args[1]).reshape(-1)]
            return args[0].mean(dim=axes, keepdim=bool(nd.attrs.get("keep_dims", False)))
        if op == "MatMul":
            a, b = args
            if nd.attrs.get("transpose_a"):
                a = a.t()
            if nd.attrs.get("transpose_b"):
                b = b.t()
            return a @ b
        if op == "Softmax":
            return torch.softmax(args[0], dim=-1)
        if op == "Sigmoid":
            return torch.sigmoid(args[0])
        if op == "Reshape":
            shape = [int(s) for s in np.asarray(args[1]).reshape(-1)]
            return args[0].reshape(shape)
        if op == "Squeeze":
            dims = nd.ints("squeeze_dims") or []
            x = args[0]
            for d in sorted(dims, reverse=True):
                x = x.squeeze(d)
            return x if dims else x.squeeze()
        if op == "Pad":
            p = np.asarray(args[1]).reshape(-1, 2)
            assert p[0].sum() == 0 and p[3].sum() == 0
            x = args[0].permute(0, 3, 1, 2)
            x = F.pad(x, (int(p[2, 0]), int(p[2, 1]), int(p[1, 0]), int(p[1, 1])))
            return x.permute(0, 2, 3, 1).contiguous()
        if op == "MaxPool":
            return self._pool(nd, args[0], "max")
        if op == "AvgPool":
            return self._pool(nd, args[0], "avg")
        if op in ("FusedBatchNorm", "FusedBatchNormV2", "FusedBatchNormV3"):
            assert not nd.attrs.get("is_training", True), "training-mode FusedBatchNorm"
            x, g, b, m, v = args
            eps = nd.attrs.get("epsilon", 1e-3)
            y = (x - m) * torch.rsqrt(v + eps) * g + b
            return (y, m, v, m, v, m)
        raise NotImplementedError(f"oracle: op {op} ({nd.name})")


# ---- reference post-/pre-processing restatements ------------------------------------------------

IMAGENET_BGR_MEAN = (103.939, 116.779, 123.68)      # facerec_test.py:99-102, facial_analysis.py:105-107
VGGFACE2_BGR_MEAN = (91.4953, 103.8827, 131.0912)   # facerec_test.py:103-106


def preprocess_rgb_u8(img_u8: np.ndarray, convert2BGR=True, imageNetUtilsMean=True, dtype=np.float32):
    """facerec_test.py:96-110 on an already-resized RGB uint8 array (..., H, W, 3)."""
    x = img_u8.astype(np.float64)
    if convert2BGR:
        x = x[..., ::-1].copy()
        mean = IMAGENET_BGR_MEAN if imageNetUtilsMean else VGGFACE2_BGR_MEAN
        x[..., 0] -= mean[0]
        x[..., 1] -= mean[1]
        x[..., 2] -= mean[2]
    else:
        x /= 127.5
        x -= 1.0
    return x.astype(dtype)


def age_from_probs(age_preds: np.ndarray):
    """facial_analysis.py:113-124: top-2 expectation, min_age=1.  age_preds: (100,)."""
    indices = age_preds.argsort()[::-1][:2]
    norm_preds = age_preds[indices] / np.sum(age_preds[indices])
    res_age = 1
    for age, p in zip(indices, norm_preds):
        res_age += age * p
    return float(res_age), indices


def is_male(gender_preds):
    """facial_analysis.py:76-81 (use_sota=False branch)."""
    return gender_preds >= 0.6
