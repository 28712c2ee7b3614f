"""ORACLE (test infrastructure, not product code): TensorFlow GraphDef wire-format reader.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product has its own C++ reader (hse_facerec_tf_b200/csrc/graphdef.cc).

Restates what the reference does at facerec_test.py:41-48 (`load_graph`: GraphDef.ParseFromString)
and age_gender_identity/facial_analysis.py:319-325 (`load_graph_def`), without TensorFlow: a plain
protobuf wire-format walk over the message layouts of tensorflow/core/framework/{graph,node_def,
attr_value,tensor,tensor_shape}.proto (TF 1.x; field numbers listed in SURVEY.md section 7).

Parity: unpinned by the reference (no TF in this image) - pinned by the known-answer tables in
SURVEY.md section 4, see tests/test_oracle_kat.py.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

# TF DataType enum -> numpy dtype (types.proto)
DTYPES = {
    1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8,
    9: np.int64, 10: np.bool_, 12: np.uint8,  # 12 = quint8, stored as raw bytes
}
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT64, DT_BOOL, DT_QUINT8 = 1, 2, 3, 4, 9, 10, 12


def _varint(buf: bytes, pos: int):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message; value is int or bytes."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val, wt):
    if wt == 0:
        return [val]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(v)
    return out


def parse_shape(buf: bytes):
    """TensorShapeProto{2: dim{1: size}, 3: unknown_rank} -> list[int] (-1 = unknown) or None."""
    dims = []
    for fno, wt, val in _fields(buf):
        if fno == 2:
            size = 0
            for f2, _, v2 in _fields(val):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif fno == 3 and val:
            return None
    return dims


def parse_tensor(buf: bytes):
    """TensorProto -> (np.ndarray, tf_dtype_enum).  Single *_val with a larger shape broadcasts."""
    dtype, shape, content = DT_FLOAT, [], None
    fvals, dvals, ivals, i64vals, bvals = [], [], [], [], []
    for fno, wt, val in _fields(buf):
        if fno == 1:
            dtype = val
        elif fno == 2:
            shape = parse_shape(val) or []
        elif fno == 4:
            content = val
        elif fno == 5:
            fvals += list(struct.unpack(f"<{len(val) // 4}f", val)) if wt in (2, 5) else []
        elif fno == 6:
            dvals += list(struct.unpack(f"<{len(val) // 8}d", val))
        elif fno == 7:
            ivals += [_signed64(v) for v in _packed_varints(val, wt)]
        elif fno == 10:
            i64vals += [_signed64(v) for v in _packed_varints(val, wt)]
        elif fno == 11:
            bvals += [bool(v) for v in _packed_varints(val, wt)]
    npdt = DTYPES[dtype]
    n = int(np.prod(shape)) if shape else 1
    if content:
        arr = np.frombuffer(content, dtype=npdt).copy()
    else:
        vals = fvals or dvals or ivals or i64vals or bvals or [0]
        arr = np.array(vals, dtype=npdt)
        if arr.size == 1 and n > 1:
            arr = np.full(n, arr[0], dtype=npdt)
        elif arr.size < n:  # TF pads with the last value
            arr = np.concatenate([arr, np.full(n - arr.size, arr[-1], dtype=npdt)])
    return arr.reshape(shape), dtype


def parse_attr(buf: bytes):
    """AttrValue -> python value."""
    for fno, wt, val in _fields(buf):
        if fno == 2:
            return val  # bytes (s)
        if fno == 3:
            return _signed64(val)
        if fno == 4:
            return struct.unpack("<f", val)[0]
        if fno == 5:
            return bool(val)
        if fno == 6:
            return ("type", val)
        if fno == 7:
            return ("shape", parse_shape(val))
        if fno == 8:
            return ("tensor",) + parse_tensor(val)
        if fno == 1:
            lst = {"s": [], "i": [], "f": [], "b": [], "type": [], "shape": []}
            for f2, w2, v2 in _fields(val):
                if f2 == 2:
                    lst["s"].append(v2)
                elif f2 == 3:
                    lst["i"] += [_signed64(v) for v in _packed_varints(v2, w2)]
                elif f2 == 4:
                    lst["f"] += list(struct.unpack(f"<{len(v2) // 4}f", v2))
                elif f2 == 5:
                    lst["b"] += [bool(v) for v in _packed_varints(v2, w2)]
                elif f2 == 6:
                    lst["type"] += _packed_varints(v2, w2)
                elif f2 == 7:
                    lst["shape"].append(parse_shape(v2))
            return ("list", lst)
    return None


@dataclass
class Node:
    name: str = ""
    op: str = ""
    inputs: list = field(default_factory=list)
    attrs: dict = field(default_factory=dict)

    def ints(self, key):
        a = self.attrs.get(key)
        return a[1]["i"] if a else None

    def tensor(self):
        a = self.attrs["value"]
        return a[1], a[2]


def parse_graphdef(data: bytes) -> list[Node]:
    """GraphDef{1: repeated NodeDef, 2: library (skipped), 4: versions (skipped)}."""
    nodes = []
    for fno, wt, val in _fields(data):
        if fno != 1:
            continue
        nd = Node()
        for f2, w2, v2 in _fields(val):
            if f2 == 1:
                nd.name = v2.decode()
            elif f2 == 2:
                nd.op = v2.decode()
            elif f2 == 3:
                nd.inputs.append(v2.decode())
            elif f2 == 5:
                key, attr = None, None
                for f3, w3, v3 in _fields(v2):
                    if f3 == 1:
                        key = v3.decode()
                    elif f3 == 2:
                        attr = parse_attr(v3)
                nd.attrs[key] = attr
        nodes.append(nd)
    return nodes


def load_graph(path: str) -> dict[str, Node]:
    with open(path, "rb") as f:
        nodes = parse_graphdef(f.read())
    return {n.name: n for n in nodes}
