"""Python handle over hfr_model: one compiled network (weights + activation arena) on one GPU."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class HfrModel:
    """Loads a frozen GraphDef (.pb) / Keras .h5 and runs batched forwards through libhfr.so.

    device=None compiles on the host only (plan inspection; no GPU required)."""

    def __init__(self, path, input_tensor, output_tensors, learning_phase_tensor=None, additional_input_value=0,
                 input_hw=0, device="cuda:0", precision="bf16"):
        if isinstance(output_tensors, str):
            output_tensors = [output_tensors]
        self.output_tensors = list(output_tensors)
        self.precision = precision
        self.device = None if device is None else torch.device(device)
        dev_index = -1 if self.device is None else (self.device.index or 0)
        if self.device is not None and not torch.cuda.is_available():
            raise _lib.HfrError("no CUDA device available; this library has no CPU fallback")
        h = C.c_void_p()
        check(lib.hfr_model_load(str(path).encode(), (input_tensor or "").encode(),
                                 ",".join(self.output_tensors).encode(),
                                 learning_phase_tensor.encode() if learning_phase_tensor else None,
                                 float(additional_input_value), int(input_hw), dev_index, _lib.PREC[precision],
                                 C.byref(h)))
        self._h = h
        ih, iw, ic, no = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        dims = (C.c_int * 16)()
        check(lib.hfr_model_info(h, C.byref(ih), C.byref(iw), C.byref(ic), C.byref(no), dims))
        self.h, self.w, self.c = ih.value, iw.value, ic.value
        self.out_dims = [dims[i] for i in range(no.value)]

    def close(self):
        if getattr(self, "_h", None):
            lib.hfr_model_free(self._h)
            self._h = None

    __del__ = close

    def plan(self) -> dict:
        n = lib.hfr_model_plan_json(self._h, None, 0)
        buf = C.create_string_buffer(int(n))
        lib.hfr_model_plan_json(self._h, buf, n)
        return json.loads(buf.value.decode())

    def layer_weights(self, layer: int):
        n = lib.hfr_model_layer_weights(self._h, layer, None, 0, None, 0)
        check(int(n))
        cout = self.plan()["layers"][layer]["cout"]
        w = np.zeros(int(n), np.float32)
        b = np.zeros(cout, np.float32)
        lib.hfr_model_layer_weights(self._h, layer, w.ctypes.data, w.size, b.ctypes.data, b.size)
        return w, b

    # ---- device path ---------------------------------------------------------------------------
    def _flags(self, convert2BGR, imageNetUtilsMean, l2norm, graph, u8):
        f = 0
        if u8:
            if convert2BGR:
                f |= _lib.FLAG_BGR | (_lib.FLAG_MEAN_IMAGENET if imageNetUtilsMean else _lib.FLAG_MEAN_VGGFACE2)
            else:
                f |= _lib.FLAG_SCALE_PM1
        if l2norm:
            f |= _lib.FLAG_L2NORM
        if graph:
            f |= _lib.FLAG_CUDA_GRAPH
        return f

    def forward(self, x: torch.Tensor, convert2BGR=True, imageNetUtilsMean=True, l2norm=False, graph=False, outs=None):
        """x: CUDA tensor [B,H,W,3], uint8 RGB crops (pre-processing fused on the GPU) or float32 already
        pre-processed (what the reference feeds the placeholder).  Returns a list of float32 CUDA tensors."""
        if not (x.is_cuda and x.is_contiguous()):
            raise ValueError("x must be a contiguous CUDA tensor")
        if x.device != self.device:
            raise ValueError(f"x lives on {x.device}, the model on {self.device}")
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.h, self.w, self.c):
            raise ValueError(f"expected input [B,{self.h},{self.w},{self.c}], got {tuple(x.shape)}")
        if x.dtype == torch.uint8:
            dt = _lib.IN_U8
        elif x.dtype == torch.float32:
            dt = _lib.IN_F32
        else:
            raise ValueError("x must be uint8 or float32")
        B = x.shape[0]
        if outs is None:
            outs = [torch.empty((B, d), dtype=torch.float32, device=x.device) for d in self.out_dims]
        elif any(o.device != self.device or o.dtype != torch.float32 or not o.is_contiguous() for o in outs):
            raise ValueError("outs must be contiguous float32 tensors on the model's device")
        ptrs = (C.c_void_p * len(outs))(*[o.data_ptr() for o in outs])
        flags = self._flags(convert2BGR, imageNetUtilsMean, l2norm, graph, dt == _lib.IN_U8)
        with torch.cuda.device(self.device):     # the library selects the model's device; torch's current one is restored
            check(lib.hfr_model_forward(self._h, x.data_ptr(), dt, B, flags, ptrs, _stream_ptr(x.device)))
        return outs

    def forward_host(self, x: np.ndarray, convert2BGR=True, imageNetUtilsMean=True, l2norm=False, graph=False,
                     outs=None):
        """numpy in / numpy out (the reference's calling convention); H2D and D2H copies included."""
        x = np.ascontiguousarray(x)
        if x.ndim != 4 or tuple(x.shape[1:]) != (self.h, self.w, self.c):
            raise ValueError(f"expected input [B,{self.h},{self.w},{self.c}], got {tuple(x.shape)}")
        if x.dtype == np.uint8:
            dt = _lib.IN_U8
        else:
            x = x.astype(np.float32, copy=False)  # TF casts the fp64 feed to the fp32 placeholder
            dt = _lib.IN_F32
        B = x.shape[0]
        if outs is None:
            outs = [np.empty((B, d), np.float32) for d in self.out_dims]
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        flags = self._flags(convert2BGR, imageNetUtilsMean, l2norm, graph, dt == _lib.IN_U8)
        with torch.cuda.device(self.device):
            check(lib.hfr_model_forward_host(self._h, x.ctypes.data, dt, B, flags, ptrs, _stream_ptr(self.device)))
        return outs

    HOST_SLOTS = 4

    def submit_host(self, slot, x: np.ndarray, outs, convert2BGR=True, imageNetUtilsMean=True, l2norm=False, graph=True):
        """Asynchronous half of forward_host: enqueue upload + forward + download of one batch on `slot`; `outs` (float32
        [B, out_dims[i]] arrays, ideally page-locked like `x`) are valid after wait_host(slot).  The arrays must stay
        alive and untouched until then."""
        if x.ndim != 4 or tuple(x.shape[1:]) != (self.h, self.w, self.c) or not x.flags["C_CONTIGUOUS"]:
            raise ValueError(f"expected a contiguous input [B,{self.h},{self.w},{self.c}], got {tuple(x.shape)}")
        if x.dtype == np.uint8:
            dt = _lib.IN_U8
        elif x.dtype == np.float32:
            dt = _lib.IN_F32
        else:
            raise ValueError("submit_host takes uint8 or float32 batches")
        B = x.shape[0]
        for o, d in zip(outs, self.out_dims):
            if o.dtype != np.float32 or tuple(o.shape) != (B, d) or not o.flags["C_CONTIGUOUS"]:
                raise ValueError("outs must be contiguous float32 arrays of shape [B, out_dim]")
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        flags = self._flags(convert2BGR, imageNetUtilsMean, l2norm, graph, dt == _lib.IN_U8)
        with torch.cuda.device(self.device):
            check(lib.hfr_model_submit_host(self._h, int(slot), x.ctypes.data, dt, B, flags, ptrs, _stream_ptr(self.device)))

    def wait_host(self, slot):
        check(lib.hfr_model_wait_host(self._h, int(slot)))

    def stream_host(self, batches, depth=2, **kw):
        """Generator over an iterable of host batches -> lists of page-locked float32 outputs, `depth` batches in flight
        (upload of batch i+1 and download of batch i-1 overlap the compute of batch i).  A yielded list is valid ONLY
        until the next item is requested: its slot (and its pinned output arrays) is resubmitted at that moment - copy
        what must outlive that."""
        depth = max(1, min(int(depth), self.HOST_SLOTS))
        pinned, pending, bufs = {}, [], {}
        for i, x in enumerate(batches):
            slot = i % depth
            if len(pending) == depth:
                s0 = pending.pop(0)
                self.wait_host(s0)
                yield bufs[s0]
            x = np.ascontiguousarray(x)
            if x.dtype != np.uint8:
                x = x.astype(np.float32, copy=False)
            key = (slot, x.shape, x.dtype.str)
            if key not in pinned:   # page-locked staging copies of the input and the outputs, one set per slot
                pinned[key] = (torch.empty(x.shape, dtype=torch.uint8 if x.dtype == np.uint8 else torch.float32).pin_memory(),
                               [torch.empty((x.shape[0], d), dtype=torch.float32).pin_memory() for d in self.out_dims])
            hx, ho = pinned[key]
            hx.numpy()[...] = x
            bufs[slot] = [o.numpy() for o in ho]
            self.submit_host(slot, hx.numpy(), bufs[slot], **kw)
            pending.append(slot)
        for s0 in pending:
            self.wait_host(s0)
            yield bufs[s0]

    # ---- per-layer timing ----------------------------------------------------------------------
    def layer_timing(self, enable=True):
        check(lib.hfr_model_set_layer_timing(self._h, int(enable)))

    def layer_times(self):
        """(ms per layer summed over the timed steps, steps)"""
        n = len(self.plan()["layers"])
        ms = (C.c_double * n)()
        steps = C.c_int()
        check(lib.hfr_model_get_layer_times(self._h, ms, C.byref(steps)))
        return [ms[i] for i in range(n)], steps.value

    # ---- debugging -----------------------------------------------------------------------------
    def keep_activations(self, keep=True):
        check(lib.hfr_model_set_keep_activations(self._h, int(keep)))

    def layer_output(self, layer: int, batch: int) -> torch.Tensor:
        L = self.plan()["layers"][layer]
        n = L["hw_out"][0] * L["hw_out"][1] * L["cout"]
        dst = torch.empty((batch, n), dtype=torch.float32, device=self.device)
        check(int(lib.hfr_model_debug_layer(self._h, layer, batch, dst.data_ptr(), _stream_ptr(self.device))))
        return dst.view(batch, L["hw_out"][0], L["hw_out"][1], L["cout"])
