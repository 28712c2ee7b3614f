"""GPU input staging (scope table row 8f-1): the per-face crop + resize that sits immediately in front of the hot path.

  expand_and_clamp_boxes   facial_analysis.py:236-263  (box + 10 px on every side, clamped to the frame)
  crop_resize              facial_analysis.py:267 + :95 (img[y1:y2, x1:x2] -> cv2.resize(..., (w, h)), INTER_LINEAR)

The resize is bit-exact with OpenCV's uint8 INTER_LINEAR path, so feeding its output to the network is the same as
feeding the reference's own crops.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib
from .model import _stream_ptr


def expand_and_clamp_boxes(bounding_boxes, img_h, img_w, dw=10, dh=10):
    """Integer boxes (x1, y1, x2, y2[, score]) -> list of [x1, y1, x2, y2] as process_image builds them; degenerate
    boxes (x2 <= x1 or y2 <= y1) are dropped, as in the reference."""
    out = []
    for b in bounding_boxes:
        x1, y1, x2, y2 = [int(bi) for bi in b[:4]]
        if x2 > x1 and y2 > y1:
            x1, x2 = x1 - dw, x2 + dw
            y1, y2 = y1 - dh, y2 + dh
            out.append([max(x1, 0), max(y1, 0), min(x2, img_w), min(y2, img_h)])
    return out


def crop_resize(frames, boxes, out_hw, frame_index=None):
    """frames: uint8 RGB [H,W,3] or [F,H,W,3] (numpy or CUDA tensor); boxes: [n,4] (x1,y1,x2,y2) already clamped;
    frame_index: [n] frame of each box (default 0).  Returns a CUDA uint8 tensor [n, out_h, out_w, 3]."""
    t = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
    if t.dtype != torch.uint8:
        raise ValueError("frames must be uint8")
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4 or t.shape[3] != 3:
        raise ValueError("frames must be [H,W,3] or [F,H,W,3]")
    t = t.cuda().contiguous() if not t.is_cuda else t.contiguous()
    F, H, W, _ = t.shape
    boxes = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
    n = len(boxes)
    fi = np.zeros(n, np.int64) if frame_index is None else np.asarray(frame_index, dtype=np.int64)
    if n and (boxes[:, 0].min() < 0 or boxes[:, 1].min() < 0 or boxes[:, 2].max() > W or boxes[:, 3].max() > H
              or (boxes[:, 2] <= boxes[:, 0]).any() or (boxes[:, 3] <= boxes[:, 1]).any() or fi.min() < 0 or fi.max() >= F):
        raise ValueError("boxes must be non-empty rectangles inside the frame")
    oh, ow = (out_hw, out_hw) if np.isscalar(out_hw) else out_hw
    out = torch.empty((n, oh, ow, 3), dtype=torch.uint8, device=t.device)
    if n == 0:
        return out
    b5 = torch.from_numpy(np.concatenate([fi[:, None], boxes], axis=1).astype(np.int32)).to(t.device)
    check(lib.hfr_crop_resize_u8(t.data_ptr(), F, H, W, b5.data_ptr(), n, out.data_ptr(), oh, ow, t.device.index or 0,
                                 _stream_ptr(t.device)))
    return out
