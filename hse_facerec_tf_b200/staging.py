"""GPU input staging (scope table row 8f-1): the per-face crop + resize that sits immediately in front of the hot path.

  expand_and_clamp_boxes   facial_analysis.py:236-263  (box + 10 px on every side, clamped to the frame)
  crop_resize              facial_analysis.py:267 + :95 (img[y1:y2, x1:x2] -> cv2.resize(..., (w, h)), INTER_LINEAR)
  resize_pil               facerec_test.py:84,93        (scipy.misc.imresize(img, size, interp='bilinear') = Pillow)
  load_resized_batch       facerec_test.py:80-93        (decode on the host, both resizes and the centre crop on the GPU)

Both resizes are bit-exact with the library the reference calls (OpenCV's uint8 INTER_LINEAR path, Pillow's BILINEAR
resample), so feeding their output to the network is the same as feeding the reference's own arrays.
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


def resize_pil(images, out_hw, device="cuda:0", crop=None):
    """scipy.misc.imresize(img, out_hw, interp='bilinear') for a batch, on the GPU, bit-exact with Pillow.
    images: a list of uint8 RGB arrays [H_i, W_i, 3] of arbitrary sizes (numpy; packed and uploaded in one copy), or one
    uint8 tensor / array [n, H, W, 3] (a CUDA tensor is used in place).  crop = (y0, x0, h, w): resize that window of
    every image instead (img[y0:y0+h, x0:x0+w], as in the reference's centre crop).  Returns CUDA uint8
    [n, out_h, out_w, 3]."""
    oh, ow = (out_hw, out_hw) if np.isscalar(out_hw) else out_hw
    dev = torch.device(device)
    if isinstance(images, (list, tuple)):
        arrs = [np.ascontiguousarray(a) for a in images]
        for a in arrs:
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("images must be uint8 [H,W,3]")
        sizes = [a.size for a in arrs]
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        host = torch.empty(int(offs[-1]), dtype=torch.uint8).pin_memory() if offs[-1] else torch.empty(0, dtype=torch.uint8)
        hv = host.numpy()
        for a, o in zip(arrs, offs):
            hv[o:o + a.size] = a.reshape(-1)
        buf = host.to(dev, non_blocking=True)
        desc = np.array([[offs[i], a.shape[0], a.shape[1], a.shape[1] * 3] for i, a in enumerate(arrs)], np.int64).reshape(-1, 4)
    else:
        t = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[3] != 3:
            raise ValueError("images must be uint8 [n,H,W,3]")
        buf = t.to(dev).contiguous()
        dev = buf.device
        n, H, W, _ = buf.shape
        desc = np.array([[i * H * W * 3, H, W, W * 3] for i in range(n)], np.int64).reshape(-1, 4)
    if crop is not None:
        y0, x0, ch, cw = [int(v) for v in crop]
        if len(desc) and (y0 < 0 or x0 < 0 or ch <= 0 or cw <= 0 or (desc[:, 1] < y0 + ch).any() or (desc[:, 2] < x0 + cw).any()):
            raise ValueError("crop window outside an image")
        desc[:, 0] += y0 * desc[:, 3] + x0 * 3
        desc[:, 1], desc[:, 2] = ch, cw
    n = len(desc)
    out = torch.empty((n, oh, ow, 3), dtype=torch.uint8, device=dev)
    if n:
        desc = np.ascontiguousarray(desc)
        with torch.cuda.device(dev):
            check(lib.hfr_resize_pil_u8(buf.data_ptr(), desc.ctypes.data, n, out.data_ptr(), oh, ow, dev.index or 0,
                                        _stream_ptr(dev)))
        torch.cuda.current_stream(dev).synchronize()   # `buf` / the pinned staging copy may be released after return
    return out


def load_resized_batch(paths, out_hw, crop_center=False, device="cuda:0"):
    """facerec_test.py:80-93 for a list of files: decode (host, PIL) -> [250x250 resize + centre 128x128 crop] -> resize
    to out_hw, the resizes on the GPU.  Returns CUDA uint8 [n, out_h, out_w, 3], identical to the per-file host path."""
    from PIL import Image
    imgs = []
    for p in paths:
        with Image.open(p) as im:
            imgs.append(np.asarray(im.convert("RGB")))
    if crop_center:
        orig_w, orig_h, w1, h1 = 250, 250, 128, 128
        dw, dh = (orig_w - w1) // 2, (orig_h - h1) // 2
        big = resize_pil(imgs, (orig_h, orig_w), device)                       # misc.imresize(img, (250, 250))
        return resize_pil(big, out_hw, device, crop=(dh, dw, orig_h - 2 * dh, orig_w - 2 * dw))   # img[dh:-dh, dw:-dw]
    return resize_pil(imgs, out_hw, device)
