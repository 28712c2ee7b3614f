"""MTCNN face detection (scope row 8f-4): drop-in for the detector half of FacialImageProcessing.

  load_mtcnn -> (pnet_fun, rnet_fun, onet_fun)   facial_analysis.py:334-352   the three networks run in libhfr.so
  mtcnn_detect_faces(img)                         facial_analysis.py:478-604   pyramid -> P-Net -> R-Net -> O-Net
  bbreg / generateBoundingBox / nms / pad / rerec facial_analysis.py:354-476   box bookkeeping, numpy on the host

As in the reference the cascade is host logic around three network calls; what changes is that every network call is a
batch on the GPU (hfr_mtcnn_run): one launch sequence per pyramid level for P-Net, ONE batch of all candidates for R-Net
and for O-Net.  Image resampling stays cv2.INTER_AREA on the host (it is what the reference calls; the detector is
upstream of the hot path).  The result feeds FacialImageProcessing.process_image through `detector=MTCNN(...)`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib
from .model import _stream_ptr

PNET, RNET, ONET = 0, 1, 2
STEP_THRESHOLDS = (0.6, 0.7, 0.9)      # facial_analysis.py:480
PYRAMID_FACTOR = 0.709                 # facial_analysis.py:481


class MTCNN:
    """detector = MTCNN("mtcnn.pb");  boxes, points = detector(img_rgb)   (boxes [n, 5], points [10, n])"""

    def __init__(self, model_file="mtcnn.pb", minsize=32, device="cuda:0"):
        if not torch.cuda.is_available():
            from ._lib import HfrError
            raise HfrError("no CUDA device available; the detector networks have no CPU fallback")
        self.device = torch.device(device)
        self.minsize = minsize
        h = C.c_void_p()
        check(lib.hfr_mtcnn_load(str(model_file).encode(), self.device.index or 0, C.byref(h)))
        self._h = h
        # the reference's three lambdas (numpy in, tuple of numpy out)
        self.pnet = lambda img: self._run(PNET, img)
        self.rnet = lambda img: self._run(RNET, img)
        self.onet = lambda img: self._run(ONET, img)

    def close(self):
        if getattr(self, "_h", None):
            lib.hfr_mtcnn_free(self._h)
            self._h = None

    __del__ = close

    def _out_shape(self, net, h, w, slot):
        oh, ow, oc = C.c_int(), C.c_int(), C.c_int()
        check(lib.hfr_mtcnn_out_shape(self._h, net, h, w, slot, C.byref(oh), C.byref(ow), C.byref(oc)))
        return oh.value, ow.value, oc.value

    def _run(self, net, x):
        """x: float array [n, h, w, 3] as fed to the placeholder -> tuple of float32 arrays (the bound output tensors)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4 or x.shape[3] != 3:
            raise ValueError(f"expected [n, h, w, 3], got {x.shape}")
        n, h, w, _ = x.shape
        nout = 3 if net == ONET else 2
        shapes = [self._out_shape(net, h, w, s) for s in range(nout)]
        with torch.cuda.device(self.device):
            xd = torch.from_numpy(x).to(self.device)
            outs = [torch.empty((n,) + s, dtype=torch.float32, device=self.device) for s in shapes]
            ptr = [o.data_ptr() for o in outs] + [None] * (3 - nout)
            check(lib.hfr_mtcnn_run(self._h, net, xd.data_ptr(), n, h, w, ptr[0], ptr[1], ptr[2], _stream_ptr(self.device)))
            res = [o.cpu().numpy() for o in outs]
        if net != PNET:                                   # the dense heads are [n, c]
            res = [r.reshape(n, -1) for r in res]
        return tuple(res)

    # ---- the cascade -------------------------------------------------------------------------------------------------
    def __call__(self, img):
        return self.detect_faces(img)

    def detect_faces(self, img, minsize=None):
        """img: RGB uint8 [H, W, 3].  Returns (total_boxes [n, 5] float: x1, y1, x2, y2, score; points [10, n])."""
        return run_cascade(self, img, self.minsize if minsize is None else minsize)


def run_cascade(nets, img, minsize):
    """The cascade of mtcnn_detect_faces around any object with the reference's three network callables
    (pnet / rnet / onet: float [n, h, w, 3] in, tuple of arrays out)."""
    import cv2
    img_h, img_w = img.shape[0], img.shape[1]
    cand = np.empty((0, 9))
    for scale in _scales(img_h, img_w, minsize):
        hs, ws = int(np.ceil(img_h * scale)), int(np.ceil(img_w * scale))
        level = (cv2.resize(img, (ws, hs), interpolation=cv2.INTER_AREA) - 127.5) * 0.0078125
        reg, prob = nets.pnet(np.transpose(level[None], (0, 2, 1, 3)))            # width-major, as the reference feeds it
        reg, prob = reg.transpose(0, 2, 1, 3)[0], prob.transpose(0, 2, 1, 3)[0, :, :, 1]
        boxes = _heatmap_boxes(prob, reg, scale, STEP_THRESHOLDS[0])
        keep = non_max_suppression(boxes, 0.5, "Union")
        if boxes.size and keep.size:
            cand = np.vstack([cand, boxes[keep]])
    points = np.array([])
    if len(cand):
        cand = cand[non_max_suppression(cand, 0.7, "Union")]
        bw, bh = cand[:, 2] - cand[:, 0], cand[:, 3] - cand[:, 1]
        cand = np.column_stack([cand[:, 0] + cand[:, 5] * bw, cand[:, 1] + cand[:, 6] * bh,
                                cand[:, 2] + cand[:, 7] * bw, cand[:, 3] + cand[:, 8] * bh, cand[:, 4]])
        cand = to_square(cand)
        cand[:, :4] = np.fix(cand[:, :4]).astype(np.int32)
    if len(cand):
        reg, prob = nets.rnet(_candidate_crops(img, cand, 24))
        score = prob[:, 1]
        ok = np.flatnonzero(score > STEP_THRESHOLDS[1])
        cand = np.column_stack([cand[ok, :4], score[ok]])
        if len(cand):
            keep = non_max_suppression(cand, 0.7, "Union")
            cand = to_square(box_regression(cand[keep], reg[ok][keep]))
    if len(cand):
        cand = np.fix(cand).astype(np.int32)
        reg, pts, prob = nets.onet(_candidate_crops(img, cand, 48))
        score = prob[:, 1]
        ok = np.flatnonzero(score > STEP_THRESHOLDS[2])
        cand = np.column_stack([cand[ok, :4].astype(np.float64), score[ok]])
        reg = reg[ok]
        points = pts[ok].T.astype(np.float64)
        bw, bh = cand[:, 2] - cand[:, 0] + 1, cand[:, 3] - cand[:, 1] + 1
        points[:5] = points[:5] * bw + cand[:, 0] - 1
        points[5:] = points[5:] * bh + cand[:, 1] - 1
        if len(cand):
            cand = box_regression(cand, reg)
            keep = non_max_suppression(cand, 0.7, "Min")
            cand, points = cand[keep], points[:, keep]
    return cand, points


def _scales(h, w, minsize):
    """facial_analysis.py:487-496: pyramid levels from 12 / minsize down to a 12-pixel shorter side."""
    base = 12.0 / minsize
    shorter = min(h, w) * base
    out = []
    while shorter >= 12:
        out.append(base * PYRAMID_FACTOR ** len(out))
        shorter *= PYRAMID_FACTOR
    return out


def _heatmap_boxes(prob, reg, scale, threshold):
    """facial_analysis.py:370-398: every P-Net cell above the threshold becomes a 12 x 12 box at stride 2, mapped back to
    the image (MATLAB-style 1-based corners, np.fix), with its score and regression offsets: rows [x1 y1 x2 y2 s dx1 dy1
    dx2 dy2].  prob [H', W'], reg [H', W', 4] in image orientation; the reference works on the transposes."""
    pt = prob.T
    ii, jj = np.nonzero(pt >= threshold)
    planes = [reg[:, :, k].T for k in range(4)]
    if ii.shape[0] == 1:                       # quirk kept from the reference: a single hit reads the flipped planes
        planes = [np.flipud(p) for p in planes]
    cells = np.column_stack([ii, jj])
    top_left = np.fix((2 * cells + 1) / scale)
    bottom_right = np.fix((2 * cells + 12) / scale)
    offs = np.column_stack([p[ii, jj] for p in planes]) if ii.size else np.empty((0, 4))
    return np.hstack([top_left, bottom_right, pt[ii, jj][:, None], offs])


def non_max_suppression(boxes, threshold, method):
    """facial_analysis.py:401-433: greedy NMS from the best score down; overlap = intersection over union, or over the
    smaller box ('Min').  Returns the indices kept, best first."""
    if boxes.size == 0:
        return np.empty((0,), dtype=np.int64)
    x1, y1, x2, y2, s = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3], boxes[:, 4]
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = np.argsort(s)
    kept = []
    while order.size:
        best, order = order[-1], order[:-1]
        kept.append(best)
        iw = np.maximum(0.0, np.minimum(x2[best], x2[order]) - np.maximum(x1[best], x1[order]) + 1)
        ih = np.maximum(0.0, np.minimum(y2[best], y2[order]) - np.maximum(y1[best], y1[order]) + 1)
        inter = iw * ih
        denom = np.minimum(area[best], area[order]) if method == "Min" else area[best] + area[order] - inter
        order = order[inter / denom <= threshold]
    return np.asarray(kept, dtype=np.int64)


def to_square(boxes):
    """facial_analysis.py:468-476: grow the shorter side around the centre."""
    boxes = boxes.copy()
    w, h = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    side = np.maximum(w, h)
    boxes[:, 0] += 0.5 * w - 0.5 * side
    boxes[:, 1] += 0.5 * h - 0.5 * side
    boxes[:, 2] = boxes[:, 0] + side
    boxes[:, 3] = boxes[:, 1] + side
    return boxes


def box_regression(boxes, reg):
    """facial_analysis.py:355-367: corners move by the predicted fractions of the (inclusive) width / height."""
    boxes = boxes.copy()
    w, h = boxes[:, 2] - boxes[:, 0] + 1, boxes[:, 3] - boxes[:, 1] + 1
    boxes[:, 0] += reg[:, 0] * w
    boxes[:, 1] += reg[:, 1] * h
    boxes[:, 2] += reg[:, 2] * w
    boxes[:, 3] += reg[:, 3] * h
    return boxes


def _candidate_crops(img, boxes, size):
    """facial_analysis.py:437-465 + 537-546: every (possibly out-of-image) box is cut out of the frame into a zero canvas
    of its own size, resized to size x size with cv2.INTER_AREA, normalised and handed over width-major [n, size, size, 3]."""
    import cv2
    img_h, img_w = img.shape[0], img.shape[1]
    out = np.zeros((len(boxes), size, size, 3))
    for k, b in enumerate(boxes):
        x1, y1, x2, y2 = (int(v) for v in b[:4])                  # 1-based inclusive corners
        cw, ch = x2 - x1 + 1, y2 - y1 + 1
        canvas = np.zeros((ch, cw, 3))
        sx1, sy1, sx2, sy2 = max(x1, 1), max(y1, 1), min(x2, img_w), min(y2, img_h)
        dx1, dy1 = sx1 - x1, sy1 - y1                             # where the visible part starts inside the canvas
        canvas[dy1:dy1 + (sy2 - sy1 + 1), dx1:dx1 + (sx2 - sx1 + 1)] = img[sy1 - 1:sy2, sx1 - 1:sx2]
        out[k] = cv2.resize(canvas, (size, size), interpolation=cv2.INTER_AREA)
    return np.transpose((out - 127.5) * 0.0078125, (0, 2, 1, 3))
