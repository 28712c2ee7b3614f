"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's MTCNN face detector.

  FacialImageProcessing.load_mtcnn          facial_analysis.py:334-352   three sess.run lambdas over mtcnn.pb
  FacialImageProcessing.mtcnn_detect_faces  facial_analysis.py:478-604   image pyramid -> P-Net -> R-Net -> O-Net
  bbreg / generateBoundingBox / nms / pad / rerec                        facial_analysis.py:354-476

The three networks are evaluated by oracle/tfnet.GraphOracle on the reference's own mtcnn.pb (a copy travels as
tests/golden/mtcnn.pb); the cascade around them is restated here in numpy, quirks included: images enter the networks
transposed (W before H), box corners use the MATLAB-style +1 conventions, np.fix truncation, NMS picks by ascending
argsort of the scores, the O-Net stage merges with the 'Min' overlap rule.

Parity status: pinned by the reference's notebook (AgeGenderIdentityDemo.ipynb cell 7: FOUR faces found in
test_image.jpg with minsize=32 defaults) - tests/test_oracle_kat.py checks the count; TensorFlow itself cannot run here.
"""
from __future__ import annotations

import cv2
import numpy as np

from .tfnet import GraphOracle

THRESHOLDS = (0.6, 0.7, 0.9)   # facial_analysis.py:480
FACTOR = 0.709                 # facial_analysis.py:481


class MtcnnOracle:
    def __init__(self, pb_path: str):
        self.g = GraphOracle(pb_path)

    # facial_analysis.py:349-351: the three lambdas (inputs float NHWC, already normalised and transposed)
    def pnet(self, x):
        return self.g.run(["pnet/conv4-2/BiasAdd:0", "pnet/prob1:0"], {"pnet/input:0": x.astype(np.float32)})

    def rnet(self, x):
        return self.g.run(["rnet/conv5-2/conv5-2:0", "rnet/prob1:0"], {"rnet/input:0": x.astype(np.float32)})

    def onet(self, x):
        return self.g.run(["onet/conv6-2/conv6-2:0", "onet/conv6-3/conv6-3:0", "onet/prob1:0"],
                          {"onet/input:0": x.astype(np.float32)})


def pyramid_scales(h, w, minsize):
    """facial_analysis.py:487-496"""
    m = 12.0 / minsize
    minl = min(h, w) * m
    scales, k = [], 0
    while minl >= 12:
        scales.append(m * FACTOR ** k)
        minl *= FACTOR
        k += 1
    return scales


def boxes_from_heatmap(prob, reg, scale, thr):
    """facial_analysis.py:370-398 (generateBoundingBox): prob [H', W'], reg [H', W', 4] in image orientation."""
    stride, cell = 2, 12
    pm = prob.T
    d = [reg[:, :, i].T for i in range(4)]
    ys, xs = np.where(pm >= thr)
    if ys.shape[0] == 1:
        d = [np.flipud(v) for v in d]
    score = pm[ys, xs]
    offs = np.stack([v[ys, xs] for v in d], axis=1) if ys.size else np.empty((0, 4))
    bb = np.stack([ys, xs], axis=1)
    q1 = np.fix((stride * bb + 1) / scale)
    q2 = np.fix((stride * bb + cell - 1 + 1) / scale)
    return np.hstack([q1, q2, score[:, None], offs])


def nms(boxes, thr, method):
    """facial_analysis.py:401-433"""
    if boxes.size == 0:
        return np.empty((0,), dtype=np.int64)
    x1, y1, x2, y2, s = (boxes[:, i] for i in range(5))
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = np.argsort(s)
    keep = []
    while order.size > 0:
        i = order[-1]
        keep.append(i)
        rest = order[:-1]
        w = np.maximum(0.0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
        h = np.maximum(0.0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        inter = w * h
        o = inter / np.minimum(area[i], area[rest]) if method == "Min" else inter / (area[i] + area[rest] - inter)
        order = rest[o <= thr]
    return np.asarray(keep, dtype=np.int64)


def square(b):
    """facial_analysis.py:468-476 (rerec)"""
    h, w = b[:, 3] - b[:, 1], b[:, 2] - b[:, 0]
    side = np.maximum(w, h)
    b[:, 0] = b[:, 0] + w * 0.5 - side * 0.5
    b[:, 1] = b[:, 1] + h * 0.5 - side * 0.5
    b[:, 2:4] = b[:, 0:2] + side[:, None]
    return b


def regress(b, reg):
    """facial_analysis.py:355-367 (bbreg)"""
    w, h = b[:, 2] - b[:, 0] + 1, b[:, 3] - b[:, 1] + 1
    b[:, 0:4] = np.stack([b[:, 0] + reg[:, 0] * w, b[:, 1] + reg[:, 1] * h, b[:, 2] + reg[:, 2] * w, b[:, 3] + reg[:, 3] * h], 1)
    return b


def crop_windows(b, w, h):
    """facial_analysis.py:437-465 (pad): 1-based inclusive source / destination windows of every box."""
    tw = (b[:, 2] - b[:, 0] + 1).astype(np.int32)
    th = (b[:, 3] - b[:, 1] + 1).astype(np.int32)
    dx, dy = np.ones_like(tw), np.ones_like(th)
    edx, edy = tw.copy(), th.copy()
    x, y, ex, ey = (b[:, i].astype(np.int32) for i in range(4))
    m = ex > w
    edx[m] = -ex[m] + w + tw[m]
    ex[m] = w
    m = ey > h
    edy[m] = -ey[m] + h + th[m]
    ey[m] = h
    m = x < 1
    dx[m] = 2 - x[m]
    x[m] = 1
    m = y < 1
    dy[m] = 2 - y[m]
    y[m] = 1
    return dy, edy, dx, edx, y, ey, x, ex, tw, th


def crops(img, b, size):
    """facial_analysis.py:537-546 / 564-573: zero-padded box crops resized with cv2.INTER_AREA, normalised, in the
    networks' transposed orientation [n, size(x), size(y), 3]."""
    h, w = img.shape[:2]
    dy, edy, dx, edx, y, ey, x, ex, tw, th = crop_windows(b.copy(), w, h)
    out = np.zeros((b.shape[0], size, size, 3))
    for k in range(b.shape[0]):
        tmp = np.zeros((int(th[k]), int(tw[k]), 3))
        tmp[dy[k] - 1:edy[k], dx[k] - 1:edx[k], :] = img[y[k] - 1:ey[k], x[k] - 1:ex[k], :]
        out[k] = cv2.resize(tmp, (size, size), interpolation=cv2.INTER_AREA)
    out = (out - 127.5) * 0.0078125
    return np.transpose(out, (0, 2, 1, 3))


def detect_faces(nets, img, minsize=32):
    """facial_analysis.py:478-604.  img: RGB uint8 [H, W, 3] -> (boxes [n, 5] = x1, y1, x2, y2, score; points [10, n])."""
    h, w = img.shape[:2]
    total = np.empty((0, 9))
    for scale in pyramid_scales(h, w, minsize):
        hs, ws = int(np.ceil(h * scale)), int(np.ceil(w * scale))
        im = (cv2.resize(img, (ws, hs), interpolation=cv2.INTER_AREA) - 127.5) * 0.0078125
        reg, prob = nets.pnet(np.transpose(im[None], (0, 2, 1, 3)))
        reg, prob = np.transpose(reg, (0, 2, 1, 3)), np.transpose(prob, (0, 2, 1, 3))
        boxes = boxes_from_heatmap(prob[0, :, :, 1].copy(), reg[0].copy(), scale, THRESHOLDS[0])
        pick = nms(boxes.copy(), 0.5, "Union")
        if boxes.size > 0 and pick.size > 0:
            total = np.append(total, boxes[pick], axis=0)
    points = np.array([])
    if total.shape[0] > 0:
        total = total[nms(total.copy(), 0.7, "Union")]
        rw, rh = total[:, 2] - total[:, 0], total[:, 3] - total[:, 1]
        total = np.stack([total[:, 0] + total[:, 5] * rw, total[:, 1] + total[:, 6] * rh, total[:, 2] + total[:, 7] * rw,
                          total[:, 3] + total[:, 8] * rh, total[:, 4]], axis=1)
        total = square(total.copy())
        total[:, 0:4] = np.fix(total[:, 0:4]).astype(np.int32)
    if total.shape[0] > 0:
        reg, prob = nets.rnet(crops(img, total, 24))
        score = prob[:, 1]
        ok = np.where(score > THRESHOLDS[1])[0]
        total = np.hstack([total[ok, 0:4].copy(), score[ok, None]])
        reg = reg[ok]
        if total.shape[0] > 0:
            pick = nms(total, 0.7, "Union")
            total = square(regress(total[pick].copy(), reg[pick]))
    if total.shape[0] > 0:
        total = np.fix(total).astype(np.int32)
        reg, pts, prob = nets.onet(crops(img, total, 48))
        score = prob[:, 1]
        ok = np.where(score > THRESHOLDS[2])[0]
        points = pts[ok].T.copy()
        total = np.hstack([total[ok, 0:4].copy(), score[ok, None]])
        reg = reg[ok]
        bw, bh = total[:, 2] - total[:, 0] + 1, total[:, 3] - total[:, 1] + 1
        points[0:5, :] = bw[None, :] * points[0:5, :] + total[:, 0][None, :] - 1
        points[5:10, :] = bh[None, :] * points[5:10, :] + total[:, 1][None, :] - 1
        if total.shape[0] > 0:
            total = regress(total.copy(), reg)
            pick = nms(total.copy(), 0.7, "Min")
            total, points = total[pick], points[:, pick]
    return total, points
