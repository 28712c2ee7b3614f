"""ORACLE (test infrastructure, not product code): the uint8 INTER_LINEAR arithmetic of cv2.resize, restated in numpy.

The reference resizes every face crop with `cv2.resize(img, (w, h))` (age_gender_identity/facial_analysis.py:95).  OpenCV
(opencv-python 4.13 in this image, an un-vendored dependency of the reference) IS importable here, so the tests pin both
this restatement and the CUDA kernel directly against cv2.resize itself (bit-exact, tests/test_staging_*.py).
"""
import numpy as np


def _coefs(src, dst, is_x):
    scale = src / dst
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s).astype(np.float32)
    if is_x:                                   # x: weights forced to (1, 0) outside [0, src-1); y: only indices clamp
        lo, hi = s < 0, s >= src - 1
        f[lo], s[lo] = 0, 0
        f[hi], s[hi] = 0, src - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), a0, a1


def resize_linear_u8(img: np.ndarray, ow: int, oh: int) -> np.ndarray:
    """img: uint8 [H,W,C] -> uint8 [oh,ow,C], identical to cv2.resize(img, (ow, oh), interpolation=cv2.INTER_LINEAR)."""
    H, W, _ = img.shape
    sx, sx1, a0, a1 = _coefs(W, ow, True)
    sy, sy1, b0, b1 = _coefs(H, oh, False)
    I = img.astype(np.int64)
    r0 = I[sy][:, sx] * a0[None, :, None] + I[sy][:, sx1] * a1[None, :, None]
    r1 = I[sy1][:, sx] * a0[None, :, None] + I[sy1][:, sx1] * a1[None, :, None]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
