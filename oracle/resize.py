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


# ---------------------------------------------------------------------------------------------------------------------
# scipy.misc.imresize(img, size, interp='bilinear') (facerec_test.py:84,93) = PIL.Image.resize(size, BILINEAR) on the
# uint8 array.  scipy.misc.imresize no longer exists (SciPy >= 1.3) but Pillow - the dependency that did the work - is
# importable here (12.2), so tests pin this restatement and the CUDA kernel against PIL itself, bit for bit.
# Restated from Pillow's published algorithm (libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
# ImagingResampleHorizontal_8bpc / Vertical_8bpc): a separable triangle filter whose support is widened by the scale
# factor when shrinking (antialiasing), coefficients normalised in double precision and quantised to 22-bit fixed point,
# horizontal pass first with its result rounded to uint8, then the vertical pass.
PIL_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """-> (xmin[out], count[out], int coefficients [out][ksize]) exactly as Pillow computes them for box (0, in_size)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int64)
    cnt = np.zeros(out_size, np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        x0 = int(center - support + 0.5)          # C (int) cast truncates toward zero
        x0 = max(x0, 0)
        x1 = int(center + support + 0.5)
        x1 = min(x1, in_size)
        n = x1 - x0
        w = np.empty(n, np.float64)
        for x in range(n):
            t = (x + x0 - center + 0.5) * ss
            t = -t if t < 0 else t
            w[x] = 1.0 - t if t < 1.0 else 0.0
        ww = 0.0
        for x in range(n):
            ww += w[x]
        for x in range(n):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PIL_PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PIL_PRECISION_BITS))
        xmin[xx], cnt[xx] = x0, n
    return xmin, cnt, kk


def _pil_pass(img, xmin, cnt, kk, axis):
    """one resample pass along `axis` (1 = horizontal, 0 = vertical) of a uint8 [H,W,C] image"""
    src = np.moveaxis(img.astype(np.int64), axis, 0)            # [in, other, C]
    out = np.empty((len(xmin),) + src.shape[1:], np.int64)
    for o in range(len(xmin)):
        acc = np.full(src.shape[1:], 1 << (PIL_PRECISION_BITS - 1), np.int64)
        for j in range(cnt[o]):
            acc += src[xmin[o] + j] * kk[o, j]
        out[o] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def pil_resize_bilinear_u8(img: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """img: uint8 [H,W,C] -> uint8 [oh,ow,C], identical to np.asarray(Image.fromarray(img).resize((ow, oh), BILINEAR))."""
    H, W, _ = img.shape
    if W != ow:
        img = _pil_pass(img, *pil_bilinear_coeffs(W, ow), axis=1)
    if H != oh:
        img = _pil_pass(img, *pil_bilinear_coeffs(H, oh), axis=0)
    return img
