"""Kernel-level parity (through the C ABI single-operator entry points) against plain torch fp32 references."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from hse_facerec_tf_b200 import _lib
from hse_facerec_tf_b200._lib import check, lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TDT = {0: torch.float32, 1: torch.float32, 2: torch.bfloat16}


def _stream():
    return int(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else t.data_ptr()


def tf32_round(t):
    """round-to-nearest-even to 10 mantissa bits"""
    u = t.contiguous().view(torch.int32)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & ~0x1FFF
    return u.view(torch.float32)


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    (128, 64, 32), (128, 64, 64), (300, 128, 256), (1000, 256, 1024), (4096, 512, 512), (12544, 64, 32),
    (2304, 1024, 512), (130, 1024, 1024), (77, 128, 128), (50176, 128, 64),
    # large enough for the CTA-pair (cta_group::2) tiles: odd m-block count with a ragged tail, 256- and 128-wide pair tiles
    (20000, 256, 512), (33000, 512, 192), (9600, 1024, 2048), (40100, 384, 64),
]


@pytest.mark.parametrize("prec", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_bias_act(prec, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    res = torch.randn(M, N, generator=g).to(DEV)
    dt = TDT[prec]
    if prec == 1:
        a, b = tf32_round(a), tf32_round(b)
    a_t, b_t, res_t = a.to(dt).contiguous(), b.to(dt).contiguous(), res.to(dt).contiguous()
    for act, use_res in ((0, False), (2, False), (1, True)):
        y = torch.full((M, N), float("nan"), dtype=dt, device=DEV)
        check(lib.hfr_op_gemm_bias_act(a_t.data_ptr(), b_t.data_ptr(), bias.data_ptr(), _ptr(res_t) if use_res else None,
                                       y.data_ptr(), M, N, K, act, prec, 0, _stream()))
        torch.cuda.synchronize()
        ref = a_t.double() @ b_t.double().t() + bias.double()
        if use_res:
            ref = ref + res_t.double()
        if act == 1:
            ref = torch.relu(ref)
        if act == 2:
            ref = torch.clamp(ref, 0, 6)
        got = y.double()
        assert torch.isfinite(got).all(), "kernel left unwritten / non-finite outputs"
        # operands are exactly representable, accumulation is fp32: only the output rounding differs
        tol = 2e-2 if prec == 2 else 2e-4
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item() + 1.0
        assert err <= tol * scale, f"M={M} N={N} K={K} prec={prec} act={act}: max err {err} (scale {scale})"


# (M, K0, K1, N1, N2): K0 = columns of the concatenated operand (0: none), N2 = 0: single K-concatenated GEMM
PAIR_CASES = [
    (802816 // 16, 0, 64, 256, 64), (12544, 64, 64, 256, 64), (9408, 0, 128, 512, 128), (20000, 0, 256, 1024, 256),
    (130, 0, 64, 128, 64), (128 * 149 + 5, 0, 128, 256, 128), (4000, 64, 128, 512, 256), (3000, 128, 128, 384, 64),
    (5000, 64, 64, 256, 0), (7777, 128, 256, 512, 0), (20000, 256, 512, 1024, 0), (300, 512, 1024, 2048, 0),
]


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("M,K0,K1,N1,N2", PAIR_CASES)
def test_gemm_pair_and_k_concatenation(prec, M, K0, K1, N1, N2, monkeypatch):
    """gemm_pair_kernel (two dependent 1x1 convolutions in one launch, the second fed from shared memory) and the
    K-concatenated A operand, against fp64 matmuls of the same (exactly representable) operands: with and without the
    shortcut, ragged M, fewer and more units than SMs, every (N2, staging-buffer) instantiation the launcher picks."""
    monkeypatch.setenv("HFR_SEAM", "1")
    es = 2 if prec == 2 else 4
    if N2 and (K0 + K1) * es // 128 > 4:
        pytest.skip("A rows of a unit exceed the resident buffer: not eligible in this precision (the launcher refuses)")
    g = torch.Generator(device="cpu").manual_seed(M + 3 * K1 + 5 * N1 + 7 * N2 + K0)
    dt = TDT[prec]
    rnd = (lambda t: tf32_round(t)) if prec == 1 else (lambda t: t)
    a = rnd(torch.randn(M, K1, generator=g).to(DEV)).to(dt).contiguous()
    a0 = rnd(torch.randn(M, K0, generator=g).to(DEV)).to(dt).contiguous() if K0 else None
    w1 = rnd((torch.randn(N1, K0 + K1, generator=g) / (K0 + K1) ** 0.5).to(DEV)).to(dt).contiguous()
    b1 = torch.randn(N1, generator=g).to(DEV)
    res = rnd(torch.randn(M, N1, generator=g).to(DEV)).to(dt).contiguous()
    w2 = rnd((torch.randn(max(N2, 1), N1, generator=g) / N1 ** 0.5).to(DEV)).to(dt).contiguous()
    b2 = torch.randn(max(N2, 1), generator=g).to(DEV)
    for use_res in (True, False):
        y = torch.full((M, N1), float("nan"), dtype=dt, device=DEV)
        z = torch.full((M, max(N2, 1)), float("nan"), dtype=dt, device=DEV)
        check(lib.hfr_op_gemm_pair(_ptr(a0), K0, a.data_ptr(), K1, w1.data_ptr(), b1.data_ptr(), _ptr(res) if use_res else None,
                                   y.data_ptr(), M, N1, 1, w2.data_ptr() if N2 else None, b2.data_ptr() if N2 else None,
                                   z.data_ptr() if N2 else None, N2, 1, prec, 0, _stream()))
        torch.cuda.synchronize()
        acat = a.double() if a0 is None else torch.cat([a0.double(), a.double()], dim=1)
        ref_y = acat @ w1.double().t() + b1.double()
        if use_res:
            ref_y = ref_y + res.double()
        ref_y = torch.relu(ref_y)
        tol = 2e-2 if prec == 2 else 2e-3
        assert torch.isfinite(y.double()).all(), "kernel left unwritten / non-finite outputs"
        err = (y.double() - ref_y).abs().max().item()
        assert err <= tol * (ref_y.abs().max().item() + 1.0), f"Y: max err {err}"
        if N2:
            # the second GEMM consumes the ROUNDED y the kernel stored: compare against that, tightly
            ref_z = torch.relu(y.double() @ w2.double().t() + b2.double())
            assert torch.isfinite(z.double()).all()
            errz = (z.double() - ref_z).abs().max().item()
            assert errz <= tol * (ref_z.abs().max().item() + 1.0), f"Z: max err {errz}"


# ------------------------------------------------------------------------------------------------ depthwise
DW_CASES = [  # (B, H, W, C, stride)
    (2, 16, 16, 32, 1), (2, 16, 16, 64, 2), (3, 12, 12, 128, 1), (2, 14, 14, 512, 2), (2, 7, 7, 1024, 1),
    (1, 6, 6, 256, 1), (2, 96, 96, 32, 1), (1, 112, 112, 64, 2), (2, 13, 9, 64, 2), (2, 24, 24, 256, 1),
]


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("B,H,W,C,stride", DW_CASES)
def test_dwconv3x3(prec, B, H, W, C, stride):
    g = torch.Generator(device="cpu").manual_seed(H * 31 + C + stride)
    dt = TDT[prec]
    x = torch.randn(B, H, W, C, generator=g).to(DEV).to(dt).contiguous()
    w = torch.randn(9, C, generator=g).to(DEV)
    bias = torch.randn(C, generator=g).to(DEV)
    ho, wo = -(-H // stride), -(-W // stride)
    pt = max((ho - 1) * stride + 3 - H, 0) // 2
    pl = max((wo - 1) * stride + 3 - W, 0) // 2
    pb = max((ho - 1) * stride + 3 - H, 0) - pt
    pr = max((wo - 1) * stride + 3 - W, 0) - pl
    y = torch.full((B, ho, wo, C), float("nan"), dtype=dt, device=DEV)
    check(lib.hfr_op_dwconv3x3(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, C, stride, pt, pl, ho,
                               wo, 2, prec, 0, _stream()))
    torch.cuda.synchronize()
    xp = F.pad(x.float().permute(0, 3, 1, 2), (pl, pr, pt, pb))
    ref = F.conv2d(xp, w.view(3, 3, C).permute(2, 0, 1).unsqueeze(1), bias, stride=stride, groups=C)
    ref = torch.clamp(ref, 0, 6).permute(0, 2, 3, 1)
    got = y.float()
    assert torch.isfinite(got).all()
    tol = 4e-2 if prec == 2 else 1e-4
    assert (got - ref).abs().max().item() <= tol


# ------------------------------------------------------------------------------------------------ stem
@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("case", [(2, 32, 32, 3, 2, 32, "u8"), (2, 30, 30, 3, 2, 32, "f32"), (1, 38, 38, 7, 2, 64, "u8")])
def test_stem_conv(prec, case):
    B, H, W, k, stride, cout, kind = case
    g = torch.Generator(device="cpu").manual_seed(H + k)
    dt = TDT[prec]
    w = (torch.randn(k, k, 3, cout, generator=g) * 0.03).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    ho, wo = -(-H // stride), -(-W // stride)
    tot_h = max((ho - 1) * stride + k - H, 0)
    tot_w = max((wo - 1) * stride + k - W, 0)
    pt, pl = tot_h // 2, tot_w // 2
    if kind == "u8":
        x = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).to(DEV)
        flags = _lib.FLAG_BGR | _lib.FLAG_MEAN_IMAGENET
        xf = x.float().flip(-1) - torch.tensor([103.939, 116.779, 123.68], device=DEV)
        in_dt = 1
    else:
        x = (torch.randn(B, H, W, 3, generator=g) * 60).to(DEV)
        flags, xf, in_dt = 0, x, 0
    y = torch.full((B, ho, wo, cout), float("nan"), dtype=dt, device=DEV)
    check(lib.hfr_op_stem_conv(x.data_ptr(), in_dt, w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, k, k, stride,
                               pt, pl, ho, wo, cout, flags, 2, prec, 0, _stream()))
    torch.cuda.synchronize()
    xp = F.pad(xf.permute(0, 3, 1, 2), (pl, tot_w - pl, pt, tot_h - pt))
    ref = torch.clamp(F.conv2d(xp, w.permute(3, 2, 0, 1), bias, stride=stride), 0, 6).permute(0, 2, 3, 1)
    got = y.float()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= (4e-2 if prec == 2 else 2e-3)


@pytest.mark.parametrize("case", [(2, 32, 32, 3, 32, 0), (3, 64, 48, 7, 64, 3), (2, 192, 192, 3, 32, 0), (1, 224, 224, 7, 64, 3),
                                  (2, 20, 20, 5, 64, 2), (2, 16, 16, 3, 64, 1)])
@pytest.mark.parametrize("flags", [_lib.FLAG_BGR | _lib.FLAG_MEAN_IMAGENET, _lib.FLAG_BGR | _lib.FLAG_MEAN_VGGFACE2,
                                   _lib.FLAG_SCALE_PM1])
def test_stem_conv_tensor_core(case, flags):
    """space-to-depth + STEM16 implicit GEMM vs the fp32 convolution of the pre-processed image.  The image enters
    exactly (uint8 in bf16), the mean term is exact; only the weights are rounded to bf16."""
    B, H, W, k, cout, pad = case
    g = torch.Generator(device="cpu").manual_seed(H + k + cout)
    w = (torch.randn(k, k, 3, cout, generator=g) * (0.5 / k)).float()
    if flags & _lib.FLAG_SCALE_PM1:
        w = w * 60
    else:
        w = w / 60
    bias = torch.randn(cout, generator=g).to(DEV)
    x = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).to(DEV)
    ho = (H + 2 * pad - k) // 2 + 1 if pad else -(-H // 2)
    wo = (W + 2 * pad - k) // 2 + 1 if pad else -(-W // 2)
    if pad:
        pt = pl = pad
        pb, pr = (ho - 1) * 2 + k - H - pt, (wo - 1) * 2 + k - W - pl
    else:   # TF SAME
        th, tw = max((ho - 1) * 2 + k - H, 0), max((wo - 1) * 2 + k - W, 0)
        pt, pl, pb, pr = th // 2, tw // 2, th - th // 2, tw - tw // 2
    y = torch.full((B, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    wc = w.contiguous()
    check(lib.hfr_op_stem_conv_tc(x.data_ptr(), wc.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, k, k, pt, pl, ho, wo,
                                  cout, flags, 2, 0, _stream()))
    # reference = what the kernel is specified to compute: sum(u8 * bf16(scale * W)) - sum_valid(mean * W) + bias
    if flags & _lib.FLAG_SCALE_PM1:
        xs, scale, mean = x.double(), 1.0 / 127.5, [1.0, 1.0, 1.0]
    else:
        xs, scale = x.double().flip(-1), 1.0
        mean = [103.939, 116.779, 123.68] if flags & _lib.FLAG_MEAN_IMAGENET else [91.4953, 103.8827, 131.0912]
    wd = w.to(DEV)
    wq = (wd * scale).bfloat16().double()
    pads = (pl, max(pr, 0), pt, max(pb, 0))
    xp = F.pad(xs.permute(0, 3, 1, 2), pads)
    mp = F.pad(torch.ones_like(xs).permute(0, 3, 1, 2) * torch.tensor(mean, device=DEV, dtype=torch.float64).view(1, 3, 1, 1), pads)
    ref = F.conv2d(xp, wq.permute(3, 2, 0, 1), bias.double(), stride=2) - F.conv2d(mp, wd.double().permute(3, 2, 0, 1), None, stride=2)
    ref = torch.clamp(ref, 0, 6).permute(0, 2, 3, 1)[:, :ho, :wo]
    got = y.double()
    assert torch.isfinite(got).all()
    # weights are compared at their bf16 values, so what is left is the bf16 rounding of the output (<= 2^-8 * 6)
    # and of the scaled weights in the x/127.5-1 mode
    assert (got - ref).abs().max().item() <= 0.05


# ------------------------------------------------------------------------------------------------ KxK conv / pool
CONV_CASES = [  # (B, H, W, cin, cout, k, stride)
    (2, 56, 56, 64, 64, 3, 1), (3, 28, 28, 128, 128, 3, 1), (2, 14, 14, 256, 256, 3, 1), (5, 7, 7, 512, 512, 3, 1),
    (1, 9, 11, 64, 128, 3, 1), (2, 16, 16, 64, 64, 3, 2),
    (128, 14, 14, 256, 256, 3, 1), (32, 28, 28, 128, 128, 3, 1), (40, 28, 28, 64, 256, 3, 2),  # CTA-pair tiles
    # strided 1x1 convolutions (the gather the plan states as a subsample layer, fetched through the im2col map):
    # pair tiles with K = 1024, plain tiles, odd spatial size
    (64, 28, 28, 1024, 512, 1, 2), (2, 56, 56, 256, 128, 1, 2), (3, 27, 27, 128, 64, 1, 2),
]


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("B,H,W,cin,cout,k,stride", CONV_CASES)
def test_conv2d_implicit_gemm(prec, B, H, W, cin, cout, k, stride):
    g = torch.Generator(device="cpu").manual_seed(H * 13 + cin)
    dt = TDT[prec]
    x = torch.randn(B, H, W, cin, generator=g).to(DEV)
    w = (torch.randn(cout, k, k, cin, generator=g) / (k * k * cin) ** 0.5).to(DEV)
    if prec == 1:
        x, w = tf32_round(x), tf32_round(w)
    x, w = x.to(dt).contiguous(), w.to(dt).contiguous()
    bias = torch.randn(cout, generator=g).to(DEV)
    ho, wo = -(-H // stride), -(-W // stride)
    th, tw = max((ho - 1) * stride + k - H, 0), max((wo - 1) * stride + k - W, 0)
    res = torch.randn(B, ho, wo, cout, generator=g).to(DEV).to(dt).contiguous()
    y = torch.full((B, ho, wo, cout), float("nan"), dtype=dt, device=DEV)
    check(lib.hfr_op_conv2d(x.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr(), y.data_ptr(), B, H, W, cin, k, k,
                            stride, th // 2, tw // 2, ho, wo, cout, 1, prec, 0, _stream()))
    torch.cuda.synchronize()
    xp = F.pad(x.float().permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2))
    ref = F.conv2d(xp.double(), w.double().permute(0, 3, 1, 2), bias.double(), stride=stride).permute(0, 2, 3, 1)
    ref = torch.relu(ref + res.double())
    got = y.double()
    assert torch.isfinite(got).all()
    tol = 2e-2 if prec == 2 else 2e-4
    assert (got - ref).abs().max().item() <= tol * (ref.abs().max().item() + 1.0)


@pytest.mark.parametrize("B,H,W,cin,cout,k", [(2, 56, 56, 64, 64, 3), (3, 28, 20, 64, 64, 3), (1, 9, 11, 32, 32, 3),
                                             (2, 16, 16, 16, 64, 4), (1, 17, 23, 64, 64, 1)])
def test_conv2d_window(B, H, W, cin, cout, k):
    """smem-window implicit GEMM (no-swizzle UMMA descriptors over the TMA-loaded halo window) vs torch."""
    g = torch.Generator(device="cpu").manual_seed(H * 5 + cin + k)
    x = torch.randn(B, H, W, cin, generator=g).to(DEV).bfloat16().contiguous()
    w = (torch.randn(cout, k, k, cin, generator=g) / (k * k * cin) ** 0.5).bfloat16().float().contiguous()
    bias = torch.randn(cout, generator=g).to(DEV)
    pad = (k - 1) // 2
    ho, wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    y = torch.full((B, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    check(lib.hfr_op_conv2d_window(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, cin, k, k, pad, pad,
                                   ho, wo, cout, 1, 0, _stream()))
    xp = F.pad(x.double().permute(0, 3, 1, 2), (pad, pad, pad, pad))
    ref = torch.relu(F.conv2d(xp, w.to(DEV).double().permute(0, 3, 1, 2), bias.double())).permute(0, 2, 3, 1)
    got = y.double()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 2e-2 * (ref.abs().max().item() + 1.0)


@pytest.mark.parametrize("prec", [1, 2])
@pytest.mark.parametrize("explicit_zero", [0, 1])
@pytest.mark.parametrize("shape", [(3, 112, 112, 64), (2, 13, 13, 64), (2, 9, 14, 32), (1, 5, 3, 128)])
def test_maxpool(prec, explicit_zero, shape):
    """3x3 / stride 2 SAME pooling (pad_before = total // 2): even sizes pad bottom/right only, odd sizes pad both
    sides and give an odd output width (the paired-output bf16 kernel's ragged last column)."""
    dt = TDT[prec]
    B, H, W, C = shape
    ho, wo = -(-H // 2), -(-W // 2)
    th, tw = max((ho - 1) * 2 + 3 - H, 0), max((wo - 1) * 2 + 3 - W, 0)
    x = (torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(4)) - 0.5).to(DEV).to(dt).contiguous()
    y = torch.full((B, ho, wo, C), float("nan"), dtype=dt, device=DEV)
    check(lib.hfr_op_maxpool(x.data_ptr(), y.data_ptr(), B, H, W, C, 3, 2, th // 2, tw // 2, ho, wo, explicit_zero, prec, 0,
                             _stream()))
    torch.cuda.synchronize()
    xp = F.pad(x.float().permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2),
               value=0.0 if explicit_zero else -float("inf"))
    ref = F.max_pool2d(xp, 3, 2).permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref)


# ------------------------------------------------------------------------------------------------ small ops
def test_l2_normalize_matches_sklearn():
    from sklearn import preprocessing
    x = np.random.RandomState(0).randn(1000, 1024).astype(np.float32)
    x[5] = 0
    from hse_facerec_tf_b200 import normalize
    got = normalize(x)
    ref = preprocessing.normalize(x, norm="l2")
    assert got.dtype == np.float32 and (got[5] == 0).all()
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-8)


def test_age_post_matches_reference_rule():
    from oracle.tfnet import age_from_probs
    rs = np.random.RandomState(1)
    p = rs.dirichlet(np.ones(100) * 0.3, size=257).astype(np.float32)
    p[0, :] = 0
    p[0, [10, 20, 30]] = [0.3, 0.3, 0.1]   # exact tie -> higher index first
    t = torch.from_numpy(p).to(DEV)
    age = torch.empty(p.shape[0], device=DEV)
    check(lib.hfr_age_gender_post(t.data_ptr(), p.shape[0], 100, age.data_ptr(), 0, _stream()))
    ref = np.array([age_from_probs(r)[0] for r in p])
    np.testing.assert_allclose(age.cpu().numpy(), ref, rtol=1e-5)
