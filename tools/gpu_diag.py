"""GPU diagnostics for the hand-written kernels: runs each kernel family on tiny problems and prints WHERE results
differ (row/column structure), so one gpurun round-trip is enough to localise a descriptor / layout mistake.
Usage: python tools/gpu_diag.py [gemm|dw|stem|knn|model ...]"""
import ctypes as C
import os
import sys
import traceback

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hse_facerec_tf_b200._lib import check, lib  # noqa: E402

DEV = "cuda:0"
TDT = {0: torch.float32, 1: torch.float32, 2: torch.bfloat16}


def err_map(got, ref, bm=32, bn=32):
    e = (got - ref).abs()
    M, N = e.shape
    Mp, Np = -(-M // bm) * bm, -(-N // bn) * bn
    ep = torch.zeros(Mp, Np, device=e.device, dtype=e.dtype)
    ep[:M, :N] = e
    blk = ep.view(Mp // bm, bm, Np // bn, bn).amax(dim=(1, 3))
    return blk


def diag_gemm():
    for prec in (0, 2, 1):
        for (M, N, K) in [(128, 64, 64), (128, 64, 32), (256, 128, 128), (300, 256, 512), (1000, 512, 96)]:
            g = torch.Generator().manual_seed(1)
            dt = TDT[prec]
            # small-integer operands: every product/sum is exact in bf16/tf32/fp32 -> any error is structural
            a = torch.randint(-3, 4, (M, K), generator=g).float().to(DEV).to(dt)
            b = torch.randint(-3, 4, (N, K), generator=g).float().to(DEV).to(dt)
            y = torch.full((M, N), float("nan"), dtype=dt, device=DEV)
            rc = lib.hfr_op_gemm_bias_act(a.data_ptr(), b.data_ptr(), None, None, y.data_ptr(), M, N, K, 0, prec, 0, None)
            torch.cuda.synchronize()
            if rc:
                print(f"gemm prec={prec} {M}x{N}x{K}: rc={rc} {lib.hfr_last_error().decode()}")
                continue
            ref = a.float() @ b.float().t()
            got = y.float()
            nan = torch.isnan(got).sum().item()
            bad = (torch.nan_to_num(got, nan=1e9) != ref)
            print(f"gemm prec={prec} {M}x{N}x{K}: nan={nan} mismatches={bad.sum().item()}/{M * N}")
            if bad.any():
                rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
                cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
                print("   first bad rows", rows, "cols", cols)
                print("   got[0,:8]", got[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())
                print("   block err map (32x32 blocks):\n", err_map(torch.nan_to_num(got, nan=1e9), ref).cpu().numpy().round(1))


def diag_dw():
    for prec in (1, 2):
        for (B, H, W, Cc, s) in [(1, 8, 8, 32, 1), (1, 16, 16, 64, 1), (2, 16, 16, 64, 2), (1, 12, 12, 128, 1)]:
            dt = TDT[prec]
            g = torch.Generator().manual_seed(2)
            x = torch.randint(-2, 3, (B, H, W, Cc), generator=g).float().to(DEV).to(dt)
            w = torch.randint(-2, 3, (9, Cc), generator=g).float().to(DEV)
            bias = torch.zeros(Cc, device=DEV)
            ho, wo = -(-H // s), -(-W // s)
            th, tw = max((ho - 1) * s + 3 - H, 0), max((wo - 1) * s + 3 - W, 0)
            y = torch.full((B, ho, wo, Cc), float("nan"), dtype=dt, device=DEV)
            rc = lib.hfr_op_dwconv3x3(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, Cc, s, th // 2,
                                      tw // 2, ho, wo, 0, prec, 0, None)
            torch.cuda.synchronize()
            if rc:
                print(f"dw prec={prec} {B}x{H}x{W}x{Cc} s{s}: rc={rc} {lib.hfr_last_error().decode()}")
                continue
            xp = F.pad(x.float().permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2))
            ref = F.conv2d(xp, w.view(3, 3, Cc).permute(2, 0, 1).unsqueeze(1), None, stride=s, groups=Cc).permute(0, 2, 3, 1)
            got = torch.nan_to_num(y.float(), nan=1e9)
            bad = got != ref
            print(f"dw prec={prec} {B}x{H}x{W}x{Cc} s{s}: mismatches={bad.sum().item()}/{bad.numel()}")
            if bad.any():
                idx = bad.nonzero()[:8].tolist()
                print("   first bad (b,y,x,c):", idx)
                print("   bad per output row:", bad.sum(dim=(0, 2, 3)).tolist(), " per col:", bad.sum(dim=(0, 1, 3)).tolist())


def diag_conv():
    for prec in (2, 1):
        for (B, H, W, cin, cout, k, s) in [(1, 8, 8, 64, 64, 3, 1), (2, 12, 12, 64, 64, 3, 1), (3, 7, 7, 128, 64, 3, 1),
                                           (2, 14, 14, 64, 128, 3, 1), (1, 16, 16, 64, 64, 3, 2), (2, 56, 56, 64, 64, 3, 1)]:
            dt = TDT[prec]
            g = torch.Generator().manual_seed(3)
            x = torch.randint(-2, 3, (B, H, W, cin), generator=g).float().to(DEV).to(dt)
            w = torch.randint(-1, 2, (cout, k, k, cin), generator=g).float().to(DEV).to(dt)   # [cout][kh*kw][cin]
            ho, wo = -(-H // s), -(-W // s)
            th, tw = max((ho - 1) * s + k - H, 0), max((wo - 1) * s + k - W, 0)
            y = torch.full((B, ho, wo, cout), float("nan"), dtype=dt, device=DEV)
            rc = lib.hfr_op_conv2d(x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), B, H, W, cin, k, k, s, th // 2,
                                   tw // 2, ho, wo, cout, 0, prec, 0, None)
            torch.cuda.synchronize()
            if rc:
                print(f"conv prec={prec} {B}x{H}x{W}x{cin}->{cout} k{k} s{s}: rc={rc} {lib.hfr_last_error().decode()}")
                continue
            xp = F.pad(x.float().permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2))
            ref = F.conv2d(xp, w.float().permute(0, 3, 1, 2), None, stride=s).permute(0, 2, 3, 1)
            got = torch.nan_to_num(y.float(), nan=1e9)
            bad = got != ref
            print(f"conv prec={prec} {B}x{H}x{W}x{cin}->{cout} k{k} s{s}: mismatches={bad.sum().item()}/{bad.numel()}")
            if bad.any():
                print("   first bad (b,y,x,c):", bad.nonzero()[:6].tolist())
                print("   bad per image:", bad.sum(dim=(1, 2, 3)).tolist(), " per out row (img0):", bad[0].sum(dim=(1, 2)).tolist(),
                      " per out col (img0):", bad[0].sum(dim=(0, 2)).tolist())
                print("   got[0,0,:4,0]", got[0, 0, :4, 0].tolist(), "ref", ref[0, 0, :4, 0].tolist(),
                      " got[0,1,:4,0]", got[0, 1, :4, 0].tolist(), "ref", ref[0, 1, :4, 0].tolist())


def diag_stem():
    for (B, H, W, k, cout, pad) in [(1, 16, 16, 3, 32, 0), (1, 16, 16, 3, 64, 0), (2, 32, 32, 7, 64, 3), (1, 224, 224, 7, 64, 3)]:
        g = torch.Generator().manual_seed(5)
        x = torch.randint(0, 4, (B, H, W, 3), generator=g, dtype=torch.uint8).to(DEV)
        w = torch.randint(-1, 2, (k, k, 3, cout), generator=g).float().contiguous()
        ho = (H + 2 * pad - k) // 2 + 1 if pad else H // 2
        wo = (W + 2 * pad - k) // 2 + 1 if pad else W // 2
        if pad:
            pt = pl = pad
            pb, pr = (ho - 1) * 2 + k - H - pt, (wo - 1) * 2 + k - W - pl
        else:
            th, tw = max((ho - 1) * 2 + k - H, 0), max((wo - 1) * 2 + k - W, 0)
            pt, pl, pb, pr = th // 2, tw // 2, th - th // 2, tw - tw // 2
        y = torch.full((B, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
        rc = lib.hfr_op_stem_conv_tc(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), B, H, W, k, k, pt, pl, ho, wo, cout, 0, 0, 0, None)
        torch.cuda.synchronize()
        if rc:
            print(f"stem {B}x{H}x{W} k{k} ->{cout}: rc={rc} {lib.hfr_last_error().decode()}")
            continue
        xp = F.pad(x.float().permute(0, 3, 1, 2), (pl, max(pr, 0), pt, max(pb, 0)))
        ref = F.conv2d(xp, w.to(DEV).permute(3, 2, 0, 1), None, stride=2).permute(0, 2, 3, 1)[:, :ho, :wo]
        got = torch.nan_to_num(y.float(), nan=1e9)
        bad = (got - ref).abs() > 0.01 * ref.abs() + 0.01
        print(f"stem {B}x{H}x{W} k{k} ->{cout}: mismatches={bad.sum().item()}/{bad.numel()}")
        if bad.any():
            print("   first bad (b,y,x,c):", bad.nonzero()[:6].tolist())
            print("   bad per out row (img0):", bad[0].sum(dim=(1, 2)).tolist(), " per col:", bad[0].sum(dim=(0, 2)).tolist(),
                  " per channel:", bad[0].sum(dim=(0, 1)).tolist())
            print("   got[0,0,:4,0]", got[0, 0, :4, 0].tolist(), "ref", ref[0, 0, :4, 0].tolist())
            print("   got[0,1,:4,1]", got[0, 1, :4, 1].tolist(), "ref", ref[0, 1, :4, 1].tolist())


def diag_win():
    for (B, H, W, cin, cout, k) in [(1, 16, 8, 16, 64, 1), (1, 16, 8, 16, 64, 3), (1, 16, 8, 64, 64, 3), (2, 56, 56, 64, 64, 3),
                                    (1, 20, 13, 32, 32, 3), (3, 12, 12, 16, 64, 4)]:
        g = torch.Generator().manual_seed(7)
        x = torch.randint(-2, 3, (B, H, W, cin), generator=g).float().to(DEV).bfloat16()
        w = torch.randint(-1, 2, (cout, k, k, cin), generator=g).float().contiguous()
        pad = (k - 1) // 2
        ho, wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
        y = torch.full((B, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
        rc = lib.hfr_op_conv2d_window(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), B, H, W, cin, k, k, pad, pad, ho, wo, cout, 0, 0, None)
        torch.cuda.synchronize()
        if rc:
            print(f"win {B}x{H}x{W}x{cin}->{cout} k{k}: rc={rc} {lib.hfr_last_error().decode()}")
            continue
        xp = F.pad(x.float().permute(0, 3, 1, 2), (pad, pad, pad, pad))
        ref = F.conv2d(xp, w.to(DEV).permute(0, 3, 1, 2), None).permute(0, 2, 3, 1)
        got = torch.nan_to_num(y.float(), nan=1e9)
        bad = (got - ref).abs() > 0.01 * ref.abs() + 0.01
        print(f"win {B}x{H}x{W}x{cin}->{cout} k{k}: mismatches={bad.sum().item()}/{bad.numel()}")
        if bad.any():
            print("   first bad (b,y,x,c):", bad.nonzero()[:6].tolist())
            print("   bad per out row (img0):", bad[0].sum(dim=(1, 2)).tolist(), " per col:", bad[0].sum(dim=(0, 2)).tolist())
            print("   bad per channel (img0):", bad[0].sum(dim=(0, 1)).tolist())
            print("   got[0,0,:8,0]", got[0, 0, :8, 0].tolist(), "\n   ref[0,0,:8,0]", ref[0, 0, :8, 0].tolist())
            print("   got[0,:8,0,0]", got[0, :8, 0, 0].tolist(), "\n   ref[0,:8,0,0]", ref[0, :8, 0, 0].tolist())


def diag_dwpw():
    for (B, H, W, cin, cout) in [(1, 16, 8, 64, 64), (2, 16, 16, 64, 128), (2, 12, 12, 128, 256), (3, 6, 6, 256, 512),
                                 (2, 24, 24, 128, 128), (5, 7, 7, 512, 1024)]:
        g = torch.Generator().manual_seed(9)
        x = torch.randint(-2, 3, (B, H, W, cin), generator=g).float().to(DEV).bfloat16()
        dw = torch.randint(-1, 2, (9, cin), generator=g).float().to(DEV)
        dwb = torch.randint(-1, 2, (cin,), generator=g).float().to(DEV)
        pw = torch.randint(-1, 2, (cout, cin), generator=g).float().to(DEV).bfloat16()
        y = torch.full((B, H, W, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
        rc = lib.hfr_op_dwpw(x.data_ptr(), dw.data_ptr(), dwb.data_ptr(), pw.data_ptr(), None, y.data_ptr(), B, H, W, cin, cout, 0, 0, 0, None)
        torch.cuda.synchronize()
        if rc:
            print(f"dwpw {B}x{H}x{W}x{cin}->{cout}: rc={rc} {lib.hfr_last_error().decode()}")
            continue
        xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))
        mid = F.conv2d(xp, dw.view(3, 3, cin).permute(2, 0, 1).unsqueeze(1), dwb, groups=cin)
        mid = mid.bfloat16().float()      # the depthwise output is rounded to bf16 when it becomes the A operand
        ref = torch.einsum("bchw,oc->bhwo", mid, pw.float())
        got = torch.nan_to_num(y.float(), nan=1e9)
        bad = (got - ref).abs() > 0.01 * ref.abs() + 0.01
        print(f"dwpw {B}x{H}x{W}x{cin}->{cout}: mismatches={bad.sum().item()}/{bad.numel()}")
        if bad.any():
            print("   first bad (b,y,x,c):", bad.nonzero()[:6].tolist())
            print("   bad per image:", bad.sum(dim=(1, 2, 3)).tolist(), "per row(img0):", bad[0].sum(dim=(1, 2)).tolist(),
                  " per col:", bad[0].sum(dim=(0, 2)).tolist())
            print("   got[0,0,:4,0]", got[0, 0, :4, 0].tolist(), "ref", ref[0, 0, :4, 0].tolist())


def diag_fuserace():
    """Is the fused dw+pw path deterministic / batch-invariant?  Per-layer comparison of the same images run at B=64 and B=70."""
    import hse_facerec_tf_b200 as hfr
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pb = os.path.join(root, "tests", "golden", "age_gender_quantized.pb")
    os.environ["HFR_FUSE_KEEP"] = "1"
    m = hfr.HfrModel(pb, "input_1:0", ["global_pooling/Mean:0"], precision="bf16", input_hw=192)
    m.keep_activations(True)
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.randint(0, 256, (70, 192, 192, 3)).astype(np.uint8)).cuda()
    layers = m.plan()["layers"]
    def run(b):
        m.forward(x[:b].contiguous())
        torch.cuda.synchronize()
        return [m.layer_output(li, b).clone() for li in range(len(layers))]
    a70, b70, a64 = run(70), run(70), run(64)
    for li, L in enumerate(layers):
        d_rep = (a70[li] - b70[li]).abs().max().item()
        d_b = (a70[li][:64] - a64[li]).abs()
        nbad = (d_b > 0).sum().item()
        msg = f"layer {li:2d} {L['kind']:5s} {L['name'][:26]:26s} repeat-diff {d_rep:.4g}  B70-vs-B64 diff {d_b.max().item():.4g} ({nbad} elems)"
        if nbad and L['kind'] in ('pw', 'conv', 'stem'):
            idx = (d_b > 0).nonzero()
            imgs = sorted(set(idx[:, 0].tolist()))[:10]
            msg += f" imgs {imgs} first {idx[:3].tolist()}"
        print(msg)


def diag_knn():
    import hse_facerec_tf_b200 as hfr
    for prec in ("bf16", "tf32"):
        rs = np.random.RandomState(0)
        g = rs.randn(3000, 256).astype(np.float32)
        q = g[rs.randint(0, 3000, 200)] + 0.01 * rs.randn(200, 256).astype(np.float32)
        d = ((q[:, None, :].astype(np.float64) - g[None].astype(np.float64)) ** 2).sum(-1)
        ref = d.argmin(1)
        try:
            clf = hfr.KNeighborsClassifier(precision=prec).fit(g, np.arange(3000))
            dist, ind = clf.kneighbors(q)
            print(f"knn {prec}: agree {(ind[:, 0] == ref).mean():.3f}  max|d-dref| {np.abs(dist[:, 0] - np.sqrt(d.min(1))).max():.3g}")
            if (ind[:, 0] != ref).any():
                print("   first got", ind[:10, 0].tolist(), "ref", ref[:10].tolist())
        except Exception:
            traceback.print_exc()


def diag_model():
    import hse_facerec_tf_b200 as hfr
    from oracle.tfnet import preprocess_rgb_u8
    from tests.helpers import run_plan_cpu
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pb = os.path.join(root, "tests", "golden", "age_gender_quantized.pb")
    crops = np.load(os.path.join(root, "tests", "golden", "face_crops_u8.npz"))["c224"][:2]
    for prec in ("fp32", "tf32", "bf16"):
        try:
            m = hfr.HfrModel(pb, "input_1:0", ["global_pooling/Mean:0"], precision=prec)
            m.keep_activations(True)
            m.forward(torch.from_numpy(crops).cuda())
            torch.cuda.synchronize()
            _, kept = run_plan_cpu(m, preprocess_rgb_u8(crops), keep=True)
            for li, L in enumerate(m.plan()["layers"]):
                got = m.layer_output(li, 2).cpu().numpy().reshape(kept[li].shape)
                err = np.abs(got - kept[li]).max()
                print(f"model {prec} layer {li:2d} {L['kind']:5s} {L['name'][:28]:28s} max|err| {err:9.4g}  ref range {np.abs(kept[li]).max():8.3f} nan {np.isnan(got).sum()}")
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "dw", "knn", "model"]
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for w in what:
        print(f"===== {w}")
        try:
            {"gemm": diag_gemm, "dw": diag_dw, "knn": diag_knn, "model": diag_model, "conv": diag_conv, "stem": diag_stem, "win": diag_win, "dwpw": diag_dwpw, "fuserace": diag_fuserace}[w]()
        except Exception:
            traceback.print_exc()
