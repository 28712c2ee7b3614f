"""Host-side launch heuristics (no GPU): the GEMM tile choice must reproduce the kernel selection recorded in the
committed ncu launch list of the ResNet-50 benchmark step, and the k-NN gallery split must keep its invariants."""
import csv
import ctypes as C
import os
import sys

import numpy as np

import bench
import hse_facerec_tf_b200 as hfr
from hse_facerec_tf_b200._lib import check, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tile_choice(m, n, k, taps, sms=148):
    ctas, bn = C.c_int(), C.c_int()
    check(lib.hfr_debug_gemm_tile_choice(m, n, k, taps, sms, C.byref(ctas), C.byref(bn)))
    return ctas.value, bn.value


def pair_config(m, k0, k1, n1, n2, has_res=1, prec=2):
    vals = [C.c_int() for _ in range(6)]
    check(lib.hfr_debug_gemm_pair_config(m, k0, k1, n1, n2, has_res, prec, *[C.byref(v) for v in vals]))
    return [v.value for v in vals]          # eligible, nbuf, pf, na, stages, smem bytes


def test_tile_choice_reproduces_the_profiled_resnet50_step():
    spec = bench.model_spec("resnet50")
    m = hfr.HfrModel(spec["path"], spec["input"], spec["outputs"], input_hw=spec["hw"], device=None, precision="bf16")
    layers = m.plan()["layers"]
    names, seen = [], set()
    with open(os.path.join(ROOT, "profiles", "r2_resnet50_launches.csv")) as f:
        for r in csv.DictReader(l for l in f if not l.startswith("==")):
            if r["ID"] not in seen:
                seen.add(r["ID"])
                names.append(r["Kernel Name"])
    seq = []            # one entry per launch of an eager step, in order (see tools/traffic_from_launches.py)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from traffic_from_launches import kcat_absorbed
    sub_outs = {L["out"] for L in layers if L["kind"] == "subsample"}
    skip = False
    for i, L in enumerate(layers):
        if L["kind"] == "subsample" and all(U["kind"] == "pw" for U in layers if U["in"] == L["out"]):
            continue    # bypassed: its consumers gather through the im2col map
        if skip:        # the 'reduce' half of a gemm_pair_kernel launch
            skip = False
            continue
        if kcat_absorbed(layers, i, sub_outs):
            continue    # 'increase' absorbed into the projection shortcut's GEMM (K-concatenation)
        k0 = layers[i - 1]["cin"] if i > 0 and kcat_absorbed(layers, i - 1, sub_outs) else 0
        pair = None
        if L["kind"] == "pw" and i + 1 < len(layers) and L["in"] not in sub_outs:
            N = layers[i + 1]   # launch.cu: gemm_pair_eligible
            if (N["kind"] == "pw" and N["in"] == L["out"] and N.get("in2", -1) < 0 and L["cout"] % 128 == 0
                    and N["cout"] in (64, 128, 256) and (k0 + L["cin"] + 63) // 64 <= 4):
                pair, skip = N, True
        seq += [(L, None, 0), (L, None, 0)] if L["kind"] == "stem" else [(L, pair, k0)]
    names = names[:len(seq)]
    assert len(names) == len(seq) == 44
    batch, checked, pairs = 256, 0, 0
    for (L, pair, k0), name in zip(seq, names):
        if pair is not None:                                   # <T, N2, NBUF, PF>
            assert "gemm_pair_kernel" in name, (L["name"], name)
            targs = name.split("<")[1].split(">")[0].replace(" ", "").split(",")      # <T, N2, NBUF, PF>
            assert int(targs[1]) == pair["cout"]
            M = batch * L["hw_out"][0] * L["hw_out"][1]
            e, nbuf, pf, na, stages, smem = pair_config(M, k0, L["cin"], L["cout"], pair["cout"], has_res=int(k0 == 0))
            assert e == 1 and (nbuf, pf) == (int(targs[2]), int(targs[3])), (L["name"], name, nbuf, pf)
            pairs += 1
            continue
        if "gemm_tc_kernel" not in name:
            continue
        targs = name.split("<")[1].split(">")[0].replace(" ", "").split(",")      # <T, BLOCK_N, EPI, AMODE, CTAS>
        bn_seen, amode, ctas_seen = int(targs[1]), int(targs[3]), int(targs[4])
        taps = L["k"][0] * L["k"][1] if L["kind"] == "conv" else (1 if amode == 1 else 0)
        M, N, K = batch * L["hw_out"][0] * L["hw_out"][1], L["cout"], k0 + L["cin"] * max(taps, 1)
        assert tile_choice(M, N, K, taps) == (ctas_seen, bn_seen), (L["name"], M, N, K, taps)
        checked += 1
    assert pairs == 8 and checked == 49 - 16 - 4


def test_gemm_pair_eligibility_and_configuration():
    """The launcher's own predicate and configuration choice for the fused 'increase' + 'reduce' launch (launch.cu:
    gemm_pair_eligible / gemm_pair_config), on ResNet-50's seams at batch 256 and on the shapes that must be refused."""
    M = 256 * 56 * 56
    # stage 2 (K1 = 64): two A buffers, four staging buffers with the residual two chunks ahead, >= 4 ring slots
    assert pair_config(M, 0, 64, 256, 64) == [1, 4, 2, 2, 4, 1024 + (2 + 4 + 8) * 16384 + 512]
    # ... and its K-concatenated first block (64 + 64 columns, no residual tensor): three staging buffers
    e, nbuf, pf, na, stages, smem = pair_config(M, 64, 64, 256, 64, has_res=0)
    assert (e, nbuf, pf, na) == (1, 3, 1, 2) and stages >= 4
    # stage 3 (K1 = 128) and stage 4 (K1 = 256: one resident A buffer of 64 KB)
    assert pair_config(M // 4, 0, 128, 512, 128)[:4] == [1, 3, 1, 2]
    assert pair_config(M // 16, 0, 256, 1024, 256)[:4] == [1, 3, 1, 1]
    for args in ((M, 0, 64, 256, 64), (M, 64, 64, 256, 64), (M // 4, 0, 128, 512, 128), (M // 16, 0, 256, 1024, 256)):
        e, nbuf, pf, na, stages, smem = pair_config(*args)
        assert e == 1 and 4 <= stages <= 8 and smem <= 227 * 1024 and 1 <= pf < nbuf
    # refused: stage 5 (the second accumulator would need 512 TMEM columns), A rows beyond the resident buffer (K1 = 512;
    # tf32 stage 4: 256 floats = 8 k-blocks), a first GEMM whose width is not whole 128-column tiles, fp32 mode
    assert pair_config(M // 64, 0, 512, 2048, 512)[0] == 0
    assert pair_config(M // 16, 0, 512, 1024, 256)[0] == 0
    assert pair_config(M // 16, 0, 256, 1024, 256, prec=1)[0] == 0
    assert pair_config(M // 4, 0, 128, 512, 128, prec=1)[:2] == [1, 3]
    assert pair_config(M, 0, 64, 192, 64)[0] == 0
    assert pair_config(M, 0, 64, 256, 64, prec=0)[0] == 0


def test_tile_choice_policy_edges():
    assert tile_choice(50176, 256, 1024, 0) == (2, 256)        # long-K 1x1: CTA pair
    assert tile_choice(50176, 1024, 256, 0) == (1, 128)        # short-K 1x1: 128-wide single-CTA tiles
    assert tile_choice(50176, 256, 2304, 9) == (2, 256)        # 3x3 implicit GEMM: CTA pair
    assert tile_choice(128 * 3, 256, 2304, 9)[0] == 1          # odd number of 128-pixel blocks: no pair for im2col
    assert tile_choice(49, 1024, 1024, 0) == (1, 64)           # batch-1 tail layer: not even one wave -> narrow tiles
    assert tile_choice(802816, 64, 64, 0) == (1, 64)           # N = 64 has a single choice
    assert tile_choice(9216, 512, 512, 0) == (1, 128)          # MobileNet-192 tail at B = 64


def test_knn_plan_invariants():
    rs = np.random.RandomState(0)
    cases = [(100000, 1000000), (1, 1), (300, 5000), (130, 30000), (4096, 125000), (7, 10 ** 7)]
    cases += [(int(rs.randint(1, 200000)), int(rs.randint(1, 3000000))) for _ in range(50)]
    for nq, n in cases:
        sp, per = C.c_int(), C.c_int()
        check(lib.hfr_debug_knn_plan(nq, n, C.byref(sp), C.byref(per)))
        nb = (n + 255) // 256
        assert 1 <= per.value <= max(64, 1) and sp.value >= 1
        assert sp.value * per.value >= nb                       # every gallery block belongs to a split
        assert (sp.value - 1) * per.value < nb                  # and no split is empty


def test_bench_books_fused_launches_with_their_own_work():
    """bench.py: per-layer event times (-1 where a layer had no launch of its own) -> one row per launch.  On the
    ResNet-50 plan with the layers the device plan fuses marked -1: 43 rows (44 launches, the stem's two count once),
    24 of them 1x1 GEMMs; the flops of the network are conserved, the bytes drop by what the fusions remove, and the
    1x1 class carries the byte count of the committed benchmark line."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from traffic_from_launches import kcat_absorbed
    spec = bench.model_spec("resnet50")
    m = hfr.HfrModel(spec["path"], spec["input"], spec["outputs"], input_hw=spec["hw"], device=None, precision="bf16")
    layers = m.plan()["layers"]
    work = bench.plan_work(m.plan(), 2)
    sub_outs = {L["out"] for L in layers if L["kind"] == "subsample"}
    t = [1.0] * len(layers)
    for i, L in enumerate(layers):
        if L["kind"] == "subsample" or kcat_absorbed(layers, i, sub_outs):
            t[i] = -1.0
    for i in range(len(layers) - 1):        # the eight seams: 'reduce' layers fused into the preceding launch
        A, B = layers[i], layers[i + 1]
        k0 = layers[i - 1]["cin"] if i > 0 and kcat_absorbed(layers, i - 1, sub_outs) else 0
        if (A["kind"] == "pw" and B["kind"] == "pw" and B["in"] == A["out"] and B["in2"] < 0 and A["in"] not in sub_outs
                and t[i] > 0 and pair_config(256 * A["hw_out"][0] * A["hw_out"][1], k0, A["cin"], A["cout"], B["cout"])[0]):
            t[i + 1] = -1.0
    merged = bench.merge_fused_launches(work, t)
    kinds = [w["kind"] for w, _ in merged]
    assert len(merged) == 43 and kinds.count("pw") == 24 and kinds.count("conv") == 16
    assert sum(" + " in w["name"] for w, _ in merged) == 8 and sum("(+)" in w["name"] for w, _ in merged) == 4
    flops_all = sum(w["flops"] for w in work)
    assert abs(sum(w["flops"] for w, _ in merged) - flops_all) < 1e-6 * flops_all
    pw_bytes = sum(w["bytes"] for w, _ in merged if w["kind"] == "pw")
    pw_bytes_unfused = sum(w["bytes"] for w in work if w["kind"] == "pw")
    assert pw_bytes < 0.75 * pw_bytes_unfused
    line = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_all.json")))
    assert line["roofline"]["bound"] == "hbm" and line["kernels"]["pw"]["launches"] == 24
    assert abs(pw_bytes * 256 / 24 - line["roofline"]["algorithmic_bytes_per_launch"]) < 2
