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
            assert int(name.split("<")[1].split(">")[0].replace(" ", "").split(",")[1]) == pair["cout"]
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
