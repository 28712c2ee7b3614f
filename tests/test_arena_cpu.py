"""Arena planning (api.cu: plan_arena) without a GPU: the activation arena re-uses the slots of dead values, and the fused
launches widen what "dead" has to mean.  Checked on the compiled plans of every benchmark workload, for the plan a device
handle executes (K-concatenation planned through HFR_PLAN_ASSUME_KCAT=1) and for the plain one:

  * two values that overlap in memory are never live at the same time (a layer's output is placed BEFORE the values that
    layer reads for the last time are released);
  * the seam rule: when two 1x1 convolutions may run as one gemm_pair_kernel launch, the second layer's output is written
    while other CTAs still read the first layer's inputs, so it must not overlap any of them.  (Without that rule
    first-fit handed ResNet-50's `conv2_3_1x1_reduce` output the first quarter of the slot `conv2_2_relu`'s shortcut
    had just left: units 1-3 would overwrite shortcut rows unit 0 may not have read yet.)
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import hse_facerec_tf_b200 as hfr  # noqa: E402


def overlaps(a, b):
    return not (a[2] + a[3] <= b[2] or b[2] + b[3] <= a[2])


@pytest.mark.parametrize("assume_kcat", [False, True])
@pytest.mark.parametrize("workload,precision", [("resnet50", "bf16"), ("resnet50", "tf32"), ("mobilenet192", "bf16"),
                                                ("agegender224", "bf16"), ("agegender224", "tf32")])
def test_arena_slots_are_reused_only_by_dead_values(workload, precision, assume_kcat, monkeypatch):
    if assume_kcat:
        monkeypatch.setenv("HFR_PLAN_ASSUME_KCAT", "1")
    else:
        monkeypatch.delenv("HFR_PLAN_ASSUME_KCAT", raising=False)
    spec = bench.model_spec(workload)
    m = hfr.HfrModel(spec["path"], spec["input"], spec["outputs"], input_hw=spec["hw"], device=None, precision=precision)
    plan = m.plan()
    layers = plan["layers"]
    arena = {a[0]: a for a in plan["arena"]}          # value -> [value, producer layer, offset, bytes, released after layer]
    assert arena and max(a[2] + a[3] for a in arena.values()) == plan["arena_bytes_per_image"]
    outputs = {o["value"] for o in plan["outputs"]}
    for v, a in arena.items():
        assert (a[4] == -1) == (v in outputs), f"value {v}: only the model outputs stay in the arena for good ({a})"
        assert a[4] == -1 or a[4] >= a[1]
    # liveness: [producer, release]; a slot freed after layer r can be taken by the output of layer r + 1 at the earliest
    vals = sorted(arena.values())
    for i, x in enumerate(vals):
        for y in vals[i + 1:]:
            if not overlaps(x, y):
                continue
            x_end = x[4] if x[4] >= 0 else 10 ** 9
            y_end = y[4] if y[4] >= 0 else 10 ** 9
            assert x_end < y[1] or y_end < x[1], f"values {x} and {y} share memory while both are live"
    # every reader finds its inputs alive
    absorbed = set()
    for li, L in enumerate(layers):
        for v in (L["in"], L["in2"]):
            if v > 0 and v in arena:
                a = arena[v]
                assert a[1] < li and (a[4] == -1 or a[4] >= li), f"layer {li} {L['name']} reads value {v} outside its life {a}"
            elif v > 0:
                absorbed.add(v)      # never materialised: a bypassed gather or a K-concatenated 'increase'
    if workload == "resnet50":
        assert len(absorbed) == (4 + 3 if assume_kcat else 3), absorbed   # 3 strided gathers (+ 4 'increase' tensors)
    # seam rule
    seams = 0
    for li in range(len(layers) - 1):
        A, B = layers[li], layers[li + 1]
        if not (A["kind"] == "pw" and B["kind"] == "pw" and B["in"] == A["out"] and B["in2"] < 0 and A["out"] in arena):
            continue
        seams += 1
        z = arena[B["out"]]
        ins = [A["in"], A["in2"]]
        if li > 0 and layers[li - 1]["out"] == A["in2"] and A["in2"] not in arena:   # K-concatenated: reads that layer's input
            ins = [A["in"], layers[li - 1]["in"]]
        for v in ins:
            if v > 0 and v in arena:
                assert not overlaps(z, arena[v]), f"seam {A['name']} -> {B['name']}: output {z} overlaps input {arena[v]}"
    if workload == "resnet50":
        assert seams >= 8


def test_fused_dense_tail_keeps_the_pooled_vector_alive():
    """dense_heads_kernel runs the hidden Dense layer and both heads in one launch: with only the heads requested (the
    pooled embedding is then an ordinary intermediate) its slot must not be handed to a head output."""
    spec = bench.model_spec("agegender224")
    m = hfr.HfrModel(spec["path"], spec["input"], ["age_pred/Softmax:0", "gender_pred/Sigmoid:0"], input_hw=0, device=None,
                     precision="bf16")
    plan = m.plan()
    layers = plan["layers"]
    arena = {a[0]: a for a in plan["arena"]}
    fc = [i for i, L in enumerate(layers) if L["kind"] == "fc"]
    assert len(fc) == 3 and fc == list(range(fc[0], fc[0] + 3))
    pooled = arena[layers[fc[0]]["in"]]
    assert pooled[4] == fc[-1], pooled                      # released after the LAST layer of the fused group
    for i in fc:
        assert not overlaps(arena[layers[i]["out"]], pooled)
