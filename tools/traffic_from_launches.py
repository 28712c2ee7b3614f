"""profiles/traffic_<workload>.json from an ncu launch list (gpu time + dram bytes per launch), one bench step.
usage: python tools/traffic_from_launches.py <launches.csv> <workload> <launches per step> [steps to skip]"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def kernel_class(name):
    """class from the kernel name alone (fallback when the plan is not available)"""
    if "gemm_pair_kernel" in name:
        return "pw"
    if "gemm_tc_kernel" in name:
        mode = name.split("<")[1].split(">")[0].replace(" ", "").split(",")
        epi, amode = mode[2], mode[3]
        if epi == "1":
            return "knn"
        return {"0": "pw", "1": "conv", "2": "stem"}[amode]
    if "conv_window_kernel<4" in name.replace(" ", ""):
        return "stem"
    for key, cls in (("conv_window", "conv"), ("stem_s2d", "stem"), ("stem_conv", "stem"), ("dwconv3x3", "dw"), ("dwpw", "dwpw"),
                     ("maxpool", "maxpool"), ("subsample", "subsample"), ("gap_kernel", "gap"), ("fc_kernel", "fc"), ("dense_heads", "fc")):
        if key in name:
            return cls
    return "other"


def kcat_absorbed(layers, i, sub_outs):
    """api.cu kcat_candidate (bf16): layer i is a linear 1x1 convolution whose only reader is the next 1x1 convolution's
    residual input - the two run as one K-concatenated GEMM and layer i is not launched."""
    if i + 1 >= len(layers):
        return False
    A, B = layers[i], layers[i + 1]
    if A["kind"] != "pw" or B["kind"] != "pw" or A.get("act", "none") not in ("none", 0) or A.get("in2", -1) >= 0:
        return False
    if B.get("in2", -1) != A["out"] or B["in"] == A["out"] or A["cout"] != B["cout"] or A["hw_out"] != B["hw_out"]:
        return False
    if A["cin"] % 64 or B["cin"] % 64 or A["in"] in sub_outs:
        return False
    return not any(U is not B and (U["in"] == A["out"] or U.get("in2", -1) == A["out"]) for U in layers)


def plan_classes(workload):
    """Class of every launch of one eager step, in launch order, from the compiled plan (host-only load): the same
    layer kinds bench.py books times under.  A bf16 stem is two launches (space-to-depth + window conv); a subsample
    whose consumers are all 1x1 layers is not launched (they fetch the strided pixels through an im2col map)."""
    sys.path.insert(0, ROOT)
    import bench
    import hse_facerec_tf_b200 as hfr
    spec = bench.model_spec(workload)
    m = hfr.HfrModel(spec["path"], spec["input"], spec["outputs"], input_hw=spec["hw"], device=None, precision="bf16")
    layers = m.plan()["layers"]
    seq = []
    sub_outs = {L["out"] for L in layers if L["kind"] == "subsample"}
    fused_into_prev = False
    for i, L in enumerate(layers):
        if L["kind"] == "subsample":
            users = [U for U in layers if U["in"] == L["out"] or U.get("in2", -1) == L["out"]]
            if users and all(U["kind"] == "pw" and U["in"] == L["out"] for U in users):
                continue
        if fused_into_prev:          # second half of a gemm_pair_kernel launch (launch.cu: gemm_pair_eligible, bf16)
            fused_into_prev = False
            continue
        if kcat_absorbed(layers, i, sub_outs):   # 'increase' absorbed by the projection shortcut's GEMM (api.cu: plan_kcat)
            continue
        if L["kind"] == "pw" and i + 1 < len(layers) and L["in"] not in sub_outs:
            N = layers[i + 1]
            k0 = layers[i - 1]["cin"] if i > 0 and kcat_absorbed(layers, i - 1, sub_outs) else 0
            fused_into_prev = (N["kind"] == "pw" and N["in"] == L["out"] and N.get("in2", -1) < 0 and L["cout"] % 128 == 0
                               and N["cout"] in (64, 128, 256) and (k0 + L["cin"] + 63) // 64 <= 4)
        seq += ["stem", "stem"] if L["kind"] == "stem" else [L["kind"]]
    return seq


def main():
    path, workload, per_step = sys.argv[1], sys.argv[2], int(sys.argv[3])
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    lines = [l for l in open(path) if not l.startswith("==")]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    launches = list(by.values())[skip * per_step:(skip + 1) * per_step]
    try:
        seq = plan_classes(workload)
    except Exception as e:  # noqa: BLE001
        print("plan not available, classifying by kernel name:", e, file=sys.stderr)
        seq = None
    if seq is not None and len(seq) != len(launches):
        print(f"plan has {len(seq)} launches, list has {len(launches)}: classifying by kernel name", file=sys.stderr)
        seq = None
    agg = {}
    for i, d in enumerate(launches):
        cls = seq[i] if seq else kernel_class(d["name"])
        c = agg.setdefault(cls, dict(launches=0, us=0.0, dram_bytes=0.0, l2_bytes=0.0))
        c["launches"] += 1
        c["us"] += d.get("gpu__time_duration.sum", 0.0)
        c["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        c["l2_bytes"] += d.get("lts__t_bytes.sum", 0.0)
    tot = sum(c["us"] for c in agg.values()) or 1.0
    out = {"workload": workload, "source": os.path.basename(path), "note": "ncu --clock-control none --cache-control none, "
           "one eager bench step; per-launch averages", "classes": {}}
    for k, c in agg.items():
        out["classes"][k] = dict(launches=c["launches"], us_total=round(c["us"], 1), share=round(c["us"] / tot, 4),
                                 dram_bytes_per_launch=round(c["dram_bytes"] / c["launches"]),
                                 l2_bytes_per_launch=round(c["l2_bytes"] / c["launches"]))
    dst = os.path.join(ROOT, "profiles", f"traffic_{workload}.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
