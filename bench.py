#!/usr/bin/env python
"""bench.py - benchmark of the hot path (contract: DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W]            # headline + every other BASELINE workload
  python bench.py --only mobilenet192|agegender224|resnet50|knn [--precision bf16|tf32] [--layers]
  python bench.py --impl reference ...                            # the reference's CPU path on the host cores

One "step" = one pass of the hot path over one batch of synthetic input.  Rank 0 prints ONE JSON line.  The top level
is the headline workload (ResNet-50, batch 256, bf16 - BASELINE.json configs[1]); `workloads` holds the same record
(value, e2e, roofline, cpu_baseline, clocks) for the other configurations BASELINE.json's metric names:
MobileNet-192 batch 64, the age/gender MobileNet-224 batch 256, ResNet-50 in tf32 mode and the 1-NN identification
(100k queries x 1M x 1024 gallery, row-sharded over the N ranks, NCCL all-gather merge: the one collective of the path).
  value   whole-job throughput, inputs resident in HBM, CUDA-graph replay, CUDA-event timed, max over ranks
  e2e     the same metric through the host-buffer C-ABI call (pinned host input -> H2D -> run -> D2H of the result)
  roofline  dominant kernel class: algorithmic bytes|flops per launch / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  oracle (torch-CPU port; sklearn for 1-NN = the reference's real dependency) on a bounded sample,
          rank 0 at N=1 only
The reference arm never imports the product package (no libhfr.so in that process).
"""
from __future__ import annotations

import os
import sys

_NCPU = os.cpu_count() or 1
if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core, set before numpy / torch / sklearn load
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(_NCPU)

import argparse
import importlib.util
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PB = os.path.join(ROOT, "tests", "golden", "age_gender_quantized.pb")

WORKLOADS = {
    # name: (description, per-GPU batch, input size, metric, unit)
    "mobilenet192": ("MobileNet-192 embedding extraction (real VGGFace2 MobileNet body from the shipped graph, run at "
                     "192x192), batch 64 synthetic uint8 RGB crops", 64, 192, "face images/sec", "images/s"),
    "agegender224": ("age/gender MobileNet-224 + heads (age softmax top-2 expectation, gender sigmoid, 1024-D feature), "
                     "batch 256 synthetic crops", 256, 224, "face images/sec", "images/s"),
    "resnet50": ("ResNet-50 embedding extraction (Caffe-style VGGFace2 resnet50_ft topology, synthetic weights), batch "
                 "256 synthetic 224x224 crops", 256, 224, "face images/sec", "images/s"),
    "knn": ("1-NN identification: 100k queries vs 1M x 1024 L2-normalised synthetic gallery (row-sharded over the "
            "GPUs, NCCL all-gather top-1 merge)", 100_000, 1024, "1-NN queries/sec", "queries/s"),
}
HEADLINE = ("resnet50", "bf16")
OTHERS = [("mobilenet192", "bf16"), ("agegender224", "bf16"), ("resnet50", "tf32"), ("knn", "bf16")]


def _synth():
    """tests' fixture writer (hse_facerec_tf_b200/synth.py) loaded BY PATH: importing the package would dlopen
    libhfr.so, which the reference arm must not do."""
    spec = importlib.util.spec_from_file_location("_hfr_synth", os.path.join(ROOT, "hse_facerec_tf_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}
    if os.path.exists(p):
        d = json.load(open(p))
        out = {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d["bf16_tflops_sustained"],
               "source": "measured"}
    # tf32: MEASURED_PEAKS.json has no tf32 figure; profiles/peaks_probe.json holds ours, taken with the same probe
    q = os.path.join(ROOT, "profiles", "peaks_probe.json")
    if os.path.exists(q):
        d = json.load(open(q))
        out["tensor_tf32"] = d["tf32_tflops_sustained"]
        out["tf32_source"] = "profiles/peaks_probe.json (same probe as MEASURED_PEAKS.json, tf32 matmul)"
    else:
        out["tensor_tf32"] = out["tensor"] / 2
        out["tf32_source"] = "bf16 sustained / 2 (no tf32 probe available)"
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------ workloads
def synth_images(batch, size, seed):
    return np.random.RandomState(seed).randint(0, 256, (batch, size, size, 3)).astype(np.uint8)


def model_spec(workload):
    if workload == "mobilenet192":
        return dict(path=PB, input="input_1:0", outputs=["global_pooling/Mean:0"], hw=192, bgr=True, imagenet=True)
    if workload == "agegender224":
        return dict(path=PB, input="input_1:0",
                    outputs=["age_pred/Softmax:0", "gender_pred/Sigmoid:0", "global_pooling/Mean:0"], hw=0, bgr=True,
                    imagenet=True)
    if workload == "resnet50":
        return dict(path=_synth().ensure_resnet50_pb(), input="input:0", outputs=["pool5_7x7_s1:0"], hw=0, bgr=True,
                    imagenet=False)
    raise ValueError(workload)


def input_rotation(batch, size):
    """distinct input batches the timed loop rotates over: enough that the inputs alone exceed the 126 MB L2"""
    in_bytes = batch * size * size * 3
    return max(2, min(32, -(-140_000_000 // in_bytes))), in_bytes


def workload_config(workload, world, batch=0, gallery=1_000_000):
    """The `config` object of a workload: identical in the product arm and in the reference arm."""
    desc, b, size, metric, unit = WORKLOADS[workload]
    b = batch or b
    if workload == "knn":
        return dict(workload="knn", description=desc, queries=b, gallery=gallery, dim=size, shards=world,
                    parallelism=f"gallery row-sharded x{world}",
                    l2_policy=f"gallery shard {(gallery // world) * size * 2 / 1e9:.1f} GB (bf16) >> L2")
    nrot, in_bytes = input_rotation(b, size)
    return dict(workload=workload, description=desc, per_gpu_batch=b, global_batch=b * world,
                input=f"{size}x{size}x3 uint8", parallelism=f"dp{world}",
                l2_policy=f"inputs rotate over {nrot} distinct batches ({nrot * in_bytes / 1e6:.0f} MB > L2)")


def plan_work(plan, es):
    """Algorithmic work per image from the compiled plan: per layer (flops, bytes moved in+out; weights ignored)."""
    rows = []
    for L in plan["layers"]:
        hin, win = L["hw_in"]
        ho, wo = L["hw_out"]
        k = L["k"][0] * L["k"][1]
        kind = L["kind"]
        if kind in ("pw", "conv", "stem"):
            flops = 2.0 * ho * wo * L["cout"] * L["cin"] * k
        elif kind == "dw":
            flops = 2.0 * ho * wo * L["cout"] * 9
        elif kind in ("fc", "head"):
            flops = 2.0 * L["cin"] * L["cout"]
        else:
            flops = 0.0
        in_b = hin * win * L["cin"] * (1 if kind == "stem" else es)
        if kind == "subsample":
            in_b = ho * wo * L["cin"] * es      # only every stride-th pixel is read
        out_b = ho * wo * L["cout"] * (4 if kind in ("gap", "fc", "head") else es)
        if kind in ("fc", "head"):
            in_b, out_b = L["cin"] * 4, L["cout"] * 4
        res_b = ho * wo * L["cout"] * es if L["in2"] >= 0 else 0
        rows.append(dict(kind=kind, name=L["name"], flops=flops, bytes=float(in_b + res_b + out_b), out_bytes=float(out_b),
                         in_bytes=float(in_b), res_bytes=float(res_b), out=L["out"], in2=L["in2"], **{"in": L["in"]}))
    return rows


def merge_fused_launches(work, per_step):
    """Per-layer work rows (plan_work) + per-layer times (-1: nothing was launched for that layer) -> one (row, time) per
    LAUNCH, with the algorithmic work the launch really has:
      * a bypassed gather (strided 1x1 consumers fetch through an im2col map) disappears;
      * an 'increase' 1x1 convolution absorbed by the projection shortcut it is the residual of (api.cu: plan_kcat) adds its
        flops and its INPUT bytes to that GEMM, whose residual tensor no longer exists;
      * two 1x1 convolutions in one gemm_pair_kernel launch: both layers' flops, the first layer's bytes plus the second
        layer's OUTPUT only (its input never leaves the SM)."""
    merged = []
    i = 0
    absorbed = None
    while i < len(work):
        w, t = dict(work[i]), per_step[i]
        if t < 0:
            if (w["kind"] == "pw" and i + 1 < len(work) and work[i + 1]["kind"] == "pw" and per_step[i + 1] >= 0
                    and work[i + 1]["in2"] == w["out"]):
                absorbed = w
            i += 1
            continue
        if absorbed is not None:
            w["name"] = absorbed["name"] + " (+) " + w["name"]
            w["flops"] += absorbed["flops"]
            w["bytes"] += absorbed["in_bytes"] - w["res_bytes"]
            absorbed = None
        if (w["kind"] == "pw" and i + 1 < len(work) and work[i + 1]["kind"] == "pw" and per_step[i + 1] < 0
                and work[i + 1]["in"] == w["out"] and work[i + 1]["in2"] < 0):
            w2 = work[i + 1]
            w["name"] += " + " + w2["name"]
            w["flops"] += w2["flops"]
            w["bytes"] += w2["out_bytes"]
            i += 1
        merged.append((w, t))
        i += 1
    return merged


KERNEL_NAMES = {"pw": "gemm_tc_kernel / gemm_pair_kernel (1x1 conv)", "fc": "dense_heads_kernel",
                "conv": "gemm_tc_kernel (im2col) / conv_window_kernel", "dw": "dwconv3x3_pipe_kernel"}


def bench_network(workload, precision, args, rank, world, dev):
    import torch
    import torch.distributed as dist
    import hse_facerec_tf_b200 as hfr
    desc, batch, size, metric, unit = WORKLOADS[workload]
    batch = args.batch or batch
    spec = model_spec(workload)
    model = hfr.HfrModel(spec["path"], spec["input"], spec["outputs"], input_hw=spec["hw"], device=f"cuda:{dev}",
                         precision=precision)
    es = 2 if precision == "bf16" else 4
    size = model.h
    nrot, in_bytes = input_rotation(batch, size)
    xs = [torch.from_numpy(synth_images(batch, size, 1000 * rank + i)).to(f"cuda:{dev}") for i in range(nrot)]
    outs = [[torch.empty((batch, d), dtype=torch.float32, device=f"cuda:{dev}") for d in model.out_dims]
            for _ in range(nrot)]
    stream = torch.cuda.Stream(device=dev)
    kw = dict(convert2BGR=spec["bgr"], imageNetUtilsMean=spec["imagenet"])
    steps, warmup = args.steps, max(args.warmup, 3)

    def step(i, graph=not args.no_graph):
        model.forward(xs[i % nrot], graph=graph, outs=outs[i % nrot], **kw)

    e2e_s, host_s = None, [0.0, 0.0]
    with torch.cuda.stream(stream):
        for i in range(warmup + nrot):          # warm-up (also captures one graph per input buffer)
            step(i)
        stream.synchronize()
        launches0 = hfr.launch_count()
        step(0, graph=False)
        stream.synchronize()
        launches_per_step = hfr.launch_count() - launches0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(dev)
        sampler.start()
        # untimed pre-roll of the same steps so that nvidia-smi (100 ms period) sees the GPU under this load even when
        # the K timed steps last only a few milliseconds; the sampler keeps running through the timed region
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 1.0:
            for i in range(8):
                step(i)
            stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step(i)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        clocks = sampler.stop()

        if not args.no_e2e:
            # ---- e2e: pinned host batch -> H2D -> run -> D2H of the result, through the host-buffer C-ABI call.
            # Streaming form (hfr_model_submit_host / hfr_model_wait_host, `depth` batches in flight): every step's
            # input still travels pinned host -> device and its result device -> pinned host inside the timed region,
            # but step i+1's upload and step i-1's download overlap step i's compute.
            depth = int(os.environ.get("HFR_BENCH_E2E_DEPTH", "3"))
            hx = [torch.from_numpy(synth_images(batch, size, 5000 + 1000 * rank + i)).pin_memory() for i in range(depth)]
            houts = [[torch.empty((batch, d), dtype=torch.float32).pin_memory().numpy() for d in model.out_dims]
                     for _ in range(depth)]
            hxn = [h.numpy() for h in hx]

            def e2e_run(n):
                for i in range(n):
                    ta = time.perf_counter()
                    if i >= depth:
                        model.wait_host(i % depth)
                    tb = time.perf_counter()
                    model.submit_host(i % depth, hxn[i % depth], houts[i % depth], graph=True, **kw)
                    host_s[0] += time.perf_counter() - tb
                    host_s[1] += tb - ta
                for sl in range(depth):
                    model.wait_host(sl)

            e2e_run(4)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            host_s[0] = host_s[1] = 0.0
            e2e_run(steps)
            torch.cuda.synchronize(dev)
            e2e_s = time.perf_counter() - t0

        # ---- per-kernel durations (CUDA events around every layer launch, eager, same stream)
        model.layer_timing(True)
        for i in range(min(steps, 20)):
            step(i, graph=False)
        lms, lsteps = model.layer_times()
        model.layer_timing(False)

    if world > 1:
        t = torch.tensor([ms, e2e_s or 0.0], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), (float(t[1]) if e2e_s is not None else None)
    plan = model.plan()
    out_dims = list(model.out_dims)
    model.close()
    del xs, outs
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    pk = peaks()
    tensor_peak = pk["tensor"] if precision == "bf16" else pk["tensor_tf32"]
    work = plan_work(plan, es)
    classes = {}
    per_step = [t / max(lsteps, 1) for t in lms]
    merged = merge_fused_launches(work, per_step)
    for w, t in merged:
        c = classes.setdefault(w["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        c["ms"] += t
        c["flops"] += w["flops"] * batch
        c["bytes"] += w["bytes"] * batch
        c["launches"] += 1
    kernels = {}
    for kname, c in classes.items():
        if c["ms"] <= 0:
            continue
        # each class is compared with BOTH roofs (algorithmic flops vs measured tensor peak of this precision,
        # algorithmic bytes vs measured HBM copy bandwidth); the binding roof is the one it sits closer to
        tf = c["flops"] / (c["ms"] * 1e-3) / 1e12
        gb = c["bytes"] / (c["ms"] * 1e-3) / 1e9
        frac_t, frac_h = tf / tensor_peak, gb / pk["hbm"]
        tensor_bound = kname in ("pw", "conv") and frac_t >= frac_h
        kernels[kname] = dict(ms_per_step=round(c["ms"], 5), launches=c["launches"],
                              bound="tensor" if tensor_bound else "hbm",
                              achieved=round(tf if tensor_bound else gb, 2),
                              unit="TFLOP/s" if tensor_bound else "GB/s",
                              frac=round(frac_t if tensor_bound else frac_h, 4),
                              tflops=round(tf, 2), hbm_gbs=round(gb, 1))
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    kd = kernels[dom]
    stem_name = "stem_s2d_kernel + conv_window_kernel" if precision == "bf16" else "stem_conv_kernel"
    roof = dict(kernel=KERNEL_NAMES.get(dom, stem_name if dom == "stem" else dom),
                bound=kd["bound"], achieved=kd["achieved"], peak=tensor_peak if kd["bound"] == "tensor" else pk["hbm"],
                unit=kd["unit"], frac=kd["frac"], traffic=None,
                peak_source=(pk["source"] + " (sustained)") if (kd["bound"] == "hbm" or precision == "bf16") else pk["tf32_source"],
                per_launch_ms=round(kd["ms_per_step"] / kd["launches"], 5),
                algorithmic_bytes_per_launch=round(classes[dom]["bytes"] / classes[dom]["launches"]),
                note="class times are eager CUDA-event pairs around each launch; `value` is CUDA-graph replay with "
                     "programmatic dependent launch, so the class times sum to slightly more than ms_per_step")
    # DRAM traffic per launch of the dominant class, from the committed ncu launch list of this workload (if present)
    tag = workload if precision == "bf16" else f"{workload}_{precision}"
    tpath = os.path.join(ROOT, "profiles", f"traffic_{tag}.json")
    if os.path.exists(tpath) and batch == WORKLOADS[workload][1]:
        tr = json.load(open(tpath)).get("classes", {}).get(dom)
        if tr:
            roof["traffic"] = tr["dram_bytes_per_launch"]
            roof["traffic_source"] = f"profiles/traffic_{tag}.json (ncu, share of step {tr['share']})"
    total = batch * world * steps
    rec = dict(metric=metric, value=round(total / (ms * 1e-3), 1), unit=unit, n_gpus=world, steps=steps, warmup=warmup,
               ms_per_step=round(ms / steps, 4), higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype=precision, data="synthetic", config=workload_config(workload, world, batch), cuda_graph=not args.no_graph)
    rec["per_gpu_value"] = round(rec["value"] / world, 1)     # `value` is the whole job; ratios against ONE host belong to this
    if e2e_s is not None:
        rec["e2e"] = dict(value=round(total / e2e_s, 1), unit=unit, h2d_bytes_per_step=in_bytes,
                          d2h_bytes_per_step=sum(out_dims) * 4 * batch,
                          note=f"hfr_model_submit_host/wait_host, {os.environ.get('HFR_BENCH_E2E_DEPTH', '3')} batches in "
                               f"flight; host in submit {host_s[0] / steps * 1e3:.3f} ms, in wait {host_s[1] / steps * 1e3:.3f} "
                               "ms per step. e2e can exceed `value`: the device-resident loop draws more power and sits "
                               "lower under the board's power cap")
    rec.update(gpu_launches=int(launches_per_step * steps), launches_per_step=int(launches_per_step), clocks=clocks,
               roofline=roof, kernels=kernels)
    if args.layers:
        rec["layers"] = [dict(kind=w["kind"], name=w["name"], us=round(t * 1e3, 2), gflop=round(w["flops"] * batch / 1e9, 2),
                              mb=round(w["bytes"] * batch / 1e6, 1)) for w, t in merged]
    return rec


# ------------------------------------------------------------------------------------------------------ 1-NN
def knn_data(n, nq, d, rank, world):
    """gallery = normalize(randn) per shard (seed + rank); queries = gallery rows + noise."""
    import torch
    n_local = n // world
    g = torch.Generator(device="cpu").manual_seed(rank)
    gal = torch.randn(n_local, d, generator=g)
    gal = gal / gal.norm(dim=1, keepdim=True)
    g0 = torch.Generator(device="cpu").manual_seed(12345)       # queries identical on every rank
    base = torch.randn(nq, d, generator=g0)
    q = base / base.norm(dim=1, keepdim=True)
    if rank == 0:                                               # plant the first nq/world queries' neighbours in shard 0
        m = min(nq, n_local)
        q[:m] = gal[:m] + 0.05 * torch.randn(m, d, generator=g0)
        q[:m] = q[:m] / q[:m].norm(dim=1, keepdim=True)
    return gal.contiguous(), q.contiguous()


def bench_knn(precision, args, rank, world, dev):
    import torch
    import torch.distributed as dist
    import hse_facerec_tf_b200 as hfr
    from hse_facerec_tf_b200.parallel import shard_rows
    desc, nq, d, metric, unit = WORKLOADS["knn"]
    n = args.gallery
    nq = args.batch or nq
    steps, warmup = args.steps, max(args.warmup, 3)
    gal, q = knn_data(n, nq, d, rank, world)
    if world > 1:
        qt = q.cuda(dev)
        dist.broadcast(qt, src=0)                  # rank 0's queries (with planted neighbours) everywhere
        q = qt.cpu()
        del qt
    clf = hfr.KNeighborsClassifier(1, 2, device=f"cuda:{dev}", precision=precision, sharded=world > 1)
    y_local = np.arange(rank * (n // world), (rank + 1) * (n // world)) % 10000
    clf.fit(gal.cuda(dev), y_local)
    qd = q.cuda(dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            clf.kneighbors(qd, return_distance=False)
        l0 = hfr.launch_count()
        clf.kneighbors(qd, return_distance=False)
        launches_per_step = hfr.launch_count() - l0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(dev)
        sampler.start()
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 1.0:     # untimed pre-roll: the clock sampler sees the GPU under this load
            clf.kneighbors(qd, return_distance=False)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            clf.kneighbors(qd, return_distance=False)   # includes the merge collective and the D2H of the indices
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        clocks = sampler.stop()
        e2e_s = None
        if not args.no_e2e:
            # e2e: pinned host queries in, indices out.  Sharded: every rank uploads ITS 1/N of the queries (the rows
            # it would have extracted itself) and the query block is all-gathered over NVLink
            a, b = shard_rows(nq, world, rank)
            hq = q[a:b].contiguous().pin_memory().numpy()
            clf.kneighbors(hq, return_distance=False, local_queries=world > 1, total_queries=nq)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                ind = clf.kneighbors(hq, return_distance=False, local_queries=world > 1, total_queries=nq)
            torch.cuda.synchronize(dev)
            e2e_s = time.perf_counter() - t0
        else:
            ind = clf.kneighbors(qd, return_distance=False)
    if world > 1:
        t = torch.tensor([ms, e2e_s or 0.0], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), (float(t[1]) if e2e_s is not None else None)
    del clf, qd, gal
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    m = min(nq, n // world)
    planted_ok = float((ind[:m, 0] == np.arange(m)).mean())
    pk = peaks()
    tensor_peak = pk["tensor"] if precision == "bf16" else pk["tensor_tf32"]
    flops = 2.0 * nq * (n // world) * d
    per_step_ms = ms / steps
    ach = flops / (per_step_ms * 1e-3) / 1e12
    cfg = workload_config("knn", world, nq, n)
    cfg["planted_neighbour_recall"] = planted_ok
    rec = dict(metric=metric, value=round(nq * steps / (ms * 1e-3), 1), unit=unit, n_gpus=world, steps=steps,
               warmup=warmup, ms_per_step=round(per_step_ms, 3), higher_is_better=True, scaling="strong",
               vs_baseline=None, dtype=precision, data="synthetic", config=cfg)
    rec["per_gpu_value"] = round(rec["value"] / world, 1)
    if e2e_s is not None:
        rec["e2e"] = dict(value=round(nq * steps / e2e_s, 1), unit=unit, h2d_bytes_per_step=(b - a) * d * 4,
                          d2h_bytes_per_step=nq * 8,
                          note="per rank: its 1/N block of the queries pinned host -> device, all-gathered over NVLink; "
                               "all indices back to the host")
    tpath = os.path.join(ROOT, "profiles", "traffic_knn.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) and world == 1 and nq == 100_000 and n == 1_000_000 else None
    rec.update(gpu_launches=int(launches_per_step * steps), launches_per_step=int(launches_per_step), clocks=clocks,
               roofline=dict(kernel="gemm_tc_kernel<EPI_KNN> (step time also holds rows_prep + finalize + merge)",
                             bound="tensor", achieved=round(ach, 2), peak=tensor_peak, unit="TFLOP/s",
                             frac=round(ach / tensor_peak, 4),
                             traffic=traffic["dram_bytes_per_launch"] if traffic else None,
                             **({"traffic_source": traffic["source"]} if traffic else {}),
                             algorithmic_flops_per_launch=flops,
                             peak_source=(pk["source"] + " (sustained)") if precision == "bf16" else pk["tf32_source"]))
    return rec


# ------------------------------------------------------------------------------------------------------ CPU legs
def cpu_network(workload, steps, warmup, batch=0, budget_s=None):
    """Oracle port (torch-CPU restatement of the frozen graph) on all host cores.  steps x the workload's batch (a
    bounded sample when budget_s cuts it short); returns (cpu_baseline dict, steps done, seconds per step)."""
    import torch
    from oracle.tfnet import GraphOracle, preprocess_rgb_u8
    desc, b, size, metric, unit = WORKLOADS[workload]
    b = batch or b
    spec = model_spec(workload)
    torch.set_num_threads(_NCPU)
    g = GraphOracle(spec["path"])
    size = spec["hw"] or (g.placeholder_shape(spec["input"]) or [0, 224])[1]
    nrot = min(input_rotation(b, size)[0], 4)
    xs = [preprocess_rgb_u8(synth_images(b, size, i), True, spec["imagenet"]) for i in range(nrot)]
    for i in range(warmup):
        g.run(list(spec["outputs"]), {spec["input"]: xs[i % nrot]})
    done, t0 = 0, time.perf_counter()
    for i in range(steps):
        g.run(list(spec["outputs"]), {spec["input"]: xs[i % nrot]})
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    cb = dict(value=round(done * b / dt, 2), unit=unit, cores=torch.get_num_threads(), kind="port",
              sample=f"{done} batches of {b} images at {size}x{size} through oracle/tfnet.py (torch-CPU fp32, {dt:.1f} s); "
                     "TensorFlow itself is not installable here")
    return cb, done, dt / max(done, 1)


def cpu_knn(gallery, steps, warmup, queries_per_step=300):
    """The reference's real dependency: sklearn KNeighborsClassifier(n_neighbors=1, p=2).kneighbors on a query subsample
    against the full gallery."""
    from sklearn import neighbors
    desc, nq, d, metric, unit = WORKLOADS["knn"]
    n = min(gallery, 1_000_000)
    gal, q = knn_data(n, queries_per_step * max(steps, 1), d, 0, 1)
    nn = neighbors.KNeighborsClassifier(n_neighbors=1, p=2).fit(gal.numpy(), np.arange(n) % 10000)
    qn = q.numpy()
    for _ in range(warmup):
        nn.kneighbors(qn[:32], return_distance=False)
    t0 = time.perf_counter()
    for i in range(steps):
        nn.kneighbors(qn[i * queries_per_step:(i + 1) * queries_per_step], return_distance=False)
    dt = time.perf_counter() - t0
    cb = dict(value=round(queries_per_step * steps / dt, 2), unit=unit, cores=_NCPU, kind="reference",
              sample=f"scikit-learn kneighbors: {steps} steps of {queries_per_step} of the {nq} queries vs the full {n} x {d} "
                     f"gallery ({dt:.1f} s), extrapolated linearly")
    return cb, steps, dt / max(steps, 1)


def reference_record(workload, args, steps, warmup):
    desc, batch, size, metric, unit = WORKLOADS[workload]
    if workload == "knn":
        cb, done, s_per = cpu_knn(args.gallery, steps, min(warmup, 1))
        ms_per_step = batch / cb["value"] * 1e3      # time this arm needs for one step of the product arm's size
    else:
        cb, done, s_per = cpu_network(workload, steps, warmup, args.batch)
        ms_per_step = s_per * 1e3
    v = cb["value"]
    return dict(impl="reference", metric=metric, value=v, unit=unit, n_gpus=args.gpus, steps=done, warmup=warmup,
                ms_per_step=round(ms_per_step, 1), higher_is_better=True,
                scaling="strong" if workload == "knn" else "weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(workload, args.gpus, args.batch, args.gallery), cpu_baseline=cb,
                e2e=dict(value=v, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores; same config / metric / unit /
    steps as the product arm.  The headline runs the full K steps + W warm-ups of the full batch; the other workloads
    run a few steps each so that the whole arm ends within a few minutes."""
    if args.only:
        emit(reference_record(args.only, args, args.steps, args.warmup))
        return
    rec = reference_record(HEADLINE[0], args, args.steps, args.warmup)
    rec["workloads"] = {}
    for w, prec in OTHERS:
        if prec != "bf16" and w == HEADLINE[0]:
            continue        # the CPU arm has one precision (fp32): resnet50_tf32 compares with the headline record
        rec["workloads"][w] = reference_record(w, args, min(args.steps, 5), 1)
    emit(rec)


_OUT = None


def emit(rec):
    out = _OUT or sys.stdout
    out.write(json.dumps(rec) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="hfr", choices=["hfr", "reference"])
    ap.add_argument("--only", "--workload", dest="only", default=os.environ.get("HFR_BENCH_WORKLOAD"),
                    choices=list(WORKLOADS), help="run one workload only (default: headline + all the others)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"], help="with --only")
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (queries for knn)")
    ap.add_argument("--gallery", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (ncu captures)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for ncu captures)")
    ap.add_argument("--layers", action="store_true", help="add per-layer event timings to the JSON line")
    args = ap.parse_args()
    # ONE JSON line on stdout: everything else that writes to file descriptor 1 (NCCL prints its version banner there,
    # from C) goes to stderr; the record is written through a private duplicate of the original stdout
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    def one(workload, precision, cpu=True):
        t0 = time.perf_counter()
        rec = (bench_knn(precision, args, rank, world, local) if workload == "knn"
               else bench_network(workload, precision, args, rank, world, local))
        if rank == 0:
            if world == 1 and cpu and not args.no_cpu_baseline:     # rank 0 at N=1 only
                if workload == "knn":
                    rec["cpu_baseline"] = cpu_knn(args.gallery, 5, 1)[0]
                else:
                    rec["cpu_baseline"] = cpu_network(workload, 1000, 1, args.batch, budget_s=10.0)[0]
            print(f"[bench] {workload} {precision}: {rec['value']} {rec['unit']} ({time.perf_counter() - t0:.1f} s)",
                  file=sys.stderr)
        return rec

    if args.only:
        res = one(args.only, args.precision)
    else:
        res = one(*HEADLINE)
        others = {}
        for w, prec in OTHERS:
            r = one(w, prec, cpu=not (w == HEADLINE[0]))
            if rank == 0:
                if w == HEADLINE[0] and "cpu_baseline" in res:
                    r["cpu_baseline"] = res["cpu_baseline"]      # same CPU path (fp32), not timed twice
                others[w if prec == "bf16" else f"{w}_{prec}"] = r
        if rank == 0:
            res["workloads"] = others
    if rank == 0:
        emit(res)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
