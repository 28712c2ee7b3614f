#!/usr/bin/env python
"""bench.py - headline benchmark of the hot path (contract: see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mobilenet192|agegender224|resnet50|knn]
  python bench.py --impl reference ...      # the CPU restatement of the reference's path on the host cores

One "step" = one pass of the hot path over one batch of synthetic input.  Prints ONE JSON line on rank 0.
  value   whole-job throughput, inputs resident in HBM, CUDA-graph replay, CUDA-event timed, max over ranks
  e2e     the same metric through the host-buffer C-ABI call (pinned host input -> H2D -> run -> D2H of the result)
  roofline  dominant kernel class: algorithmic bytes|flops per launch / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  oracle (torch-CPU port; sklearn for 1-NN = the reference's real dependency) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor": 1400.0, "source": "fallback"}


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


# ------------------------------------------------------------------------------------------------------ networks
def synth_images(batch, size, seed):
    return np.random.RandomState(seed).randint(0, 256, (batch, size, size, 3)).astype(np.uint8)


def model_spec(workload, precision):
    if workload == "mobilenet192":
        return dict(path=PB, input="input_1:0", outputs=["global_pooling/Mean:0"], hw=192, bgr=True, imagenet=True)
    if workload == "agegender224":
        return dict(path=PB, input="input_1:0",
                    outputs=["age_pred/Softmax:0", "gender_pred/Sigmoid:0", "global_pooling/Mean:0"], hw=0, bgr=True,
                    imagenet=True)
    if workload == "resnet50":
        from hse_facerec_tf_b200.synth import ensure_resnet50_pb
        return dict(path=ensure_resnet50_pb(), input="input:0", outputs=["pool5_7x7_s1:0"], hw=0, bgr=True,
                    imagenet=False)
    raise ValueError(workload)


def plan_work(plan, es):
    """Algorithmic work per image from the compiled plan: per layer (flops, bytes moved in+out+weights-once-ignored)."""
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
        elif kind == "fc":
            flops = 2.0 * L["cin"] * L["cout"]
        else:
            flops = 0.0
        in_b = hin * win * L["cin"] * (1 if kind == "stem" else es)
        if kind == "subsample":
            in_b = ho * wo * L["cin"] * es      # only every stride-th pixel is read
        out_b = ho * wo * L["cout"] * (4 if kind in ("gap", "fc") else es)
        if kind == "fc":
            in_b, out_b = L["cin"] * 4, L["cout"] * 4
        if L["in2"] >= 0:
            in_b += ho * wo * L["cout"] * es
        rows.append(dict(kind=kind, name=L["name"], flops=flops, bytes=float(in_b + out_b)))
    return rows


def bench_network(args, rank, world, dev):
    import torch
    import hse_facerec_tf_b200 as hfr
    desc, batch, size, metric, unit = WORKLOADS[args.workload]
    batch = args.batch or batch
    spec = model_spec(args.workload, args.precision)
    model = hfr.HfrModel(spec["path"], spec["input"], spec["outputs"], input_hw=spec["hw"], device=f"cuda:{dev}",
                         precision=args.precision)
    es = 2 if args.precision == "bf16" else 4
    size = model.h
    # rotate over enough distinct input batches that the inputs alone exceed the 126 MB L2
    in_bytes = batch * size * size * 3
    nrot = max(2, min(32, -(-140_000_000 // in_bytes)))
    xs = [torch.from_numpy(synth_images(batch, size, 1000 * rank + i)).to(f"cuda:{dev}") for i in range(nrot)]
    outs = [[torch.empty((batch, d), dtype=torch.float32, device=f"cuda:{dev}") for d in model.out_dims]
            for _ in range(nrot)]
    stream = torch.cuda.Stream(device=dev)
    kw = dict(convert2BGR=spec["bgr"], imageNetUtilsMean=spec["imagenet"])

    def step(i, graph=not args.no_graph):
        model.forward(xs[i % nrot], graph=graph, outs=outs[i % nrot], **kw)

    import torch.distributed as dist
    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 3) + nrot):          # warm-up (also captures one graph per input buffer)
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
        while time.perf_counter() - t_pre < 0.5:
            for i in range(8):
                step(i)
            stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            step(i)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        clocks = sampler.stop()

        # ---- e2e: pinned host batch -> H2D -> run -> D2H of the result, through the host-buffer C-ABI call
        # The streaming form of the call (hfr_model_submit_host / hfr_model_wait_host, two batches in flight): every
        # step's input still travels pinned host -> device and its result device -> pinned host inside the timed region,
        # but step i+1's upload and step i-1's download overlap step i's compute.
        depth = int(os.environ.get("HFR_BENCH_E2E_DEPTH", "2"))
        hx = [torch.from_numpy(synth_images(batch, size, 5000 + 1000 * rank + i)).pin_memory() for i in range(depth)]
        houts = [[torch.empty((batch, d), dtype=torch.float32).pin_memory().numpy() for d in model.out_dims]
                 for _ in range(depth)]

        hxn = [h.numpy() for h in hx]
        host_s = [0.0, 0.0]  # seconds the host spent inside submit / wait (reported on stderr)

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
        e2e_run(args.steps)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        print(f"[e2e] depth {depth}: {e2e_s / args.steps * 1e3:.3f} ms/step; host in submit {host_s[0] / args.steps * 1e3:.3f} ms, "
              f"in wait {host_s[1] / args.steps * 1e3:.3f} ms per step", file=sys.stderr)

        # ---- per-kernel durations (CUDA events around every layer launch, eager, same stream)
        model.layer_timing(True)
        for i in range(min(args.steps, 20)):
            step(i, graph=False)
        lms, lsteps = model.layer_times()
        model.layer_timing(False)

    if world > 1:
        t = torch.tensor([ms, e2e_s], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        return None

    pk = peaks()
    work = plan_work(model.plan(), es)
    classes = {}
    # a depthwise layer followed by a pointwise layer with an empty interval ran as ONE fused dwpw_kernel launch:
    # book the pair as class "dwpw" (flops of both, bytes = depthwise input + pointwise output)
    per_step = [t / max(lsteps, 1) for t in lms]
    merged = []
    i = 0
    while i < len(work):
        w, t = dict(work[i]), per_step[i]
        if w["kind"] == "dw" and i + 1 < len(work) and work[i + 1]["kind"] == "pw" and per_step[i + 1] < 2e-4:
            L0, L1 = model.plan()["layers"][i], model.plan()["layers"][i + 1]
            w = dict(kind="dwpw", name=work[i + 1]["name"], flops=work[i]["flops"] + work[i + 1]["flops"],
                     bytes=float((L0["hw_in"][0] * L0["hw_in"][1] * L0["cin"] + L1["hw_out"][0] * L1["hw_out"][1] * L1["cout"]) * es))
            i += 1
        if w["kind"] == "subsample" and t < 5e-3:
            i += 1      # bypassed: its consumers fetch the strided pixels themselves (im2col map), nothing was launched
            continue
        merged.append((w, t))
        i += 1
    for w, t in merged:
        t = t * max(lsteps, 1)
        c = classes.setdefault(w["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        c["ms"] += t / max(lsteps, 1)
        c["flops"] += w["flops"] * batch
        c["bytes"] += w["bytes"] * batch
        c["launches"] += 1
    kernels = {}
    for kname, c in classes.items():
        if c["ms"] <= 0:
            continue
        # each class is compared with BOTH roofs (algorithmic flops vs measured bf16 peak, algorithmic bytes vs measured
        # HBM copy bandwidth); the binding roof is the one it sits closer to
        tf = c["flops"] / (c["ms"] * 1e-3) / 1e12
        gb = c["bytes"] / (c["ms"] * 1e-3) / 1e9
        frac_t, frac_h = tf / pk["tensor"], gb / pk["hbm"]
        tensor_bound = kname in ("pw", "conv", "dwpw") and frac_t >= frac_h
        kernels[kname] = dict(ms_per_step=round(c["ms"], 5), launches=c["launches"],
                              bound="tensor" if tensor_bound else "hbm",
                              achieved=round(tf if tensor_bound else gb, 2),
                              unit="TFLOP/s" if tensor_bound else "GB/s",
                              frac=round(frac_t if tensor_bound else frac_h, 4),
                              tflops=round(tf, 2), hbm_gbs=round(gb, 1))
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    kd = kernels[dom]
    roof = dict(kernel={"dwpw": "dwpw_kernel (fused depthwise 3x3 + pointwise 1x1)", "pw": "gemm_tc_kernel (1x1 conv)", "conv": "gemm_tc_kernel (im2col) / conv_window_kernel", "dw": "dwconv3x3_pipe_kernel",
                        "stem": "stem_s2d_kernel + conv_window_kernel" if args.precision == "bf16" else "stem_conv_kernel"}.get(dom, dom),
                bound=kd["bound"], achieved=kd["achieved"], peak=pk["tensor"] if kd["bound"] == "tensor" else pk["hbm"],
                unit=kd["unit"], frac=kd["frac"], traffic=None, peak_source=pk["source"] + " (sustained)",
                per_launch_ms=round(kd["ms_per_step"] / kd["launches"], 5))
    # DRAM traffic per launch of the dominant class, from the committed ncu launch list of this workload (if present)
    tpath = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(tpath):
        tr = json.load(open(tpath)).get("classes", {}).get(dom)
        if tr and batch == WORKLOADS[args.workload][1] and args.precision == "bf16":
            roof["traffic"] = tr["dram_bytes_per_launch"]
            roof["traffic_source"] = f"profiles/traffic_{args.workload}.json (ncu, share of step {tr['share']})"
            roof["algorithmic_bytes_per_launch"] = round(classes[dom]["bytes"] / classes[dom]["launches"])
    total = batch * world * args.steps
    value = total / (ms * 1e-3)
    out_bytes = sum(model.out_dims) * 4 * batch
    return dict(metric=metric, value=round(value, 1), unit=unit, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=round(ms / args.steps, 4), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=args.precision, data="synthetic",
                config=dict(workload=args.workload, description=desc, per_gpu_batch=batch, global_batch=batch * world,
                            input=f"{size}x{size}x3 uint8", parallelism=f"dp{world}",
                            l2_policy=f"inputs rotate over {nrot} distinct batches ({nrot * in_bytes / 1e6:.0f} MB > L2); "
                                      "activations (>= 0.4 GB/step written) exceed L2",
                            cuda_graph=not args.no_graph),
                e2e=dict(value=round(total / e2e_s, 1), unit=unit, h2d_bytes_per_step=in_bytes,
                         d2h_bytes_per_step=out_bytes),
                gpu_launches=int(launches_per_step * args.steps), launches_per_step=int(launches_per_step),
                clocks=clocks, roofline=roof, kernels=kernels,
                **({"layers": [dict(kind=w["kind"], name=w["name"], us=round(t * 1e3, 2), gflop=round(w["flops"] * batch / 1e9, 2),
                                    mb=round(w["bytes"] * batch / 1e6, 1)) for w, t in merged]} if args.layers else {}))


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


def bench_knn(args, rank, world, dev):
    import torch
    import torch.distributed as dist
    import hse_facerec_tf_b200 as hfr
    desc, nq, d, metric, unit = WORKLOADS["knn"]
    n = args.gallery
    nq = args.batch or nq
    gal, q = knn_data(n, nq, d, rank, world)
    if world > 1:
        dist.broadcast_object_list([None], src=0)  # cheap sync before big allocations
        qt = q.cuda(dev)
        dist.broadcast(qt, src=0)                  # rank 0's queries (with planted neighbours) everywhere
        q = qt.cpu()
    clf = hfr.KNeighborsClassifier(1, 2, device=f"cuda:{dev}", precision=args.precision, sharded=world > 1)
    y_local = np.arange(rank * (n // world), (rank + 1) * (n // world)) % 10000
    clf.fit(gal.cuda(dev), y_local)
    qd = q.cuda(dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            clf.kneighbors(qd, return_distance=False)
        l0 = hfr.launch_count()
        clf.kneighbors(qd, return_distance=False)
        launches_per_step = hfr.launch_count() - l0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        sampler = ClockSampler(dev)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            clf.kneighbors(qd, return_distance=False)   # includes the D2H of the 100k indices (0.8 MB)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        # e2e: pinned host queries in, indices out
        hq = q.pin_memory()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ind = clf.kneighbors(hq.numpy(), return_distance=False)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=f"cuda:{dev}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        return None
    m = min(nq, n // world)
    planted_ok = float((ind[:m, 0] == np.arange(m)).mean())
    pk = peaks()
    flops = 2.0 * nq * (n // world) * d
    per_step_ms = ms / args.steps
    ach = flops / (per_step_ms * 1e-3) / 1e12
    return dict(metric=metric, value=round(nq * args.steps / (ms * 1e-3), 1), unit=unit, n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=round(per_step_ms, 3), higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype=args.precision, data="synthetic",
                config=dict(workload="knn", description=desc, queries=nq, gallery=n, dim=d, shards=world,
                            parallelism=f"gallery row-sharded x{world}",
                            l2_policy=f"gallery shard {(n // world) * d * (2 if args.precision == 'bf16' else 4) / 1e9:.1f} GB >> L2",
                            planted_neighbour_recall=planted_ok),
                e2e=dict(value=round(nq * args.steps / e2e_s, 1), unit=unit, h2d_bytes_per_step=nq * d * 4,
                         d2h_bytes_per_step=nq * 8),
                gpu_launches=int(launches_per_step * args.steps), launches_per_step=int(launches_per_step), clocks=clocks,
                roofline=dict(kernel="gemm_tc_kernel (EPI_KNN) incl. prep+finalize in the step time", bound="tensor",
                              achieved=round(ach, 2), peak=pk["tensor"], unit="TFLOP/s", frac=round(ach / pk["tensor"], 4),
                              traffic=None, peak_source=pk["source"] + " (sustained)"))


# ------------------------------------------------------------------------------------------------------ CPU legs
def cpu_network(args, budget_s=12.0):
    """Oracle port (torch-CPU restatement of the frozen graph) on all host cores; bounded sample."""
    import torch
    from oracle.tfnet import GraphOracle, preprocess_rgb_u8
    desc, batch, size, metric, unit = WORKLOADS[args.workload]
    spec = model_spec(args.workload, "tf32")
    torch.set_num_threads(os.cpu_count())
    g = GraphOracle(spec["path"])
    size = spec["hw"] or (g.placeholder_shape(spec["input"]) or [0, 224])[1]
    sample = min(args.batch or batch, 16)
    x = preprocess_rgb_u8(synth_images(sample, size, 0), True, spec["imagenet"])
    g.run([o for o in spec["outputs"]], {spec["input"]: x})      # warm-up
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        g.run([o for o in spec["outputs"]], {spec["input"]: x})
        n += sample
    dt = time.perf_counter() - t0
    return dict(value=round(n / dt, 2), unit=unit, cores=torch.get_num_threads(), kind="port",
                sample=f"{n} images in batches of {sample} at {size}x{size} through oracle/tfnet.py (torch-CPU fp32, "
                       f"{dt:.1f} s); TensorFlow itself is not installable here")


def cpu_knn(args, budget_q=1500):
    """The reference's real dependency: sklearn KNeighborsClassifier(n_neighbors=1, p=2).kneighbors; query subsample."""
    from sklearn import neighbors
    desc, nq, d, metric, unit = WORKLOADS["knn"]
    n = min(args.gallery, 1_000_000)
    gal, q = knn_data(n, budget_q, d, 0, 1)
    nn = neighbors.KNeighborsClassifier(n_neighbors=1, p=2).fit(gal.numpy(), np.arange(n) % 10000)
    t0 = time.perf_counter()
    nn.kneighbors(q.numpy(), return_distance=False)
    dt = time.perf_counter() - t0
    return dict(value=round(budget_q / dt, 2), unit=unit, cores=os.cpu_count(), kind="reference",
                sample=f"scikit-learn kneighbors on {budget_q} of the {nq} queries vs the full {n} x {d} gallery ({dt:.1f} s), "
                       "extrapolated linearly")


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores, same config/metric/unit."""
    desc, batch, size, metric, unit = WORKLOADS[args.workload]
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        cb = cpu_knn(args, 600) if args.workload == "knn" else cpu_network(args, 6.0)
        vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    print(json.dumps(dict(impl="reference", metric=metric, value=round(v, 2), unit=unit, n_gpus=args.gpus,
                          steps=len(vals), warmup=1,
                          # time this arm needs for one step of the GPU arm's size (batch images / nq queries),
                          # extrapolated from the bounded samples
                          ms_per_step=round(batch / v * 1e3, 1), higher_is_better=True,
                          scaling="strong" if args.workload == "knn" else "weak", vs_baseline=None, dtype="f32",
                          data="synthetic", config=dict(workload=args.workload, description=desc), cpu_baseline=cb,
                          e2e=dict(value=round(v, 2), unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="hfr", choices=["hfr", "reference"])
    ap.add_argument("--workload", default=os.environ.get("HFR_BENCH_WORKLOAD", "resnet50"), choices=list(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (queries for knn)")
    ap.add_argument("--gallery", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for ncu captures)")
    ap.add_argument("--layers", action="store_true", help="add per-layer event timings to the JSON line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    if args.workload == "knn" and args.steps > 10:
        args.steps = 5
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    res = bench_knn(args, rank, world, local) if args.workload == "knn" else bench_network(args, rank, world, local)
    if rank == 0:
        if not args.no_cpu_baseline:
            res["cpu_baseline"] = cpu_knn(args) if args.workload == "knn" else cpu_network(args)
        print(json.dumps(res))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
