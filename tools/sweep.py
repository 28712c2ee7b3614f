"""BASELINE config 5: end-to-end extract -> L2-normalise -> identify throughput, batch 32..4096 (GLOBAL batch, split over
the ranks), MobileNet-192 and ResNet-50, against a 100k-row gallery row-sharded over the ranks.  The flow is the
reference's facerec_test.py:390-442 with batches instead of single files and both halves on the GPUs:

  rank r: its slice of the uint8 crops (pinned host) -> H2D -> embeddings (L2 norm fused into the forward call)
          -> all-gather of the [B/P, D] embedding blocks over NCCL        (KNeighborsClassifier.kneighbors(local_queries=True))
          -> distance GEMM + top-1 against its gallery shard -> ONE all-gather of the packed (dist2, index) records -> merge
          -> label lookup, indices back on the host

  python tools/sweep.py --out gpurun_out/sweep_n1.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 tools/sweep.py --out ...
Timing: CUDA events on the compute stream around K steps after W warm-ups, barrier on both sides, max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload specs, clock sampler)
import hse_facerec_tf_b200 as hfr  # noqa: E402
from hse_facerec_tf_b200.parallel import shard_rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--batches", default="32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--nets", default="mobilenet192,resnet50")
    ap.add_argument("--gallery", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    rows = []
    sampler = bench.ClockSampler(local)
    sampler.start()
    for net in args.nets.split(","):
        spec = bench.model_spec(net)
        tfi = hfr.TensorFlowInference(spec["path"], spec["input"], spec["outputs"][0], convert2BGR=spec["bgr"],
                                      imageNetUtilsMean=spec["imagenet"], device=dev, precision=args.precision,
                                      input_hw=spec["hw"])
        d = tfi.model.out_dims[0]
        a, b = shard_rows(args.gallery, world, rank)
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        gal = torch.randn(b - a, d, generator=g).abs()          # embeddings are post-ReLU averages: non-negative
        gal = (gal / gal.norm(dim=1, keepdim=True)).to(dev)
        clf = hfr.KNeighborsClassifier(1, 2, device=dev, precision=args.precision, sharded=world > 1)
        clf.fit(gal, np.arange(a, b) % 5000)
        stream = torch.cuda.Stream(device=local)
        for B in [int(x) for x in args.batches.split(",")]:
            if B < world:
                continue
            qa, qb = shard_rows(B, world, rank)
            hx = torch.from_numpy(bench.synth_images(qb - qa, tfi.h, 7 * B + rank)).pin_memory()
            dx = torch.empty_like(hx, device=dev)

            def step():
                dx.copy_(hx, non_blocking=True)
                emb = tfi.extract_batch(dx, l2norm=True, graph=True)
                ind = clf.kneighbors(emb, return_distance=False, local_queries=world > 1, total_queries=B)   # ends with the D2H of the indices
                return clf._labels[ind[:, 0]]

            with torch.cuda.stream(stream):
                for _ in range(args.warmup):
                    labels = step()
                # split of one step, measured once (events around the two halves)
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record(stream)
                dx.copy_(hx, non_blocking=True)
                emb = tfi.extract_batch(dx, l2norm=True, graph=True)
                e[1].record(stream)
                clf.kneighbors(emb, return_distance=False, local_queries=world > 1, total_queries=B)
                e[2].record(stream)
                stream.synchronize()
                t_ext, t_knn = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize(local)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                e0.record(stream)
                for _ in range(args.steps):
                    labels = step()
                e1.record(stream)
                stream.synchronize()
                wall = time.perf_counter() - t0
                ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms, wall * 1e3, t_ext, t_knn], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms, wall, t_ext, t_knn = float(t[0]), float(t[1]) / 1e3, float(t[2]), float(t[3])
            assert len(labels) == B
            if rank == 0:
                rows.append(dict(net=net, global_batch=B, n_gpus=world, per_gpu_batch=qb - qa, gallery=args.gallery, dim=d,
                                 images_per_s=round(B * args.steps / max(ms * 1e-3, wall), 1),
                                 ms_per_step=round(max(ms, wall * 1e3) / args.steps, 4),
                                 extract_ms=round(t_ext, 4), identify_ms=round(t_knn, 4), precision=args.precision))
                print(rows[-1], flush=True)
        del clf, gal
        tfi.close_session()
        torch.cuda.empty_cache()
    clocks = sampler.stop()
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(dict(config="BASELINE configs[4]: end-to-end extract+identify sweep", n_gpus=world, steps=args.steps,
                       warmup=args.warmup, clocks=clocks, rows=rows,
                       note="images_per_s = global batch x steps / max(device time, host wall time) over the ranks; every "
                            "step includes the H2D of the crops, both NCCL all-gathers and the D2H of the indices"),
                  open(args.out, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
