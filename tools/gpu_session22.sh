#!/bin/bash
# L2-residency experiment: whole-network sub-batching with the current kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sb in 0 128 64 32; do
  HFR_SUB_BATCH=$sb timeout -k 5 300 python bench.py --workload resnet50 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_sb$sb.json 2> gpurun_out/bench_r50_sb$sb.err; echo "bench sb$sb rc=$?"
done
timeout -k 5 300 python bench.py --workload resnet50 --batch 512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_b512.json 2> gpurun_out/bench_r50_b512.err; echo "bench b512 rc=$?"
timeout -k 5 300 python bench.py --workload resnet50 --batch 128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_b128.json 2> gpurun_out/bench_r50_b128.err; echo "bench b128 rc=$?"
timeout -k 5 300 python bench.py --workload mobilenet192 --batch 256 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mn_b256.json 2> gpurun_out/bench_mn_b256.err; echo "bench mn b256 rc=$?"
python tools/show_bench.py gpurun_out/bench_r50_sb*.json gpurun_out/bench_r50_b512.json gpurun_out/bench_r50_b128.json gpurun_out/bench_mn_b256.json 2>&1 | grep -v "^     \|roofline"
