#!/bin/bash
# tile-width experiment: per-layer times with BLOCK_N capped at 256 / 128 / 64
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cap in 256 128 64; do
  HFR_BLOCK_N_MAX=$cap timeout -k 5 300 python bench.py --workload resnet50 --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_r50_bn$cap.json 2> gpurun_out/bench_r50_bn$cap.err; echo "bench bn$cap rc=$?"
done
python tools/show_bench.py gpurun_out/bench_r50_bn*.json
