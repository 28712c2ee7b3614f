#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 1 2 4; do
  HFR_BENCH_E2E_DEPTH=$d timeout -k 5 300 python bench.py --workload mobilenet192 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mn_d$d.json 2> gpurun_out/bench_mn_d$d.err; grep "e2e" gpurun_out/bench_mn_d$d.err
done
HFR_BENCH_E2E_DEPTH=3 timeout -k 5 300 python bench.py --workload resnet50 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_d3.json 2> gpurun_out/bench_r50_d3.err; grep "e2e" gpurun_out/bench_r50_d3.err
