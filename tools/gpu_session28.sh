#!/bin/bash
# narrower 1x1 tiles for the small-M MobileNet layers?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cap in 128 64; do
  HFR_BLOCK_N_MAX=$cap timeout -k 5 200 python bench.py --workload mobilenet192 --steps 30 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_mn_bn$cap.json 2> gpurun_out/bench_mn_bn$cap.err; echo "mn cap $cap rc=$?"
done
python tools/show_bench.py gpurun_out/bench_mn_bn128.json gpurun_out/bench_mn_bn64.json | grep -v "^     \|roofline"
