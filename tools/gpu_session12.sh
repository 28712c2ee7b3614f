#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for W in agegender224 mobilenet192 resnet50; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline --layers > gpurun_out/layers_$W.json 2> gpurun_out/layers_$W.err; echo "bench $W rc=$?"
done
