#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary14.txt; : > $S
timeout -k 5 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "dwconv or parity or layerwise or batch_sizes" > gpurun_out/pytest_14.log 2>&1; echo "pytest rc=$?" >> $S
for W in mobilenet192 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline --layers > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
done
cat $S; tail -4 gpurun_out/pytest_14.log
