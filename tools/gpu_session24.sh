#!/bin/bash
# pipelined host-buffer API: tests, then benches (e2e is the number that should move)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary24.txt; : > $S
timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q -x --tb=short -p no:cacheprovider -k "host or stream" > gpurun_out/pytest_24a.log 2>&1; echo "pytest host rc=$?" >> $S
for w in resnet50 mobilenet192 agegender224; do
  timeout -k 5 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?" >> $S
done
cat $S; tail -15 gpurun_out/pytest_24a.log
python tools/show_bench.py gpurun_out/bench_resnet50.json gpurun_out/bench_mobilenet192.json gpurun_out/bench_agegender224.json | grep -v "^     "
