#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary7.txt; : > $S
timeout -k 5 200 python tools/gpu_diag.py win stem > gpurun_out/diag_win.log 2>&1; echo "diag win rc=$?" >> $S
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "window or stem" > gpurun_out/pytest_k6.log 2>&1; echo "pytest window/stem rc=$?" >> $S
for W in resnet50 mobilenet192 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
done
cat $S; grep -c "mismatches=0/" gpurun_out/diag_win.log; grep -v "mismatches=0/" gpurun_out/diag_win.log | tail; tail -3 gpurun_out/pytest_k6.log
