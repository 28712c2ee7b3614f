#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary6.txt; : > $S
timeout -k 5 200 python tools/gpu_diag.py win stem > gpurun_out/diag_win.log 2>&1; echo "diag win rc=$?" >> $S
if grep -q "win 2x56x56x64->64 k3: mismatches=0/" gpurun_out/diag_win.log; then
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "window or stem" > gpurun_out/pytest_k6.log 2>&1; echo "pytest window/stem rc=$?" >> $S
timeout -k 5 900 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_m6.log 2>&1; echo "pytest model rc=$?" >> $S
for W in resnet50 mobilenet192 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
done
else
HFR_NO_WINDOW=1 timeout -k 5 300 python tools/gpu_diag.py stem > gpurun_out/diag_stem_nowin.log 2>&1
fi
cat $S; cat gpurun_out/diag_win.log | tail -40; tail -4 gpurun_out/pytest_k6.log; grep -E "cosine|passed|failed" gpurun_out/pytest_m6.log | cut -c1-200
