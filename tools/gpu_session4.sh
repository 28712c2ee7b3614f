#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary4.txt; : > $S
timeout -k 5 200 python tools/gpu_diag.py stem > gpurun_out/diag_stem.log 2>&1; echo "diag stem rc=$?" >> $S
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "stem" > gpurun_out/pytest_stem.log 2>&1; echo "pytest stem rc=$?" >> $S
timeout -k 5 900 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?" >> $S
for W in mobilenet192 resnet50 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
done
cat $S; cat gpurun_out/diag_stem.log | tail -30; tail -15 gpurun_out/pytest_stem.log; grep -E "cosine|passed|failed" gpurun_out/pytest_model.log | cut -c1-200
