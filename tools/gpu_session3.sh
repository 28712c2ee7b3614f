#!/bin/bash
# Session 3: re-validate everything after the epilogue rewrite, then bench all workloads.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary3.txt; : > $S
timeout -k 5 200 python tools/gpu_diag.py gemm conv knn > gpurun_out/diag3.log 2>&1; echo "diag rc=$?" >> $S
if grep -q "gemm prec=2 1000x512x96: nan=0 mismatches=0" gpurun_out/diag3.log; then
timeout -k 5 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$?" >> $S
for W in mobilenet192 resnet50 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
done
timeout -k 5 500 python bench.py --workload knn --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn.json 2> gpurun_out/bench_knn.err; echo "bench knn rc=$?" >> $S
fi
cat $S; tail -15 gpurun_out/diag3.log; tail -5 gpurun_out/pytest_all.log
