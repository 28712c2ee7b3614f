#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary8.txt; : > $S
timeout -k 5 300 python tools/gpu_diag.py dwpw > gpurun_out/diag_dwpw.log 2>&1; echo "diag dwpw rc=$?" >> $S
if grep -q "dwpw 5x7x7x512->1024: mismatches=0/" gpurun_out/diag_dwpw.log; then
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "fused or age_post or l2" > gpurun_out/pytest_k8.log 2>&1; echo "pytest fused rc=$?" >> $S
timeout -k 5 900 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_m8.log 2>&1; echo "pytest model rc=$?" >> $S
for W in mobilenet192 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
  HFR_NO_FUSE=1 timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_nofuse.json 2> gpurun_out/bench_${W}_nofuse.err; echo "bench $W nofuse rc=$?" >> $S
done
fi
cat $S; cat gpurun_out/diag_dwpw.log | tail -30; tail -5 gpurun_out/pytest_k8.log; grep -E "cosine|passed|failed|Error" gpurun_out/pytest_m8.log | cut -c1-200
