#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary11.txt; : > $S
timeout -k 5 200 python tools/gpu_diag.py stem > gpurun_out/diag_stem.log 2>&1; echo "diag stem rc=$?" >> $S
timeout -k 5 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "stem or end_to_end or empty or argument_errors or h5_on_gpu or parity_224" > gpurun_out/pytest_11.log 2>&1; echo "pytest rc=$?" >> $S
for W in resnet50 mobilenet192 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
done
HFR_NO_SHIFTED=1 timeout -k 5 600 python bench.py --workload resnet50 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_resnet50_noshift.json 2> gpurun_out/bench_resnet50_noshift.err; echo "bench resnet50 noshift rc=$?" >> $S
cat $S; tail -6 gpurun_out/diag_stem.log; tail -4 gpurun_out/pytest_11.log
