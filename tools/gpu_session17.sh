#!/bin/bash
# MMA-issue-loop fix: conv_window (stem + optional 64-ch 3x3) and gemm_tc with precomputed descriptors
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary17.txt; : > $S
timeout -k 5 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_17.log 2>&1; echo "pytest rc=$?" >> $S
for cfg in default noshift winconv winconv_noshift; do
  unset HFR_NO_SHIFTED HFR_WINDOW_CONV
  case $cfg in noshift) export HFR_NO_SHIFTED=1;; winconv) export HFR_WINDOW_CONV=1;; winconv_noshift) export HFR_WINDOW_CONV=1 HFR_NO_SHIFTED=1;; esac
  timeout -k 5 300 python bench.py --workload resnet50 --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_r50_$cfg.json 2> gpurun_out/bench_r50_$cfg.err; echo "bench $cfg rc=$?" >> $S
done
unset HFR_NO_SHIFTED HFR_WINDOW_CONV
HFR_WINDOW_CONV=1 timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q -x --tb=short -p no:cacheprovider -k "resnet" > gpurun_out/pytest_17b.log 2>&1; echo "pytest winconv rc=$?" >> $S
timeout -k 5 300 python bench.py --workload mobilenet192 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_mn.json 2> gpurun_out/bench_mn.err; echo "bench mn rc=$?" >> $S
timeout -k 5 300 python bench.py --workload knn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn.json 2> gpurun_out/bench_knn.err; echo "bench knn rc=$?" >> $S
cat $S; tail -5 gpurun_out/pytest_17.log; tail -5 gpurun_out/pytest_17b.log
python tools/show_bench.py gpurun_out/bench_r50_*.json gpurun_out/bench_mn.json gpurun_out/bench_knn.json
