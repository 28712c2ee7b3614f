#!/bin/bash
# CTA-pair GEMM: correctness first (short timeout: a protocol bug shows up as a hang), then A/B benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary19.txt; : > $S
timeout -k 5 240 python -m pytest tests/test_kernels_gpu.py -q -x --tb=short -p no:cacheprovider -k "gemm or implicit" > gpurun_out/pytest_19a.log 2>&1; rc=$?; echo "pytest gemm/conv rc=$rc" >> $S
if [ $rc -eq 0 ]; then
  timeout -k 5 600 python -m pytest tests/test_model_gpu.py tests/test_knn_gpu.py -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_19b.log 2>&1; echo "pytest model rc=$?" >> $S
  for cfg in pair nopair; do
    unset HFR_NO_PAIR; [ $cfg = nopair ] && export HFR_NO_PAIR=1
    timeout -k 5 300 python bench.py --workload resnet50 --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_r50_$cfg.json 2> gpurun_out/bench_r50_$cfg.err; echo "bench $cfg rc=$?" >> $S
  done
  unset HFR_NO_PAIR
  timeout -k 5 300 python bench.py --workload mobilenet192 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_mn.json 2> gpurun_out/bench_mn.err; echo "bench mn rc=$?" >> $S
  timeout -k 5 300 python bench.py --workload agegender224 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_ag.json 2> gpurun_out/bench_ag.err; echo "bench ag rc=$?" >> $S
fi
cat $S; tail -15 gpurun_out/pytest_19a.log; tail -5 gpurun_out/pytest_19b.log
python tools/show_bench.py gpurun_out/bench_r50_pair.json gpurun_out/bench_r50_nopair.json gpurun_out/bench_mn.json gpurun_out/bench_ag.json
