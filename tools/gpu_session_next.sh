#!/bin/bash
# First GPU session of the next round: validate and measure what round 1 prepared without GPU time left.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session_next.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary_next.txt; : > $S
# 1. the opt-in tests of the experimental paths (lanes, tf32 tensor-core stem)
HFR_TEST_EXPERIMENTAL=1 timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -k experimental \
  > gpurun_out/pytest_next_experimental.log 2>&1; echo "pytest experimental rc=$?" >> $S
# 2. lanes A/B on the three network workloads
for W in resnet50 mobilenet192 agegender224; do
  for L in 1 2 3; do
    HFR_LANES=$L timeout -k 5 300 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_${W}_lanes$L.json 2> gpurun_out/bench_${W}_lanes$L.err; echo "bench $W lanes=$L rc=$?" >> $S
  done
done
# 3. tf32 mode with and without the tensor-core stem
for T in 0 1; do
  for W in resnet50 mobilenet192; do
    if [ $T = 1 ]; then export HFR_TF32_TC_STEM=1; else unset HFR_TF32_TC_STEM; fi
    timeout -k 5 300 python bench.py --workload $W --precision tf32 --steps 20 --warmup 3 --layers --no-cpu-baseline \
      > gpurun_out/bench_${W}_tf32_tcstem$T.json 2> gpurun_out/bench_${W}_tf32_tcstem$T.err; echo "bench $W tf32 tcstem=$T rc=$?" >> $S
  done
done
unset HFR_TF32_TC_STEM
cat $S; tail -15 gpurun_out/pytest_next_experimental.log
python tools/show_bench.py gpurun_out/bench_*_lanes*.json gpurun_out/bench_*_tf32_tcstem*.json | grep -v "^     \|roofline"
