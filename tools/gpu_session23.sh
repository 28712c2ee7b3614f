#!/bin/bash
# pipelined persistent depthwise kernel: tests + MobileNet benches, against the one-shot kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary23.txt; : > $S
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q -x --tb=short -p no:cacheprovider -k "dwconv" > gpurun_out/pytest_23a.log 2>&1; rc=$?; echo "pytest dw rc=$rc" >> $S
if [ $rc -eq 0 ]; then
timeout -k 5 900 python -m pytest tests/test_model_gpu.py -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_23b.log 2>&1; echo "pytest model rc=$?" >> $S
for cfg in pipe oneshot; do
  unset HFR_DW_ONESHOT; [ $cfg = oneshot ] && export HFR_DW_ONESHOT=1
  timeout -k 5 300 python bench.py --workload mobilenet192 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_mn_$cfg.json 2> gpurun_out/bench_mn_$cfg.err; echo "bench mn $cfg rc=$?" >> $S
  timeout -k 5 300 python bench.py --workload agegender224 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_ag_$cfg.json 2> gpurun_out/bench_ag_$cfg.err; echo "bench ag $cfg rc=$?" >> $S
done
fi
cat $S; tail -5 gpurun_out/pytest_23a.log; tail -5 gpurun_out/pytest_23b.log
python tools/show_bench.py gpurun_out/bench_mn_pipe.json gpurun_out/bench_mn_oneshot.json gpurun_out/bench_ag_pipe.json gpurun_out/bench_ag_oneshot.json
