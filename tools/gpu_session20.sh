#!/bin/bash
# unrolled window-conv MMA issue + CTA-pair policy: full GPU test suite, then benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary20.txt; : > $S
HFR_PAIR=1 timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q -x --tb=short -p no:cacheprovider -k "gemm or implicit" > gpurun_out/pytest_20a.log 2>&1; echo "pytest pair-forced gemm/conv rc=$?" >> $S
timeout -k 5 1200 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_20.log 2>&1; echo "pytest all rc=$?" >> $S
for cfg in default winconv; do
  unset HFR_WINDOW_CONV; [ $cfg = winconv ] && export HFR_WINDOW_CONV=1
  timeout -k 5 300 python bench.py --workload resnet50 --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_r50_$cfg.json 2> gpurun_out/bench_r50_$cfg.err; echo "bench $cfg rc=$?" >> $S
done
HFR_WINDOW_CONV=1 timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q -x --tb=short -p no:cacheprovider -k "resnet" > gpurun_out/pytest_20b.log 2>&1; echo "pytest winconv rc=$?" >> $S
unset HFR_WINDOW_CONV
timeout -k 5 300 python bench.py --workload mobilenet192 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_mn.json 2> gpurun_out/bench_mn.err; echo "bench mn rc=$?" >> $S
timeout -k 5 300 python bench.py --workload agegender224 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_ag.json 2> gpurun_out/bench_ag.err; echo "bench ag rc=$?" >> $S
cat $S; tail -4 gpurun_out/pytest_20a.log; tail -4 gpurun_out/pytest_20.log; tail -3 gpurun_out/pytest_20b.log
python tools/show_bench.py gpurun_out/bench_r50_default.json gpurun_out/bench_r50_winconv.json gpurun_out/bench_mn.json gpurun_out/bench_ag.json
