#!/bin/bash
# round-1 final evidence: full GPU test suite, default benches (with the CPU baseline leg), reference arm, ncu launch
# lists + full captures of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary26.txt; : > $S
timeout -k 5 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_26.log 2>&1; echo "pytest all rc=$?" >> $S
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_26.log 2>&1; echo "smoke rc=$?" >> $S
timeout -k 5 600 python bench.py > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.err; echo "bench default rc=$?" >> $S
for W in mobilenet192 agegender224 knn; do
  timeout -k 5 600 python bench.py --workload $W > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
done
timeout -k 5 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?" >> $S
bash tools/gpu_profile_final.sh r1_resnet50 resnet50 224 56 "gemm_tc|conv_window" >> $S 2>&1
bash tools/gpu_profile_final.sh r1_mobilenet192 mobilenet192 290 29 "gemm_tc|dwconv3x3" >> $S 2>&1
cat $S; tail -4 gpurun_out/pytest_26.log; tail -3 gpurun_out/smoke_26.log
python tools/show_bench.py gpurun_out/bench_resnet50.json gpurun_out/bench_mobilenet192.json gpurun_out/bench_agegender224.json gpurun_out/bench_knn.json
cat gpurun_out/bench_reference.json
