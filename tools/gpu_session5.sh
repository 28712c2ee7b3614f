#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary5.txt; : > $S
timeout -k 5 600 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -x -k "gemm or conv2d" > gpurun_out/pytest_k5.log 2>&1; echo "pytest gemm/conv rc=$?" >> $S
if [ $? = 0 ]; then
timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -s -k resnet50 > gpurun_out/pytest_m5.log 2>&1; echo "pytest resnet rc=$?" >> $S
for SB in 0 32 64 128; do
  HFR_SUB_BATCH=$SB timeout -k 5 600 python bench.py --workload resnet50 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_resnet50_sb$SB.json 2> gpurun_out/bench_resnet50_sb$SB.err; echo "bench resnet50 sub=$SB rc=$?" >> $S
done
for SB in 64 128; do
  HFR_SUB_BATCH=$SB timeout -k 5 600 python bench.py --workload agegender224 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_agegender224_sb$SB.json 2> gpurun_out/bench_agegender224_sb$SB.err; echo "bench agegender sub=$SB rc=$?" >> $S
done
fi
cat $S; tail -4 gpurun_out/pytest_k5.log; grep -E "cosine|passed|failed" gpurun_out/pytest_m5.log | cut -c1-200
