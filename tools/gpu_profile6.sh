#!/bin/bash
# source-level stall samples of a stage-4 reduce 1x1 (K=1024 -> 256) and the 3x3 conv after it, CTA-pair kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k gemm_tc_kernel -s 27 -c 2 -f -o gpurun_out/prof_r1i_pair \
    python bench.py --workload resnet50 --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_r1i.log 2>&1
echo "pair rc=$?"
HFR_NO_PAIR=1 timeout -k 5 900 ncu --set full --clock-control none --import-source on -k gemm_tc_kernel -s 27 -c 2 -f -o gpurun_out/prof_r1i_nopair \
    python bench.py --workload resnet50 --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_r1i2.log 2>&1
echo "nopair rc=$?"
