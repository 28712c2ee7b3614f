#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k gemm_tc_kernel -s 13 -c 2 -f -o gpurun_out/prof_r1g_pw1 \
    python bench.py --workload agegender224 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_r1g.log 2>&1
echo "full rc=$?"
