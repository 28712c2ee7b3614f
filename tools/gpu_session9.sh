#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python tools/gpu_diag.py fuserace > gpurun_out/diag_fuserace.log 2>&1; echo "diag rc=$?"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:dwpw -s 20 -c 2 -f -o gpurun_out/prof_r1f_dwpw \
    python bench.py --workload mobilenet192 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_r1f.log 2>&1
echo "ncu rc=$?"
cat gpurun_out/diag_fuserace.log | tail -40
