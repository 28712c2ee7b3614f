#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:conv_window -s 3 -c 1 -f -o gpurun_out/prof_r1e_window \
    python bench.py --workload resnet50 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_r1e.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
