#!/bin/bash
# source-level stall samples of the window conv (stem, and the 64-channel 3x3 when HFR_WINDOW_CONV=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
HFR_WINDOW_CONV=1 timeout -k 5 900 ncu --set full --clock-control none --import-source on -k conv_window_kernel -c 2 -f -o gpurun_out/prof_r1h_win \
    python bench.py --workload resnet50 --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_r1h.log 2>&1
echo "full rc=$?"
