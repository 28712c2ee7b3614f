#!/bin/bash
# ncu evidence for the bench command: (1) launch list with per-launch device time, (2) full captures of the top kernels.
# usage: tools/gpu_profile.sh <tag> <bench args...>
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py "$@" --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
for K in gemm_tc dwconv3x3 stem_conv; do
  timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 3 -f -o gpurun_out/prof_${TAG}_$K \
    python bench.py "$@" --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_${TAG}_$K.log 2>&1
  echo "full $K rc=$?"
done
ls -la gpurun_out | tail -12
