#!/bin/bash
# Final evidence for one workload: launch list (time + dram traffic per launch) of one bench step, and a --set full
# capture of the dominant kernels.   usage: tools/gpu_profile_final.sh <tag> <workload> <skip> <count> <full-regex>
cd "$(dirname "$0")/.."
TAG=$1; W=$2; SKIP=$3; COUNT=$4; REGEX=$5
mkdir -p gpurun_out
timeout -k 5 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
  --cache-control none -s $SKIP -c $COUNT --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --workload $W --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list $TAG rc=$?"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 40 -c 4 -f -o gpurun_out/prof_${TAG} \
  python bench.py --workload $W --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full $TAG rc=$?"
