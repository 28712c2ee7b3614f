#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary15.txt; : > $S
for W in resnet50 mobilenet192; do
  timeout -k 5 600 python bench.py --workload $W --precision tf32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_tf32.json 2> gpurun_out/bench_${W}_tf32.err; echo "bench $W tf32 rc=$?" >> $S
done
timeout -k 5 500 python bench.py --workload knn --precision tf32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn_tf32.json 2> gpurun_out/bench_knn_tf32.err; echo "bench knn tf32 rc=$?" >> $S
# final evidence: resnet50 (59 launches/step; skip warmup steps: (3 warmup + 8 rot + 1) ... take one step late in the run)
bash tools/gpu_profile_final.sh r1_resnet50 resnet50 236 59 "gemm_tc|conv_window" >> $S 2>&1
bash tools/gpu_profile_final.sh r1_mobilenet192 mobilenet192 290 29 "gemm_tc|dwconv3x3" >> $S 2>&1
cat $S
