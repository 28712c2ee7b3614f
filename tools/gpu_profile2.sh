#!/bin/bash
# per-launch list for one ResNet-50 step + full captures of selected GEMM instantiations
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "stem" > gpurun_out/pytest_stem.log 2>&1; echo "pytest stem rc=$?"
TAG=r1c
timeout -k 5 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --cache-control none -s 240 -c 61 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --workload resnet50 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
timeout -k 5 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 236 -c 12 -f -o gpurun_out/prof_${TAG}_gemm \
    python bench.py --workload resnet50 --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/pytest_stem.log
