#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary10.txt; : > $S
timeout -k 5 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_all10.log 2>&1; echo "pytest all rc=$?" >> $S
HFR_FUSE=1 timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -k "batch_sizes or parity_224" > gpurun_out/pytest_fuse10.log 2>&1; echo "pytest fused-model rc=$?" >> $S
for W in mobilenet192 resnet50 agegender224; do
  timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench $W rc=$?" >> $S
  HFR_NO_PDL=1 timeout -k 5 600 python bench.py --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_nopdl.json 2> gpurun_out/bench_${W}_nopdl.err; echo "bench $W nopdl rc=$?" >> $S
done
timeout -k 5 500 python bench.py --workload knn --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn.json 2> gpurun_out/bench_knn.err; echo "bench knn rc=$?" >> $S
cat $S; tail -4 gpurun_out/pytest_all10.log; tail -3 gpurun_out/pytest_fuse10.log
