#!/bin/bash
# multi-GPU session: usage tools/gpu_session_mg.sh <ngpus>
cd "$(dirname "$0")/.."
N=$1
mkdir -p gpurun_out
S=gpurun_out/summary_mg$N.txt; : > $S
nvidia-smi -L >> $S
timeout -k 5 600 python -m pytest tests/test_multigpu.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_mg$N.log 2>&1; echo "pytest multigpu rc=$?" >> $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
for W in resnet50 mobilenet192; do
  timeout -k 5 600 $TR bench.py --gpus $N --workload $W --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err; echo "bench $W n=$N rc=$?" >> $S
done
timeout -k 5 600 $TR bench.py --gpus $N --workload knn --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn_n$N.json 2> gpurun_out/bench_knn_n$N.err; echo "bench knn n=$N rc=$?" >> $S
cat $S; tail -3 gpurun_out/pytest_mg$N.log; tail -2 gpurun_out/bench_knn_n$N.err
