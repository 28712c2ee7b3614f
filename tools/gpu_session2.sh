#!/bin/bash
# Session 2: im2col implicit-GEMM convolution + ResNet-50.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary2.txt; : > $S
timeout -k 5 200 python tools/gpu_diag.py conv > gpurun_out/diag_conv.log 2>&1; echo "diag conv rc=$?" >> $S
timeout -k 5 400 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "conv2d or maxpool" > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv rc=$?" >> $S
timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -s -k "resnet50 or parity_192" > gpurun_out/pytest_resnet.log 2>&1; echo "pytest resnet rc=$?" >> $S
timeout -k 5 600 python bench.py --workload resnet50 --steps 20 --warmup 3 > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.err; echo "bench resnet50 rc=$?" >> $S
timeout -k 5 600 python bench.py --workload agegender224 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_agegender224.json 2> gpurun_out/bench_agegender224.err; echo "bench agegender rc=$?" >> $S
cat $S; cat gpurun_out/diag_conv.log | tail -40
