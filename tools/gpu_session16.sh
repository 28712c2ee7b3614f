#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary16.txt; : > $S
timeout -k 5 600 python -m pytest tests/test_staging_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_16.log 2>&1; echo "pytest staging rc=$?" >> $S
timeout -k 5 300 python tools/bench_staging.py > gpurun_out/bench_staging.json 2> gpurun_out/bench_staging.err; echo "bench staging rc=$?" >> $S
cat $S; tail -15 gpurun_out/pytest_16.log; cat gpurun_out/bench_staging.json
