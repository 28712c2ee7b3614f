#!/bin/bash
# final verification of the committed tree: full GPU test suite, smoke, staging bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary27.txt; : > $S
timeout -k 5 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_27.log 2>&1; echo "pytest all rc=$?" >> $S
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_27.log 2>&1; echo "smoke rc=$?" >> $S
timeout -k 5 300 python tools/bench_staging.py > gpurun_out/bench_staging.json 2> gpurun_out/bench_staging.err; echo "bench staging rc=$?" >> $S
cat $S; tail -12 gpurun_out/pytest_27.log; tail -2 gpurun_out/smoke_27.log; cat gpurun_out/bench_staging.json
