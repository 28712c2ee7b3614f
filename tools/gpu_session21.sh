#!/bin/bash
# bf16x2 maxpool, strided-1x1 gather bypass, KNN CTA pairs, 128-wide 1x1 tiles: full GPU test suite, then benches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary21.txt; : > $S
timeout -k 5 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_21.log 2>&1; echo "pytest all rc=$?" >> $S
timeout -k 5 300 python bench.py --workload resnet50 --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_r50_default.json 2> gpurun_out/bench_r50_default.err; echo "bench r50 rc=$?" >> $S
HFR_NO_STRIDED_GEMM=1 timeout -k 5 300 python bench.py --workload resnet50 --steps 10 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_r50_nostrided.json 2> gpurun_out/bench_r50_nostrided.err; echo "bench r50 nostrided rc=$?" >> $S
timeout -k 5 300 python bench.py --workload mobilenet192 --steps 20 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_mn.json 2> gpurun_out/bench_mn.err; echo "bench mn rc=$?" >> $S
timeout -k 5 300 python bench.py --workload knn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn.json 2> gpurun_out/bench_knn.err; echo "bench knn rc=$?" >> $S
HFR_PAIR=0 timeout -k 5 300 python bench.py --workload knn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn_nopair.json 2> gpurun_out/bench_knn_nopair.err; echo "bench knn nopair rc=$?" >> $S
timeout -k 5 300 python bench.py --workload knn --precision tf32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn_tf32.json 2> gpurun_out/bench_knn_tf32.err; echo "bench knn tf32 rc=$?" >> $S
timeout -k 5 300 python bench.py --workload resnet50 --precision tf32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_tf32.json 2> gpurun_out/bench_r50_tf32.err; echo "bench r50 tf32 rc=$?" >> $S
cat $S; tail -6 gpurun_out/pytest_21.log
python tools/show_bench.py gpurun_out/bench_r50_default.json gpurun_out/bench_r50_nostrided.json gpurun_out/bench_mn.json gpurun_out/bench_knn.json gpurun_out/bench_knn_nopair.json gpurun_out/bench_knn_tf32.json gpurun_out/bench_r50_tf32.json
