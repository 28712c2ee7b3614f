#!/bin/bash
# last check of the final tree: smoke + the default bench line (what the driver runs) + MobileNet line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 5 400 python bench.py > gpurun_out/bench_resnet50.json 2> gpurun_out/bench_resnet50.err; echo "bench default rc=$?"
timeout -k 5 200 python bench.py --workload mobilenet192 > gpurun_out/bench_mobilenet192.json 2> gpurun_out/bench_mobilenet192.err; echo "bench mn rc=$?"
python tools/show_bench.py gpurun_out/bench_resnet50.json gpurun_out/bench_mobilenet192.json
