#!/bin/bash
# First GPU session: diagnostics + parity tests + a first bench line.  Every step has its own timeout so that a hung
# kernel cannot eat the whole lease; tensor-core dependent steps are skipped when the tcgen05 GEMM diagnostic fails.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary.txt; : > $S
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 5 200 python tools/gpu_diag.py gemm > gpurun_out/diag_gemm.log 2>&1; echo "diag gemm rc=$?" >> $S
timeout -k 5 150 python tools/gpu_diag.py dw > gpurun_out/diag_dw.log 2>&1; echo "diag dw rc=$?" >> $S
TC_OK=0
if grep -q "gemm prec=2 1000x512x96: nan=0 mismatches=0" gpurun_out/diag_gemm.log && grep -q "gemm prec=1 1000x512x96: nan=0 mismatches=0" gpurun_out/diag_gemm.log; then TC_OK=1; fi
echo "TC_OK=$TC_OK" >> $S
if [ $TC_OK = 1 ]; then
  timeout -k 5 200 python tools/gpu_diag.py knn model > gpurun_out/diag_model.log 2>&1; echo "diag knn+model rc=$?" >> $S
  timeout -k 5 700 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?" >> $S
  timeout -k 5 600 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?" >> $S
  timeout -k 5 600 python -m pytest tests/test_knn_gpu.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_knn.log 2>&1; echo "pytest knn rc=$?" >> $S
  timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> $S
  timeout -k 5 400 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_mobilenet192.json 2> gpurun_out/bench_mobilenet192.err; echo "bench mobilenet rc=$?" >> $S
  timeout -k 5 500 python bench.py --workload knn --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_knn.json 2> gpurun_out/bench_knn.err; echo "bench knn rc=$?" >> $S
else
  # still collect what does not need the tensor-core path
  timeout -k 5 400 python -m pytest tests/test_kernels_gpu.py -q --tb=short -p no:cacheprovider -k "dwconv or stem or l2 or age_post or (gemm and 0])" > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels(no tc) rc=$?" >> $S
  timeout -k 5 300 python -m pytest tests/test_model_gpu.py -q --tb=short -p no:cacheprovider -s -k "fp32" > gpurun_out/pytest_model.log 2>&1; echo "pytest model(fp32) rc=$?" >> $S
fi
cat $S
tail -30 gpurun_out/diag_gemm.log
