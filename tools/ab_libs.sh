#!/bin/bash
# Same-box A/B of two builds of the library: hse_facerec_tf_b200/libhfr.so (new) against libhfr_prev.so (a build of the
# previous commit placed next to it), alternating, one workload with per-layer timings.
#   gpurun -- 'bash tools/ab_libs.sh resnet50 2'
cd "$(dirname "$0")/.."
W=${1:-resnet50}; R=${2:-2}
L=hse_facerec_tf_b200
mkdir -p gpurun_out
cp $L/libhfr.so $L/libhfr_new.so
timeout 300 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "fused_gemm_pairs or gemm_pair_and_k or test_gemm_bias_act or conv2d_window" > gpurun_out/ab_pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/ab_pytest.log)"
for i in $(seq 1 $R); do
  for v in new prev; do
    cp $L/libhfr_$v.so $L/libhfr.so
    timeout 300 python bench.py --only $W --steps 30 --warmup 3 --layers --no-cpu-baseline > gpurun_out/ab_${v}_$i.json 2> gpurun_out/ab_${v}_$i.err
    echo "$v $i rc=$? $(python tools/show_bench.py gpurun_out/ab_${v}_$i.json | head -1)"
  done
done
cp $L/libhfr_new.so $L/libhfr.so
