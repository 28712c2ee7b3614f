#!/bin/bash
# One parameterised GPU session (replaces the per-session scripts of round 1).  Run from anywhere:
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <stage> [<stage> ...]'
# Every stage writes into gpurun_out/ (merged back by gpurun) and appends one line to gpurun_out/summary.txt.
# Stages:
#   tests            pytest -m gpu (whole suite)
#   pytest:<expr>[:timeout]  pytest -m gpu -k <expr>
#   experimental     the opt-in tests of the experimental switches (HFR_TEST_EXPERIMENTAL=1)
#   bench            bench.py default line (all workloads) -> bench_all.json
#   bench:<w>[:tf32] one workload, per-layer timings -> bench_<w>[_tf32].json
#   ab:<ENV>=<v>[,<ENV2>=<v2>]:<w>[:tf32]  one workload with environment switches set -> bench_<w>[_tf32]_<ENV><v>.json
#   peaks            tools/peak_probe.py -> peaks_probe.json (bf16 / tf32 matmul, HBM copy)
#   launches:<w>     ncu launch list (gpu__time_duration) of one eager step -> launches_<w>.csv
#   ncu:<w>:<kernel regex>:<skip>:<count>[:tf32]   ncu --set full of matching launches -> prof_<w>_<n>.ncu-rep
#   mg:<N>           multi-GPU: tests/test_multigpu.py + bench.py under torchrun on N GPUs -> bench_all_n<N>.json
#   sweep:<N>        tools/sweep.py (config 5) on N GPUs -> sweep_n<N>.json
#   py:<script>[:arg[:arg]]  any python tool (e.g. py:tools/knn_exactness_report.py)
#   smoke            __graft_entry__.smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=gpurun_out/summary.txt
nvidia-smi -L >> $S
PY=python
note() { echo "$(date +%H:%M:%S) $*" | tee -a $S; }
i=0
for stage in "$@"; do
  i=$((i + 1))
  IFS=: read -r kind a b c d e <<< "$stage"
  case $kind in
    tests)
      timeout -k 5 900 $PY -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
      note "tests rc=$? $(tail -1 gpurun_out/pytest_gpu.log)";;
    pytest)   # pytest:<-k expression>[:timeout]: a subset of the GPU tests under a short timeout (new kernels first);
              # a failure or a hang ENDS the session - nothing else is run on top of a kernel that does not work
      timeout -k 5 ${b:-300} $PY -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider -k "$a" > gpurun_out/pytest_$i.log 2>&1
      rc=$?
      note "pytest -k '$a' rc=$rc $(tail -1 gpurun_out/pytest_$i.log)"
      if [ $rc -ne 0 ]; then cat $S; exit $rc; fi;;
    experimental)
      HFR_TEST_EXPERIMENTAL=1 timeout -k 5 600 $PY -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k experimental \
        > gpurun_out/pytest_experimental.log 2>&1
      note "experimental rc=$? $(tail -1 gpurun_out/pytest_experimental.log)";;
    bench)
      if [ -z "$a" ]; then
        timeout -k 5 900 $PY bench.py --steps 20 --warmup 5 > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err
        note "bench (all workloads) rc=$?"
      else
        P=bf16; T=""; [ "$b" = tf32 ] && { P=tf32; T=_tf32; }
        timeout -k 5 ${STAGE_TIMEOUT:-600} $PY bench.py --only $a --precision $P --steps 30 --warmup 3 --layers --no-cpu-baseline \
          > gpurun_out/bench_$a$T.json 2> gpurun_out/bench_$a$T.err
        note "bench $a $P rc=$?"
      fi;;
    ab)
      P=bf16; T=""; [ "$c" = tf32 ] && { P=tf32; T=_tf32; }
      tag=$(echo "$a" | tr -d '=,')
      env ${a//,/ } timeout -k 5 ${STAGE_TIMEOUT:-600} $PY bench.py --only $b --precision $P --steps 30 --warmup 3 --layers --no-cpu-baseline \
        > gpurun_out/bench_$b${T}_$tag.json 2> gpurun_out/bench_$b${T}_$tag.err
      note "ab $a $b $P rc=$?";;
    peaks)
      timeout -k 5 300 $PY tools/peak_probe.py > gpurun_out/peaks_probe.json 2> gpurun_out/peaks_probe.err
      note "peaks rc=$? $(cat gpurun_out/peaks_probe.json)";;
    launches)
      timeout -k 5 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -c 4000 --csv --log-file gpurun_out/launches_$a.csv \
        $PY bench.py --only $a --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-e2e > gpurun_out/launches_$a.log 2>&1
      note "launches $a rc=$?";;
    ncu)
      P=bf16; [ "$e" = tf32 ] && P=tf32
      timeout -k 5 1200 ncu --set full --clock-control none --import-source on -k "regex:$b" -s ${c:-0} -c ${d:-1} -f \
        -o gpurun_out/prof_${a}_$i $PY bench.py --only $a --precision $P --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-e2e \
        > gpurun_out/prof_${a}_$i.log 2>&1
      note "ncu $a '$b' skip=$c count=$d rc=$?";;
    mg)
      TR="$PY -m torch.distributed.run --nnodes=1 --nproc-per-node $a --master-addr 127.0.0.1 --master-port 29544"
      timeout -k 5 600 $PY -m pytest tests/test_multigpu.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_mg$a.log 2>&1
      note "pytest multigpu n=$a rc=$? $(tail -1 gpurun_out/pytest_mg$a.log)"
      timeout -k 5 900 $TR bench.py --gpus $a --steps 20 --warmup 5 > gpurun_out/bench_all_n$a.json 2> gpurun_out/bench_all_n$a.err
      note "bench n=$a rc=$?";;
    sweep)
      TR="$PY -m torch.distributed.run --nnodes=1 --nproc-per-node $a --master-addr 127.0.0.1 --master-port 29545"
      [ "$a" = 1 ] && TR=$PY
      timeout -k 5 1500 $TR tools/sweep.py --out gpurun_out/sweep_n$a.json > gpurun_out/sweep_n$a.log 2>&1
      note "sweep n=$a rc=$?";;
    py)
      timeout -k 5 1200 $PY $a $b $c > gpurun_out/py_$i.log 2>&1
      note "py $a rc=$? $(tail -1 gpurun_out/py_$i.log | cut -c1-200)";;
    smoke)
      timeout -k 5 300 $PY -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1
      note "smoke rc=$? $(tail -1 gpurun_out/smoke.log)";;
    *) note "unknown stage $stage";;
  esac
done
cat $S
