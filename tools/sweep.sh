#!/bin/bash
# batch sweep of both networks on this box's GPUs (device-resident value + e2e per batch).
#   usage: tools/sweep.sh [ngpus]        -> gpurun_out/sweep_n<ngpus>.md
cd "$(dirname "$0")/.."
N=${1:-1}
mkdir -p gpurun_out
OUT=gpurun_out/sweep_n$N.md
echo "| workload | per-GPU batch | GPUs | img/s (device-resident) | img/s (e2e, pinned host) | ms/step | SM MHz |" > $OUT
echo "|---|---:|---:|---:|---:|---:|---:|" >> $OUT
RUN="python"
[ "$N" -gt 1 ] && RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
for W in mobilenet192 resnet50; do
  for B in 32 64 128 256 512 1024 2048 4096; do
    [ "$W" = resnet50 ] && [ "$B" -gt 1024 ] && continue      # 5.2 MB of bf16 activations per image: keep the arena modest
    timeout -k 5 300 $RUN bench.py --gpus $N --workload $W --batch $B --steps 20 --warmup 3 --no-cpu-baseline \
      > gpurun_out/sweep_${W}_b${B}_n$N.json 2> gpurun_out/sweep_${W}_b${B}_n$N.err || continue
    python - "$W" "$B" "$N" gpurun_out/sweep_${W}_b${B}_n$N.json >> $OUT <<'PY'
import json, sys
w, b, n, path = sys.argv[1:5]
d = json.loads(open(path).read().strip().splitlines()[-1])
print(f"| {w} | {b} | {n} | {d['value']:.0f} | {d['e2e']['value']:.0f} | {d['ms_per_step']} | {d['clocks'].get('sm_mhz')} |")
PY
  done
done
cat $OUT
