#!/bin/bash
# usage (GPU box with 8 GPUs): bash tools/scale_run.sh [tag] ["list of N"]  — the driver's 1/2/4/8 scaling launch of our arm (float4 and
# packed-xyz records) with the box's own H2D ceiling (tools/h2d_probe.py) at every N next to it
tag=${1:-x}
list=${2:-1 2 4 8}
mkdir -p gpurun_out
for n in $list; do
  for pb in 16 12; do
    out=gpurun_out/scale_${tag}_n${n}_pb${pb}.json
    if [ $n -eq 1 ]; then
      python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-configs --point-bytes $pb > $out 2> gpurun_out/scale_$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 --point-bytes $pb > $out 2> gpurun_out/scale_$n.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open('$out').read().strip().splitlines()[-1])
    print('n=$n point_bytes=$pb value', round(d['value']), 'e2e', round(d['e2e']['value']), 'each_call_waited', round(d['e2e']['value_each_call_waited']), 'ms/step', round(d['ms_per_step'],2))
except Exception as e:
    print($n, 'failed', e)
PY
  done
  if [ $n -eq 1 ]; then
    python tools/h2d_probe.py > gpurun_out/h2d_probe_${tag}_n$n.json 2>gpurun_out/probe_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 tools/h2d_probe.py > gpurun_out/h2d_probe_${tag}_n$n.json 2>gpurun_out/probe_$n.err
  fi
  tail -1 gpurun_out/h2d_probe_${tag}_n$n.json | cut -c1-600
done
