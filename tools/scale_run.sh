#!/bin/bash
# usage (GPU box with 8 GPUs): bash tools/scale_run.sh  — the driver's 1/2/4/8 scaling launch, our arm + reference arm at N=1
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1])
    print($n, round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],2))
except Exception as e:
    print($n, 'failed', e)
PY
done
