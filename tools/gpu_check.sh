#!/bin/bash
# usage (on the GPU box, via gpurun): bash tools/gpu_check.sh <tag>
#   GPU parity suite + 256/1024-scan bench lines + tools/entry_smoke.py (every entry-point family once, small sizes)
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
for n in 256 1024; do
python bench.py --steps 3 --warmup 3 --scans $n --no-cpu-baseline > gpurun_out/bench_${tag}_$n.json 2>gpurun_out/bench_${tag}_err.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${tag}_$n.json').read().strip().splitlines()[-1])
print($n, round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
PY
done
python tools/entry_smoke.py 2>&1 | tail -2
