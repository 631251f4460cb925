#!/bin/bash
# usage (on the GPU box, via gpurun): bash tools/gpu_check.sh <tag> [sanitize]
#   GPU parity suite + 256/1024-scan bench lines; with "sanitize": compute-sanitizer memcheck / racecheck / synccheck
#   over tools/sanitize_smoke.py (every kernel family once, small sizes)
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
if [ "$2" = "sanitize" ]; then
  for tool in memcheck racecheck synccheck; do
    timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/sanitize_${tag}_$tool.log 2>&1
    echo "$tool rc=$? $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitize_${tag}_$tool.log) $(grep 'SUMMARY' gpurun_out/sanitize_${tag}_$tool.log | tail -1) | $(grep 'sanitize smoke ok' gpurun_out/sanitize_${tag}_$tool.log)"
  done
fi
