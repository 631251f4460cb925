#!/bin/bash
# GPU box: A/B of build flags and environment settings.  Each argument is "<nvcc flags>@@<ENV=val ENV2=val>" (either
# side may be empty); for each: rebuild the library, run the 1024-scan bench line, print the per-kernel times.
# usage: bash tools/ab.sh [--tests] "<variant>" ...
tests=0; if [ "$1" == "--tests" ]; then tests=1; shift; fi
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  flags="${v%%@@*}"; envs=""; if [[ "$v" == *"@@"* ]]; then envs="${v#*@@}"; fi
  LOAMGPU_NVCC_FLAGS="$flags" python loam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  if [ $tests == 1 ]; then env $envs python -m pytest tests -m gpu -q --maxfail=5 -x > gpurun_out/pytest_ab$i.log 2>&1; echo "[$v] $(tail -1 gpurun_out/pytest_ab$i.log)"; fi
  env $envs python bench.py --steps 3 --warmup 3 --scans 1024 --no-cpu-baseline --no-configs 2>gpurun_out/ab_err$i.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$v]', '|', round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})" || tail -3 gpurun_out/ab_err$i.log
done
python loam_b200/build.py --force > /dev/null 2>&1
