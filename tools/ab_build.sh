#!/bin/bash
# usage (on the GPU box): bash tools/ab_build.sh "<nvcc flags A>" "<nvcc flags B>" ...   — rebuild with each flag set, bench 1024 scans
for flags in "$@"; do
  LOAMGPU_NVCC_FLAGS="$flags" python loam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  python bench.py --steps 3 --warmup 3 --scans 1024 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$flags]', '|', round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
done
python loam_b200/build.py --force > /dev/null 2>&1
