#!/bin/bash
# A/B of the flag-gated k-NN variants kept in the tree (each: full GPU parity suite, then a 1024-scan bench line).
# Round-1 results, 1024 pairs per launch, default 12.02 ms:
#   -DKNN_WIDE4=1      4-wide walk: 11.88 ms, NN build +0.30 ms for the widening pass
#   -DKNN_TWOPHASE=1   two-phase leaf scan: 12.33 ms (smoke parity only; the full suite has not run on it)
mkdir -p gpurun_out
i=0
for flags in "" "-DKNN_WIDE4=1" "-DKNN_TWOPHASE=1" "-DKNN_TWOPHASE=1 -DKNN_WIDE4=1"; do
  i=$((i+1))
  LOAMGPU_NVCC_FLAGS="$flags" python loam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2ab$i.log 2>&1; echo "[$flags] $(tail -1 gpurun_out/pytest_r2ab$i.log)"
  python bench.py --steps 3 --warmup 3 --scans 1024 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$flags]', '|', round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
done
python loam_b200/build.py --force > /dev/null 2>&1
