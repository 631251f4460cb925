#!/bin/bash
# First GPU call of the next round: the k-NN variants prepared (compiled, SASS inspected, not yet measured) at the end
# of round 1.  Every variant runs the full GPU parity suite before its bench line.
#   -DKNN_TWOPHASE=1               leaf scan in two phases: 8 distances + filter against the K-th entry, then the
#                                  insertion network only for the kept candidates (leaf: 176 + 53 per kept candidate
#                                  SASS instructions against 640, no spills)
#   -DKNN_WIDE4=1                  4-wide walk (measured alone: k-NN -1 %, NN build +0.30 ms)
#   -DKNN_TWOPHASE=1 -DKNN_WIDE4=1 both
mkdir -p gpurun_out
i=0
for flags in "" "-DKNN_TWOPHASE=1" "-DKNN_TWOPHASE=1 -DKNN_WIDE4=1" "-DKNN_TWOPHASE=1 -DKNN_MINBLOCKS=7"; do
  i=$((i+1))
  LOAMGPU_NVCC_FLAGS="$flags" python loam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2ab$i.log 2>&1; echo "[$flags] $(tail -1 gpurun_out/pytest_r2ab$i.log)"
  python bench.py --steps 3 --warmup 3 --scans 1024 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$flags]', '|', round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
done
python loam_b200/build.py --force > /dev/null 2>&1
