#!/bin/bash
# GPU box: per-phase cycles of the shared-memory NN build (-DBUILD_TIMING), mean per set size.  usage: bash tools/bt_run.sh [nvcc flags]
mkdir -p gpurun_out
LOAMGPU_NVCC_FLAGS="-DBUILD_TIMING $1" python loam_b200/build.py --force > /dev/null 2>&1 || { echo build failed; exit 1; }
python bench.py --steps 1 --warmup 1 --scans 128 --no-cpu-baseline --no-configs 2>&1 | grep "^BT n" | python tools/bt_summary.py
python loam_b200/build.py --force > /dev/null 2>&1
