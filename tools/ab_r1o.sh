#!/bin/bash
# GPU box: parity suite on the default build (warp-local selection sweeps + split ranking), then A/B against the
# variants with one / both switched off
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_gpu_v27.log 2>&1; tail -2 gpurun_out/pytest_gpu_v27.log
bash tools/ab_build.sh "" "-DEXTRACT_SPLIT_RANK=0" "-DEXTRACT_WARP_SWEEPS=0" "-DEXTRACT_WARP_SWEEPS=0 -DEXTRACT_SPLIT_RANK=0"
