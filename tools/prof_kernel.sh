#!/bin/bash
# GPU box: one `ncu --set full` capture (with source) of the launches of one kernel inside a short bench run.
# usage: bash tools/prof_kernel.sh <tag> <kernel regex> [launch-skip] [launch-count] [extra bench args...]
tag=$1; re=$2; skip=${3:-0}; cnt=${4:-2}; shift 4
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || { echo "bench failed"; tail -5 gpurun_out/bench_${tag}.err; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$re" --launch-skip $skip --launch-count $cnt -f \
  -o gpurun_out/prof_${tag} python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_${tag}.log 2>&1
echo "capture rc=$? $(ls -la gpurun_out/prof_${tag}.ncu-rep | awk '{print $5}') bytes"
