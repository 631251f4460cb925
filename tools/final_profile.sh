#!/bin/bash
# GPU box, end of a round: (1) the default bench line, (2) the ncu launch list of the same command,
# (3) one `ncu --set full` capture (with source) of the ten kernels of the timed 1023-pair chunk.
# usage: bash tools/final_profile.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err || { echo "bench failed"; tail -5 gpurun_out/bench_${tag}_default.err; exit 1; }
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${tag}_default.json').read().strip().splitlines()[-1])
print('default', round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_each_call_waited']), d['roofline']['frac'], d['clocks'], d.get('cpu_baseline',{}).get('value'))
PY
# only this library's kernels, first 900 launches (= the whole device-resident phase and the start of the host-buffer
# phase): profiling the ~10,000 small torch kernels that generate the synthetic scans took 5 minutes in round 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  -k regex:'extract_ring|pack_features|bvh_build|assoc_knn|assoc_fit|lm_kernel|compact_active|init_pairs|finish_pairs|widen_kernel' -c 900 \
  --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launches_${tag}.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_${tag}.csv)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'extract_ring|pack_features|bvh_build|assoc_knn|assoc_fit|lm_kernel' --launch-skip 43 --launch-count 11 -f -o gpurun_out/prof_${tag} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_full_${tag}.log 2>&1
echo "full capture rc=$? $(ls -la gpurun_out/prof_${tag}.ncu-rep | awk '{print $5}') bytes"
