#!/bin/bash
# GPU box: parity suite on the default build, then A/B of the k-NN variants (256- and 1024-scan bench lines each);
# the parity suite is repeated on every variant that is not the default
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v26.log 2>&1; tail -3 gpurun_out/pytest_gpu_v26.log
line() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], '|', round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})
PY
}
i=0
for flags in "" "-DKNN_FASTINS=1" "-DKNN_STACK4=1" "-DKNN_FASTINS=1 -DKNN_STACK4=1"; do
  i=$((i+1))
  LOAMGPU_NVCC_FLAGS="$flags" python loam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  if [ -n "$flags" ]; then
    python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v26_ab$i.log 2>&1; echo "[$flags] $(tail -1 gpurun_out/pytest_gpu_v26_ab$i.log)"
  fi
  for n in 256 1024; do
    python bench.py --steps 3 --warmup 3 --scans $n --no-cpu-baseline > gpurun_out/ab${i}_$n.json 2>gpurun_out/ab${i}_err.log && line gpurun_out/ab${i}_$n.json "[$flags] $n"
  done
done
python loam_b200/build.py --force > /dev/null 2>&1
