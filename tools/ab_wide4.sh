#!/bin/bash
# GPU box: build with the 4-wide k-NN records, full parity suite, one 1024-scan bench line (then the default build again)
LOAMGPU_NVCC_FLAGS="-DKNN_WIDE4=1" python loam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed"; exit 1; }
python -m pytest tests -m gpu -q --maxfail=5 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --scans 1024 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[wide4]', round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
