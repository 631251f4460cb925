run() { python bench.py --steps 3 --warmup 3 --scans 1024 --no-cpu-baseline --no-configs "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', '|', round(d['value']), round(d['e2e']['value']), round(d['e2e']['value_each_call_waited']), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items()})"; }
run
run --chunk-pairs 1024
run --chunk-pairs 768
run --point-bytes 12
run --point-bytes 12 --chunk-pairs 1024
run --point-bytes 12 --chunk-pairs 768
LOAMGPU_NVCC_FLAGS="-DEXTRACT_MINBLOCKS=8" python loam_b200/build.py --force > /dev/null 2>&1
run
run --point-bytes 12
python loam_b200/build.py --force > /dev/null 2>&1
