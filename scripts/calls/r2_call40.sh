# forward kernel at 4 CTAs per SM (64 registers, spills) against 3 CTAs per SM (80 registers)
set -x
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
timeout 300 python scripts/voxel_time.py 2>&1 | grep "voxel query"
