# cheaper pair test + leaky ReLU as max: parity suite, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2l_tests.log | tail -5
timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['final_loss'], {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
timeout 300 python scripts/voxel_time.py 2>&1 | grep "voxel query"
