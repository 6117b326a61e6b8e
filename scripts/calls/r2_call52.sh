# forward kernel: tiles claimed from a device counter (dynamic) against the fixed stride (NAFB_STATIC_TILES=1)
set -x
mkdir -p gpurun_out
for st in 1 0; do
  NAFB_STATIC_TILES=$st timeout 600 python bench.py --no-extra --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('static' if '$st'=='1' else 'dynamic', 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'loss', d['final_loss'], {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
timeout 900 python -m pytest tests -m gpu -q -x -k "not baseline_shapes" > gpurun_out/r2h_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2h_tests.log | tail -5
