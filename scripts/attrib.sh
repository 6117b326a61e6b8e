for f in 0 4 8 1; do
  NAFB_DEBUG_SKIP=$f python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip=$f', 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
