# forward kernel, chain only (NAFB_DEBUG_SKIP=2: no table gather) and complete, against the number of CTAs (1 / 2 / 3 per SM)
for skip in 2 0; do for g in 148 296 444; do
  NAFB_FWD_GRID=$g NAFB_DEBUG_SKIP=$skip python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --profile-steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip=$skip grid=$g', 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done; done
