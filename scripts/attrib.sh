# time attribution of the training step: NAFB_DEBUG_SKIP bits: 1 = no scatter, 2 = no gather, 4/8 = no scatter for levels <6 / >=6,
# 16 = no warp aggregation in the scatter
for f in ${ATTRIB_FLAGS:-0 16 1 3}; do
  NAFB_DEBUG_SKIP=$f python bench.py --steps 30 --warmup 5 --no-cpu-baseline --profile-steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip=$f', 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
