# backward scatter: aggregation thresholds revisited with the warp-specialised kernel (levels eligible, max runs per warp)
for cfg in "8 20" "10 20" "12 20" "8 32" "10 32" "6 20"; do set -- $cfg
f=$(( ($1+1)*256 + ($2+1)*65536 ))
NAFB_DEBUG_SKIP=$f python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra --profile-steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('levels=$1 runs=$2', 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
