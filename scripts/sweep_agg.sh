for lv in 8 6 4 2; do for rn in 20 12 8 4; do
f=$(( (lv+1)*256 + (rn+1)*65536 ))
NAFB_DEBUG_SKIP=$f python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --profile-steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('levels=$lv runs=$rn', 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done; done
