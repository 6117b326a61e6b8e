set -x
mkdir -p gpurun_out
NAFB_FWD_STAMPS=1 NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2j_fwd_stamps.log 2>&1
grep "nafb\] " gpurun_out/r2j_fwd_stamps.log | tail -34
for v in "NAFB_BWD_SW=8"; do
  echo "== variant [$v]"
  env $v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --profile-steps 10 2>gpurun_out/r2j_bench_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'loss', d['final_loss'])"
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2j_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2j_tests.log | tail -10
