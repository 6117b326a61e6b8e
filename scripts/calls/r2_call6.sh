set -x
mkdir -p gpurun_out
NAFB_BWD_SW=4 NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2g_stamps_sw4.log 2>&1
NAFB_BWD_SW=8 NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2g_stamps_sw8.log 2>&1
grep "CTA(s)\|occupancy(threads 416, smem 104688)\|occupancy(threads 544, smem 104688)" gpurun_out/r2g_stamps_sw*.log
for v in "NAFB_BWD_SW=4" "NAFB_BWD_SW=8" "NAFB_BWD=legacy" "NAFB_BWD_SW=4 NAFB_DEBUG_SKIP=1" "NAFB_BWD_SW=8 NAFB_DEBUG_SKIP=1"; do
  echo "== variant [$v]"
  env $v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --profile-steps 10 2>gpurun_out/r2g_bench_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'loss', d['final_loss'])"
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2g_tests.log | tail -10
