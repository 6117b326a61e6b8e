# round 2, call 3: 17-warp backward kernel (8 scatter warps, 56 registers), stash tail, no gather token; new tests
set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; rc=$?; echo "smoke rc $rc"; tail -3 gpurun_out/r2c_smoke.log
for v in "" "NAFB_BWD=legacy" "NAFB_DEBUG_SKIP=1"; do
  echo "== variant [$v]"
  env $v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --profile-steps 10 2>gpurun_out/r2c_bench_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'loss', d['final_loss'])"
done
NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2c_stamps.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|PSNR|adam bit" gpurun_out/r2c_tests.log | tail -20
