# round 2, call 2: the warp-specialised backward kernel, the gather token and the loss tail of the forward kernel
set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; rc=$?; echo "smoke rc $rc"; tail -3 gpurun_out/r2b_smoke.log
[ $rc -ne 0 ] && { NAFB_BWD=legacy timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc $?"; tail -15 gpurun_out/r2b_tests.log
for v in "" "NAFB_BWD=legacy" "NAFB_DEBUG_SKIP=64" "NAFB_DEBUG_SKIP=1" "NAFB_DEBUG_SKIP=3"; do
  echo "== variant [$v]"
  env $v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --profile-steps 10 2>gpurun_out/r2b_bench_err.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'loss', d['final_loss'])"
done
NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2b_stamps.log 2>&1
NAFB_DEBUG_SKIP=33 timeout 120 python scripts/stamps.py > gpurun_out/r2b_stamps_noscatter.log 2>&1
