set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
NAFB_FWD_STAMPS=1 NAFB_DEBUG_SKIP=32 timeout 120 python scripts/stamps.py > gpurun_out/r2k_fwd_stamps.log 2>&1
grep "nafb\] \|CUDA events" gpurun_out/r2k_fwd_stamps.log | tail -34
for v in "" "NAFB_DEBUG_SKIP=2"; do
  echo "== variant [$v]"
  env $v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --profile-steps 10 2>gpurun_out/r2k_bench_err.log > gpurun_out/r2k_bench_$( [ -z "$v" ] && echo default || echo nogather ).json
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'loss', d['final_loss']); print({k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()})" gpurun_out/r2k_bench_$( [ -z "$v" ] && echo default || echo nogather ).json
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2k_tests.log | tail -10
