set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in "NAFB_FWD_NQ=2" "NAFB_FWD_NQ=4"; do
  echo "== variant [$v]"
  env $v timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --profile-steps 10 2>gpurun_out/r2n_bench_err.log > gpurun_out/r2n_bench_$v.json
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'piped', round(d['e2e']['pipelined']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'loss', d['final_loss']); print({k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()}, d['workloads']['large_batch']['kernels_ms'])" gpurun_out/r2n_bench_$v.json
done
NAFB_FWD_NQ=2 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2n_tests.log | tail -6
