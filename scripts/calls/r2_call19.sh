set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-extra --profile-steps 10 2>gpurun_out/r2q_bench_err.log > gpurun_out/r2q_bench.json
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'piped', round(d['e2e']['pipelined']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'losses', d['final_loss'], d['host_loss'], d['host_loss_pipelined'])" gpurun_out/r2q_bench.json
tail -3 gpurun_out/r2q_bench_err.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2q_tests.log | tail -6
