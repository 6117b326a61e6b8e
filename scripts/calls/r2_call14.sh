set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --profile-steps 10 2>gpurun_out/r2m_bench_err.log > gpurun_out/r2m_bench.json
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'piped', round(d['e2e']['pipelined']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'loss', d['final_loss']); print({k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()}); print(d['clocks'])" gpurun_out/r2m_bench.json
tail -3 gpurun_out/r2m_bench_err.log
# launch list of the graph-replayed steps (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_density|k_adam|k_mse|k_reduce|k_draw" -c 60 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile-steps 3 > gpurun_out/r2m_ncu_launch.log 2>&1
echo "launch-list exit $?"; tail -12 gpurun_out/r2m_launches.csv | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2m_tests.log | tail -6
