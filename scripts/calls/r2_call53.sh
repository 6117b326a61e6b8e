# last verification of the round's final tree: full GPU suite, smoke, default bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2i_tests.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,1), 'sampled', round(d['sampled']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}); print({k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()}); print(d['clocks'], d['gpu_launches'])" gpurun_out/r2i_bench_n1.json
