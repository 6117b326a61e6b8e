# final single-GPU pass of the round: full GPU suite, default bench, reference arm, ncu captures (scripts/profile.sh), example
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2b_tests.log | tail -6
timeout 900 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,1), 'sampled', round(d['sampled']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}); print(d['roofline']); print({k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()}); print(d['clocks'])" gpurun_out/r2b_bench_n1.json
tail -3 gpurun_out/r2b_bench_n1.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_ref.err; echo "ref rc $?"; cut -c1-300 gpurun_out/r2b_bench_reference_arm.json
TAG=r2b bash scripts/profile.sh 2>&1 | grep -E "exit|error" 
timeout 600 python examples/train_phantom.py --epochs 1501 > gpurun_out/r2b_train_phantom_full_schedule.log 2>&1; tail -1 gpurun_out/r2b_train_phantom_full_schedule.log
