# final single-GPU measurements of the round: tests, default bench (with side workloads, reference CUDA build, CPU baseline),
# reference arm, the full chest_50 schedule on the phantom
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2t_tests.log | tail -6
timeout 900 python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,1), 'sampled', round(d['sampled']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}); print(d['roofline']); print({k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()}); print(d.get('reference_cuda')); print(d.get('cpu_baseline')); print(d['clocks'])" gpurun_out/r2t_bench_n1.json
tail -3 gpurun_out/r2t_bench_n1.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 > gpurun_out/r2t_bench_reference_arm.json 2> gpurun_out/r2t_ref.err; echo "ref rc $?"; cut -c1-400 gpurun_out/r2t_bench_reference_arm.json
timeout 600 python examples/train_phantom.py --epochs 1501 > gpurun_out/r2t_train_phantom_full_schedule.log 2>&1; tail -2 gpurun_out/r2t_train_phantom_full_schedule.log
