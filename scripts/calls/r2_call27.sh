# draw of step k+1 beside the optimizer of step k: select tests, exchange world-1 test, bench (sampled leg), full-schedule example
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_select.py tests/test_gpu_exchange.py -m gpu -q > gpurun_out/r2y_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2y_tests.log | tail -8
timeout 900 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,1), 'sampled', round(d['sampled']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2y_bench_n1.json
tail -3 gpurun_out/r2y_bench_n1.err
timeout 600 python examples/train_phantom.py --epochs 1501 > gpurun_out/r2y_train_phantom_full_schedule.log 2>&1; tail -2 gpurun_out/r2y_train_phantom_full_schedule.log
