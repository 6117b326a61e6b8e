# push exchange with self-validating 128-byte lines (no barrier between reduce-scatter and all-gather): 2 GPUs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exchange.py -m gpu -q -x > gpurun_out/r2m_exchange_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2m_exchange_tests.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 20 --no-extra --no-cpu-baseline > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), 'sampled', round(d['sampled']['ms_per_step'],4), 'div', d.get('replica_divergence'), 'err', d.get('exchange_error_word'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2m_bench_n2.json
tail -3 gpurun_out/r2m_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/exchange_stamps.py 2>&1 | grep "stamps" | sed 's/)rank/)\nrank/g'
