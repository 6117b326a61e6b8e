# line-based push exchange at 8 / 4 / 2 GPUs (own operands loaded before the polls), stamps, 2-GPU tests
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
timeout 600 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --no-cpu-baseline --no-extra --steps 400 --warmup 20 > gpurun_out/r2n_bench_n$n.json 2> gpurun_out/r2n_bench_n$n.err; echo "n$n rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), 'sampled', round(d['sampled']['ms_per_step'],4), 'div', d.get('replica_divergence'), 'err', d.get('exchange_error_word'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2n_bench_n$n.json
done
timeout 300 $TR --nproc-per-node 8 --master-port 29534 scripts/exchange_stamps.py 2>&1 | grep "stamps" | sed 's/)rank/)\nrank/g' > gpurun_out/r2n_exchange_stamps_n8.log; cat gpurun_out/r2n_exchange_stamps_n8.log | cut -c1-120
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q > gpurun_out/r2n_exchange_tests_2gpu.log 2>&1; echo "tests rc $?"; tail -2 gpurun_out/r2n_exchange_tests_2gpu.log
