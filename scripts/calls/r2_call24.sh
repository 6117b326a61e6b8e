# pipelined push exchange at 2 GPUs: tests, chunk sweep of the chest_50 bench, phase stamps
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exchange.py -m gpu -q -x > gpurun_out/r2v_exchange_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2v_exchange_tests.log | tail -8
for c in 1 4 8 16; do
  NAFB_PUSH_CHUNKS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-extra --no-cpu-baseline > gpurun_out/r2v_bench_n2_c$c.json 2> gpurun_out/r2v_bench_n2_c$c.err; echo "bench c=$c rc $?"
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('chunks', sys.argv[2], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'div', d.get('replica_divergence'), 'err', d.get('exchange_error_word'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2v_bench_n2_c$c.json $c
done
NAFB_PUSH_CHUNKS=8 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/exchange_stamps.py 2>&1 | grep "stamps"
