# role-split pipelined push exchange at 2 GPUs: tests, (chunks, pusher %) sweep of the chest_50 bench, phase stamps
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exchange.py -m gpu -q -x > gpurun_out/r2w_exchange_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2w_exchange_tests.log | tail -8
for cfg in "1 33" "2 33" "4 33" "4 50" "4 20" "8 33"; do
  set -- $cfg
  NAFB_PUSH_CHUNKS=$1 NAFB_PUSH_PCT=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 --no-extra --no-cpu-baseline > gpurun_out/r2w_bench_n2_c$1_p$2.json 2> gpurun_out/r2w_bench_n2_c$1_p$2.err; echo "bench $cfg rc $?"
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('chunks/pct', sys.argv[2], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'div', d.get('replica_divergence'), 'err', d.get('exchange_error_word'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2w_bench_n2_c$1_p$2.json "$cfg"
done
NAFB_PUSH_CHUNKS=4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/exchange_stamps.py 2>&1 | grep "stamps"
