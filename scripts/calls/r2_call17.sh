# 2 GPUs: the exchange tests (peer kernels vs NCCL, rollback, failure injection, DP rule) and a short bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exchange.py -q -x -s > gpurun_out/r2o_exchange_tests.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  |exercised" gpurun_out/r2o_exchange_tests.log | tail -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('N=2 ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, 'div', d['replica_divergence'], 'err', d['exchange_error_word'], 'loss', d['final_loss'], d['final_loss_all_ranks']); print({k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()})" gpurun_out/r2o_bench_n2.json
tail -5 gpurun_out/r2o_bench_n2.err
