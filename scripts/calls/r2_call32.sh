# final multi-GPU pass: chest_50 weak scaling at 8 / 4 / 2 GPUs (+ side workloads at 8), exchange phase stamps, 2-GPU exchange tests
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2b_bench_n8.json 2> gpurun_out/r2b_bench_n8.err; echo "n8 rc $?"
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --no-cpu-baseline --no-extra > gpurun_out/r2b_bench_n4.json 2> gpurun_out/r2b_bench_n4.err; echo "n4 rc $?"
timeout 600 $TR --nproc-per-node 2 --master-port 29523 bench.py --gpus 2 --no-cpu-baseline --no-extra > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "n2 rc $?"
for n in 8 4 2; do python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), 'div', d.get('replica_divergence'), 'err', d.get('exchange_error_word'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, {k:(v.get('ms') or v.get('ms_per_step')) for k,v in d.get('workloads',{}).items()})" gpurun_out/r2b_bench_n$n.json; done
timeout 300 $TR --nproc-per-node 8 --master-port 29524 scripts/exchange_stamps.py 2>&1 | grep "stamps" > gpurun_out/r2b_exchange_stamps_n8.log; cat gpurun_out/r2b_exchange_stamps_n8.log
CUDA_VISIBLE_DEVICES=0,1 timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q > gpurun_out/r2b_exchange_tests_2gpu.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2b_exchange_tests_2gpu.log
