# final tree: 2-GPU exchange tests + chest_50 at 2 GPUs
set -x
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'div', d.get('replica_divergence'), 'err', d.get('exchange_error_word'), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
