# NVLink byte counters around 2000 steps at 8 GPUs
set -x
mkdir -p gpurun_out
nvidia-smi nvlink -gt d -i 0 | head -8
# (scripts/nvlink_bytes.py: counters before / after 2000 steps -- removed, the counters read N/A here)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/nvlink_bytes.py 2>&1 | grep "^rank" | sort > gpurun_out/r2b_nvlink_bytes_n8.log; cat gpurun_out/r2b_nvlink_bytes_n8.log
