# the driver's launch of the reference arm under torchrun (rank 0 alone works, the other rank exits 0)
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2e_ref_n2.json 2> gpurun_out/r2e_ref_n2.err; echo "rc $?"
wc -l gpurun_out/r2e_ref_n2.json; cut -c1-260 gpurun_out/r2e_ref_n2.json; tail -2 gpurun_out/r2e_ref_n2.err
