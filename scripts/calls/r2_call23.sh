# one-pass draw kernel: select tests, a timing of the draw alone, bench (sampled leg), full-schedule example
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_select.py -m gpu -q > gpurun_out/r2u_select.log 2>&1; echo "tests rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2u_select.log | tail -8
timeout 300 python - <<'PY'
import torch
from neuralvolumetricreconstructionformedicalimages_b200.dataset.mask import PixelSampler
dev = "cuda"
projs = torch.rand(50, 256, 256, device=dev)
s = PixelSampler(projs, None, 0.0, seed=1)
for n in (1024, 4096, 8192):
    pix = torch.empty(n, 3, dtype=torch.int32, device=dev); val = torch.empty(n, device=dev); msk = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(5): s.draw_into(n, pix, val, msk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): s.draw_into(n, pix, val, msk)
    e1.record(); torch.cuda.synchronize()
    print("draw", n, "rays of 65536 candidates:", round(e0.elapsed_time(e1) / 200 * 1e3, 1), "us")
s.check()
PY
timeout 900 python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; echo "bench rc $?"
python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']/1e6,1), 'sampled', round(d['sampled']['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" gpurun_out/r2u_bench_n1.json
tail -3 gpurun_out/r2u_bench_n1.err
timeout 600 python examples/train_phantom.py --epochs 1501 > gpurun_out/r2u_train_phantom_full_schedule.log 2>&1; tail -2 gpurun_out/r2u_train_phantom_full_schedule.log
