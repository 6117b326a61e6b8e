# fine-sampling kernel: its tests, the A/B against the reference's render, then the whole GPU suite
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fine_and_guard.py -m gpu -q -x > gpurun_out/r2d_fine.log 2>&1; echo "fine rc $?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2d_fine.log | tail -12
python - <<'PY'
import torch, time, importlib
R = importlib.import_module('neuralvolumetricreconstructionformedicalimages_b200.render.render')
dev = "cuda"
N, S, NF = 1024, 192, 192
g = torch.Generator().manual_seed(0)
o = torch.randn(N, 3, generator=g) * 0.2; d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
rays = torch.cat([o, d, torch.full((N, 1), 0.1), torch.full((N, 1), 0.9)], -1).to(dev)
z = (0.1 + 0.8 * torch.sort(torch.rand(N, S, generator=g), -1).values).to(dev); w = torch.rand(N, S, generator=g).to(dev)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
def eager():
    z_mid = .5 * (z[..., 1:] + z[..., :-1]); zs = R.sample_pdf(z_mid, w[..., 1:-1], NF, det=True)
    za, _ = torch.sort(torch.cat([z, zs], -1), -1)
    pts = (rays[..., None, :3] + rays[..., None, 3:6] * za[..., :, None]).clamp(-0.3, 0.3)
    return R.compute_tv_regularization(pts)
print("fine sampling stage, 1024 rays, 192 + 192 depths (det): kernel", round(t(lambda: R.sample_fine(rays, z, w, NF, True, 0.3)), 1), "us, torch operators", round(t(eager), 1), "us")
PY
