"""Kernel-only time of the native forward op through the C ABI (CUDA events, no module glue).  It was used with an experimental
build to compare DEPTH = 1, 2, 3, 4, 8 levels of loads in flight per thread: 56.0 / 57.3 / 57.5 / 71.4 / 69.8 us (scripts/calls/r2_call48.sh)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from neuralvolumetricreconstructionformedicalimages_b200 import _lib
from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
dev = torch.device("cuda", 0)
_, rays_b, _, _, _ = bench.synthetic_batches(1, dev, seed=3)
rays = rays_b[0]; S = bench.N_SAMPLES
t = torch.linspace(0., 1., S, device=dev); z = rays[:, 6:7] * (1 - t) + rays[:, 7:8] * t
pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).clamp(-0.3 + 1e-6, 0.3 - 1e-6).reshape(-1, 3)
x01 = ((pts + 0.3) / 0.6).contiguous()
enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19).to(dev)
grid = _lib.make_grid(enc.embeddings.detach(), enc._offsets_np, 3, 2, 16)
out = torch.empty(x01.shape[0], 32, device=dev)
L_ = _lib.lib()
def run():
    _lib.check(L_.nafb_hash_encode_forward(ctypes.byref(grid), _lib.ptr(x01), _lib.ptr(out), x01.shape[0], _lib.LAYOUT_BLC, 0, None, _lib.stream_ptr()))
for _ in range(10): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): run()
e1.record(); torch.cuda.synchronize()
print(f"nafb_hash_encode_forward, {x01.shape[0]} points: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us per launch")
