"""The native op at the reference's FFI boundary (hash_encode_forward / hash_encode_backward): this library against the reference's
own CUDA extension (baseline/_ref, 2-line compile fix) on the same B200, chest_50 shape: 196 608 points in render()'s order
(192 consecutive samples per ray), 16 levels x 2 features, 2^19 tables."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from baseline import ref_loader
from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder

dev = torch.device("cuda", 0)
_, rays_b, _, _, _ = bench.synthetic_batches(1, dev, seed=3)
rays = rays_b[0]
N, S = rays.shape[0], bench.N_SAMPLES
t = torch.linspace(0., 1., S, device=dev)
z = rays[:, 6:7] * (1 - t) + rays[:, 7:8] * t
pts = (rays[:, None, :3] + rays[:, None, 3:6] * z[..., None]).clamp(-0.3 + 1e-6, 0.3 - 1e-6).reshape(-1, 3).contiguous()


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.randn(pts.shape[0], 32, device=dev, generator=torch.Generator(device=dev).manual_seed(1))


def measure(enc, name):
    enc = enc.to(dev)
    with torch.no_grad():
        fwd = timeit(lambda: enc(pts, 0.3))

    def fb():
        enc.embeddings.grad = None
        enc(pts, 0.3).backward(g)
    both = timeit(fb)
    print(f"{name}: forward {fwd:.1f} us, forward + backward {both:.1f} us (backward ~ {both - fwd:.1f} us) for {pts.shape[0]} points", flush=True)
    # kernel-only times of the op itself (the module adds the range check with its host synchronisation and the normalisation)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            fb()
        torch.cuda.synchronize()
    rows = [(e.key, e.count, getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)) for e in prof.key_averages()]
    ker = [(k, c, t / max(c, 1)) for k, c, t in rows if "hash" in k.lower() or "kernel_grid" in k.lower()]
    import re
    print("    kernels: " + "; ".join(f"{(re.search(r'(k_hash\w+|kernel_grid\w*)', k) or [k])[0]} {t:.1f} us" for k, c, t in ker), flush=True)
    enc.embeddings.grad = None
    out = enc(pts, 0.3)
    out.backward(g)
    return out.detach(), enc.embeddings.grad.detach().clone()


torch.manual_seed(0)
ours = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
o_out, o_grad = measure(ours, "this library ")
if ref_loader.available("cuda"):
    r_get_encoder = ref_loader.import_reference("cuda")[0]
    ref = r_get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    with torch.no_grad():
        ref.embeddings.copy_(ours.embeddings.cpu())
    r_out, r_grad = measure(ref, "reference CUDA")
    print("encodings bit-identical:", bool(torch.equal(o_out, r_out)),
          "| table gradient: max |diff| / max |ref| =", float((o_grad - r_grad).abs().max() / r_grad.abs().max()))
else:
    print("reference CUDA build not staged")
