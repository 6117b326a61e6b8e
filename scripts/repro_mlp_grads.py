"""MLP gradients must be bit-reproducible run to run (fixed tile assignment, fixed reduction tree): hammer the backward pass."""
import os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import test_gpu_parity as T
from neuralvolumetricreconstructionformedicalimages_b200.render import render
rng = np.random.default_rng(3)
for N, S in ((1024, 192), (384, 96), (200, 64)):
    pts = ((torch.rand(N * S, 3, device="cuda") * 2 - 1) * 0.29).contiguous()
    wts = torch.randn(N * S, 1, device="cuda")
    net = T._chest_net(table_scale=0.3)
    ref = None
    bad = 0
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
        for p in net.parameters():
            p.grad = None
        (net(pts) * wts).sum().backward()      # fixed d(sigma): the per-CTA sums and the reduction tree are deterministic
        g = torch.cat([p.grad.reshape(-1) for p in list(net.parameters())[1:]]).clone()
        if ref is None:
            ref = g
        else:
            d = (g != ref)
            if bool(d.any()):
                bad += 1
                idx = torch.nonzero(d).reshape(-1)
                print(f"N={N} S={S} iter {it}: {int(d.sum())} MLP gradient entries differ, first {idx[:12].tolist()}, max rel "
                      f"{float(((g - ref).abs() / ref.abs().clamp_min(1e-30))[d].max()):.3e}")
    print(f"N={N} S={S}: {bad} non-reproducible runs")
