"""Dump the phase timestamps of the backward kernel (NAFB_DEBUG_SKIP bit 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from neuralvolumetricreconstructionformedicalimages_b200 import fused
dev = torch.device("cuda", 0)
eng = bench.build_engine(dev)
_, rays_b, projs_b, mask_b, _ = bench.synthetic_batches(4, dev, 1)
eng.use_cuda_graph = False
for i in range(3):
    eng.train_step(rays_b[i], projs_b[i], mask_b[i])
torch.cuda.synchronize()
ws = list(fused._ws_cache.values())[0]
st = ws[-4160:-64].view(torch.int64).cpu().numpy().reshape(4, 128)
for c in range(4):
    s = st[c]
    n = int((s > 0).sum())
    d = np.diff(s[:n])
    print("cta", c, "stamps", n)
    per = 17
    for k in range(0, n - 1, per):
        print("  tile", k // per, " ".join(f"{x:6d}" for x in d[k:k + per]), " | total", int(d[k:k + per].sum()))
