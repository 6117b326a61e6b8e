"""Dump the phase timestamps of the warp-specialised backward kernel (NAFB_DEBUG_SKIP bit 5 = 32).
Epilogue thread 0 of CTA 0 / 1: 13 stamps per tile; scatter warp 9 of CTA 0 / 1: (tile start, d_enc ready) per tile + end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
dev = torch.device("cuda", 0)
eng = bench.build_engine(dev)
_, rays_b, projs_b, mask_b, _ = bench.synthetic_batches(4, dev, 1)
eng.use_cuda_graph = False
from neuralvolumetricreconstructionformedicalimages_b200.engine import EventTimer
for i in range(30):          # let the clocks ramp up: the stamps describe the LAST launch
    eng.train_step(rays_b[i % 4], projs_b[i % 4], mask_b[i % 4])
tm = EventTimer()
eng.profiled_step(rays_b[0], projs_b[0], mask_b[0], None, tm)
torch.cuda.synchronize()
print("last step, CUDA events (us):", {k: round(v[0] * 1e3, 1) for k, v in tm.summary().items()})
st = eng._bwd_ws[-4160:-64].view(torch.int64).cpu().numpy().reshape(4, 128).copy()
ks = st[3, 96:122].reshape(13, 2)
if ks[0, 0] > 0:
    names_k = ["start", "set-up done", "epilogue loop done", "all roles done (scatter included), TMEM released", "partials written", "grid barrier passed", "reduction done", "dump pass 0", "dump pass 1", "TMEM released", "column sums written", "first dW block dumped", "warp 0 dumped"]
    print("kernel milestones of CTA 0 (cycles / ns since start; MHz = cycles / us):")
    for i in (1, 2, 4, 5, 6, 3):
        dc, dn = ks[i, 0] - ks[0, 0], ks[i, 1] - ks[0, 1]
        print(f"  {names_k[i]:22s} {dc:8d} cyc {dn:8d} ns   {dc / max(dn, 1) * 1e3:7.0f} MHz")
st[3, 96:] = 0
names = ["dsig", "wM0", "epi0", "wM1", "epi1", "wM2", "head", "stG", "wM3", "epi3", "wM4", "epi4", "next"]
for c in range(2):
    s = st[c]
    n = int((s > 0).sum())
    d = np.diff(s[:n])
    print("epilogue cta", c, "stamps", n)
    for k in range(0, n - 1, 13):
        row = d[k:k + 13]
        print("  tile", k // 13, " ".join(f"{a}={b}" for a, b in zip(names, row)), "| total", int(row.sum()))
for c in range(2):
    s = st[2 + c]
    n = int((s > 0).sum())
    print("scatter cta", c, "stamps", n)
    for k in range(0, n - 2, 2):
        print(f"  tile {k // 2}: wait d_enc {s[k + 1] - s[k]}  scatter+next {s[k + 2] - s[k + 1]}")
    print("  whole", s[n - 1] - s[0], " epilogue whole", st[c][int((st[c] > 0).sum()) - 1] - st[c][0], "  scatter start - epilogue start", s[0] - st[c][0])
