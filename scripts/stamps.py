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
for i in range(3):
    eng.train_step(rays_b[i], projs_b[i], mask_b[i])
torch.cuda.synchronize()
st = eng._bwd_ws[-4160:-64].view(torch.int64).cpu().numpy().reshape(4, 128)
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
