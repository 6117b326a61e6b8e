import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda", 0)
eng = bench.build_engine(dev)
with torch.no_grad():
    eng.table.uniform_(-0.05, 0.05)
for n in (128, 512):
    s_half = (n * 0.001) / 2 - 0.001 / 2
    vol = torch.empty(n, n, n, device=dev)
    eng.voxel_query((n, n, n), (s_half,) * 3, out=vol)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.voxel_query((n, n, n), (s_half,) * 3, out=vol)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"voxel query {n}^3: {ms:.2f} ms  ({n**3 / ms / 1e6:.2f} G voxels/s)  checksum {float(vol.double().sum()):.6f}")
