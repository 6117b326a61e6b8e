"""Time stamps of the exchange kernel's phases (block 0 of every rank): run under torchrun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
eng = bench.build_engine(dev)
pix_b, rays_b, projs_b, mask_b, (data, geo) = bench.synthetic_batches(16, dev, 1 + rank)
eng.set_geometry(data["angles"], geo)
rows = []
for i in range(30):
    eng.train_step(None, projs_b[i % 16], mask_b[i % 16], pixels=pix_b[i % 16])
    if i >= 10:
        torch.cuda.synchronize()
        st = eng.px.flags.view(torch.int32)  # words
        raw = eng.px.flags.cpu().numpy().view(np.uint32)
        # stamps live beyond XFLAG_WORDS view? flags tensor has 32 words: words 20..27
        s = raw[20:28].view(np.uint64).astype(np.int64)
        rows.append([(s[1] - s[0]) / 1e3, (s[2] - s[1]) / 1e3, (s[3] - s[2]) / 1e3, (s[3] - s[0]) / 1e3])
r = np.array(rows)
print(f"rank {rank} mode {eng.exchange_mode}: stamps d01 {r[:,0].mean():.1f} us, d12 {r[:,1].mean():.1f} us, d23 {r[:,2].mean():.1f} us, total {r[:,3].mean():.1f} us (pull: wait-arrive / zero+slice / tail+done; push: push / wait-pushed / slice)", flush=True)
dist.barrier(); dist.destroy_process_group()
