"""Probe torch symmetric memory + NVLS multicast on this box (run under torchrun)."""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    t = sm.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(rank + 1)
    h = sm.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print("attrs", [a for a in dir(h) if not a.startswith("_")])
        print("buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(getattr(h, "multicast_ptr", 0)), "has_mc", getattr(h, "has_multicast_support", None))
        print("signal_pad_ptrs", [hex(p) for p in h.signal_pad_ptrs], "signal_pad_size", getattr(h, "signal_pad_size", None))
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (8,), torch.float32)
    print("rank", rank, "peer value", peer[:2].tolist())
except Exception as e:
    import traceback; traceback.print_exc()
    print("rank", rank, "FAILED", repr(e)[:300])
dist.barrier()
dist.destroy_process_group()
