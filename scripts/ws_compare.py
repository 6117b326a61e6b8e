import sys, json, subprocess, os
sys.path.insert(0, "/root/repo")
import torch, bench
from neuralvolumetricreconstructionformedicalimages_b200 import _lib
from neuralvolumetricreconstructionformedicalimages_b200.engine import EventTimer
dev = torch.device("cuda", 0)
for mode in (0, 2):
    _lib.check(_lib.lib().nafb_set_mlp_mode(mode))
    eng = bench.build_engine(dev)
    pix_b, rays_b, projs_b, mask_b, (data, geo) = bench.synthetic_batches(16, dev, 1)
    eng.set_geometry(data["angles"], geo)
    for i in range(5):
        eng.profiled_step(None, projs_b[i], mask_b[i], None, EventTimer(), pixels=pix_b[i])
    tm = EventTimer()
    for i in range(20):
        eng.profiled_step(None, projs_b[i % 16], mask_b[i % 16], None, tm, pixels=pix_b[i % 16])
    torch.cuda.synchronize()
    print("mode", mode, {k: round(v[0] * 1e3, 1) for k, v in tm.summary().items()})
