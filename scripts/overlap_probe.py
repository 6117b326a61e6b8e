"""How well do the dense Adam kernel (HBM-bound) and the forward kernel (L2 gather-bound) overlap when they run concurrently?
(feasibility probe for "Adam || next forward")"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from neuralvolumetricreconstructionformedicalimages_b200 import _lib
dev = torch.device("cuda", 0)
eng = bench.build_engine(dev)
pix_b, rays_b, projs_b, mask_b, (data, geo) = bench.synthetic_batches(8, dev, 1)
eng.set_geometry(data["angles"], geo)
for i in range(4):
    eng.train_step(None, projs_b[i], mask_b[i], pixels=pix_b[i])
torch.cuda.synchronize()
L = _lib.lib()
s = eng._static[(1024, True)]
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
grid, mlp = eng.meta.grid(eng.table), eng.meta.mlp(eng.mlp_params)
smp = eng._ray_sampler(None, s["pixels"], None)

def adam(stream):
    _lib.check(L.nafb_adam_step_dev(_lib.ptr(eng.flat_param), _lib.ptr(eng.flat_grad), _lib.ptr(eng.exp_avg), _lib.ptr(eng.exp_avg_sq), eng.n_params,
                                    0.9, 0.999, 1e-8, 1.0, 1, _lib.ptr(eng.state), stream.cuda_stream))

def fwd(stream):
    _lib.check(L.nafb_density_forward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), _lib.SRC_RAYS, None, _lib.ptr(s["acc"]), None, None, None,
                                      _lib.ptr(s["stash"]), stream.cuda_stream))

def timeit(fn, n=30):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

cur = torch.cuda.current_stream()
def both():
    sA.wait_stream(cur); sB.wait_stream(cur)
    adam(sA); fwd(sB)
    cur.wait_stream(sA); cur.wait_stream(sB)
print("adam alone %.1f us, fwd alone %.1f us, concurrent %.1f us" % (timeit(lambda: adam(cur)), timeit(lambda: fwd(cur)), timeit(both)))
