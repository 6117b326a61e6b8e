"""Measured L2 gather / scatter denominators (run on the GPU box): python scripts/microbench.py"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuralvolumetricreconstructionformedicalimages_b200 import _lib
L = _lib.diag_lib()
names = {0: "ld.f32", 1: "ld.v2.f32", 2: "ld.v4.f32", 3: "red.f32", 4: "red.v2.f32", 5: "red.v4.f32", 6: "red.v2 warp-uniform", 7: "red.v2 lane-pairs", 8: "2x ld.v2 adjacent"}
sink = torch.zeros(4, device="cuda")
out = {}
for size_mb in (4, 57, 512):
    n = size_mb * 1024 * 1024 // 4
    buf = torch.zeros(n, device="cuda")
    for mode in range(9):
        ops = ctypes.c_uint64()
        iters = 64
        for _ in range(2):
            assert 0 == (L.nafb_microbench(mode, _lib.ptr(buf), n, iters, _lib.ptr(sink), ctypes.byref(ops), _lib.stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            assert 0 == (L.nafb_microbench(mode, _lib.ptr(buf), n, iters, _lib.ptr(sink), ctypes.byref(ops), _lib.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gops = ops.value / (ms * 1e-3) / 1e9
        out[f"{size_mb}MB {names[mode]}"] = round(gops, 1)
        print(f"{size_mb:4d} MB  {names[mode]:22s} {gops:8.1f} Gop/s   ({ms*1e3:.1f} us for {ops.value/1e6:.1f} M ops)")
json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)
