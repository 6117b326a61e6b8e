"""ctypes front-end of the C oracle (``nafb_oracle.c``) and of ``oracle/_ref``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "libnafb_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_hashgrid.so")
_lib = None
_ref = None


def build_oracle(force: bool = False) -> str:
    """Compile nafb_oracle.c with gcc (explicit fmaf, contraction off)."""
    src = os.path.join(_HERE, "nafb_oracle.c")
    if force or not os.path.exists(_ORACLE_SO) or os.path.getmtime(_ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w", "-o", _ORACLE_SO, src, "-lm"]
        )
    return _ORACLE_SO


def build_ref() -> bool:
    """Run oracle/build_ref.sh if the reference mount is present. Returns have_ref()."""
    if os.path.exists("/root/reference/src/encoder/hashencoder/src/hashencoder.cu"):
        subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")])
    return have_ref()


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = ctypes.CDLL(_ORACLE_SO)
        _lib.oracle_grid_index.restype = ctypes.c_uint32
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _load_ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libref_hashgrid.so not built (run oracle/build_ref.sh where /root/reference exists)")
        _ref = ctypes.CDLL(_REF_SO)
        _ref.ref_grid_index_3.restype = ctypes.c_uint32
    return _ref


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def level_offsets(num_levels=16, base_resolution=16, log2_hashmap_size=19, input_dim=3) -> np.ndarray:
    """Entry offsets per level (hashgrid.py:92-102): T_l = min(2^log2, (H*2^l + 1)^D)."""
    offs = [0]
    for i in range(num_levels):
        res = base_resolution * 2 ** i
        offs.append(offs[-1] + min(2 ** log2_hashmap_size, (res + 1) ** input_dim))
    return np.asarray(offs, dtype=np.int32)


def oracle_grid_index(D, C, hashmap_size, resolution, pos_grid) -> int:
    lib = _load()
    pg = (ctypes.c_uint32 * len(pos_grid))(*[int(v) for v in pos_grid])
    return lib.oracle_grid_index(D, C, 0, int(hashmap_size), int(resolution), pg) // C


def ref_grid_index_3(C, hashmap_size, resolution, pos_grid) -> int:
    ref = _load_ref()
    return ref.ref_grid_index_3(C, int(hashmap_size), int(resolution), *[int(v) for v in pos_grid])


def oracle_hash_forward(x01, table, offsets, H, calc_grad_inputs=False):
    """-> outputs [L,B,C] (reference FFI layout, hashencoder.cu:96) and dy_dx [B, L*D*C] or None."""
    lib = _load()
    x01, table = _f32(x01), _f32(table)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    B, D = x01.shape
    C = table.shape[1]
    L = offsets.shape[0] - 1
    out = np.zeros((L, B, C), np.float32)
    dy_dx = np.zeros((B, L * D * C), np.float32) if calc_grad_inputs else None
    lib.oracle_hash_forward(_p(x01), _p(table), _p(offsets), _p(out), B, D, C, L, int(H), _p(dy_dx))
    return out, dy_dx


def oracle_hash_backward(grad, x01, offsets, n_entries, C, H, want_f64=False):
    """grad [B, L*C] -> grad_table [n_entries, C] (+ float64 accumulation if asked)."""
    lib = _load()
    grad, x01 = _f32(grad), _f32(x01)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    B, D = x01.shape
    L = offsets.shape[0] - 1
    gg = np.zeros((n_entries, C), np.float32)
    gg64 = np.zeros((n_entries, C), np.float64) if want_f64 else None
    lib.oracle_hash_backward(_p(grad), _p(x01), _p(offsets), _p(gg), _p(gg64), B, D, C, L, int(H))
    return (gg, gg64) if want_f64 else gg


def oracle_input_backward(grad, dy_dx, B, D, C, L):
    lib = _load()
    gi = np.zeros((B, D), np.float32)
    lib.oracle_input_backward(_p(_f32(grad)), _p(_f32(dy_dx)), _p(gi), B, D, C, L)
    return gi


def oracle_corners(x01_point, offsets, level, C, H, D=3):
    """entries[2^D], weights[2^D], pos_grid[D], frac[D] for one (point, level)."""
    lib = _load()
    x = _f32(x01_point).reshape(D)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    n = 1 << D
    entry = np.zeros(n, np.uint32)
    weight = np.zeros(n, np.float32)
    pg = np.zeros(D, np.uint32)
    fr = np.zeros(D, np.float32)
    lib.oracle_corners(_p(x), _p(offsets), int(level), D, C, int(H), _p(entry), _p(weight), _p(pg), _p(fr))
    return entry, weight, pg, fr


def ref_hash_forward(x01, table, offsets, H, calc_grad_inputs=False):
    ref = _load_ref()
    x01, table = _f32(x01), _f32(table)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    B, D = x01.shape
    C = table.shape[1]
    L = offsets.shape[0] - 1
    out = np.zeros((L, B, C), np.float32)
    dy_dx = np.zeros((B, L * D * C) if calc_grad_inputs else (1,), np.float32)
    rc = ref.ref_hash_forward(_p(x01), _p(table), _p(offsets), _p(out), B, D, C, L, int(H), int(calc_grad_inputs), _p(dy_dx))
    if rc:
        raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8.")
    return out, (dy_dx if calc_grad_inputs else None)


def ref_hash_backward(grad, x01, table, offsets, H, ordered=True):
    ref = _load_ref()
    grad, x01, table = _f32(grad), _f32(x01), _f32(table)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    B, D = x01.shape
    C = table.shape[1]
    L = offsets.shape[0] - 1
    gg = np.zeros_like(table)
    rc = ref.ref_hash_backward(_p(grad), _p(x01), _p(table), _p(offsets), _p(gg), B, D, C, L, int(H), int(ordered))
    if rc:
        raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8.")
    return gg


def num_threads() -> int:
    return _load().oracle_num_threads()


class _OracleHashFn(torch.autograd.Function):
    """CPU autograd wrapper with the contract of hashgrid.py:10-71 ([L,B,C] -> [B, L*C])."""

    @staticmethod
    def forward(ctx, x01, table, offsets_np, H, use_ref):
        x = x01.detach().contiguous().numpy()
        t = table.detach().contiguous().numpy()
        fwd = ref_hash_forward if use_ref else oracle_hash_forward
        out, _ = fwd(x, t, offsets_np, H)
        L, B, C = out.shape
        ctx.save_for_backward(x01.detach())
        ctx.meta = (offsets_np, H, t.shape[0], C, use_ref, table)
        return torch.from_numpy(np.ascontiguousarray(out.transpose(1, 0, 2)).reshape(B, L * C))

    @staticmethod
    def backward(ctx, grad):
        (x01,) = ctx.saved_tensors
        offsets_np, H, n_entries, C, use_ref, table = ctx.meta
        g = grad.contiguous().numpy()
        if use_ref:
            gg = ref_hash_backward(g, x01.numpy(), table.detach().numpy(), offsets_np, H, ordered=True)
        else:
            gg = oracle_hash_backward(g, x01.numpy(), offsets_np, n_entries, C, H)
        return None, torch.from_numpy(gg), None, None, None


class OracleHashEncoder(torch.nn.Module):
    """CPU stand-in for HashEncoder (hashgrid.py:77-137): same attributes, same maths.

    normalise="div" is what torch CPU eager does for ``(x + size) / (2 * size)``;
    normalise="mul_recip" is what ATen's CUDA ``div`` does with a python-scalar divisor
    (``x * (1.0f / b)``, BinaryDivTrueKernel.cu) -- the mode the GPU reference runs in.
    """

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                 use_ref=False, normalise="div"):
        super().__init__()
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.base_resolution, self.log2_hashmap_size = base_resolution, log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.offsets_np = level_offsets(num_levels, base_resolution, log2_hashmap_size, input_dim)
        self.offsets = torch.from_numpy(self.offsets_np)
        self.embeddings = torch.nn.Parameter(torch.zeros(int(self.offsets_np[-1]), level_dim))
        self.embeddings.data.uniform_(-1e-4, 1e-4)  # hashgrid.py:111-113
        self.use_ref = use_ref
        self.normalise = normalise

    def forward(self, inputs, size=1):
        if inputs.min().item() < -size or inputs.max().item() > size:  # hashgrid.py:122-123
            raise ValueError(f"HashGrid encoder: inputs range [{inputs.min().item()}, {inputs.max().item()}] not in [{-size}, {size}]!")
        if self.normalise == "div":
            x01 = (inputs + size) / (2 * size)
        else:
            inv = (np.float32(1.0) / np.float32(2 * size)).item()
            x01 = (inputs + size) * torch.tensor(inv, dtype=torch.float32)
        prefix = list(x01.shape[:-1])
        x01 = x01.reshape(-1, self.input_dim)
        out = _OracleHashFn.apply(x01, self.embeddings, self.offsets_np, self.base_resolution, self.use_ref)
        return out.view(prefix + [self.output_dim])
