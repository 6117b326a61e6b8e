"""ctypes binding of libnafb200.so (C ABI declared in include/nafb200.h).

This is the only place where the package touches native code.  There is NO fallback: if the
library is missing or cannot be loaded, every operator of the package raises RuntimeError.
PyTorch is used for device memory and streams only -- kernels receive raw device pointers
and the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnafb200.so")
DIAG_LIB_PATH = os.path.join(_HERE, "libnafb200_diag.so")   # diagnostics (tests / scripts only): include/nafb200_diag.h

ABI_VERSION = 10
NAFB_MAX_LEVELS = 32
NAFB_MAX_LAYERS = 8
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
ACT = {"sigmoid": 0, "relu": 1, "tanh": 2, "none": 3}
LAYOUT_LBC, LAYOUT_BLC = 0, 1
SRC_POINTS, SRC_RAYS, SRC_VOXELS = 0, 1, 2
ARITH_TC, ARITH_SIMT = 0, 1          # enum nafb_arith (nafb_mlp.arith)

c_f32p = ctypes.c_void_p
u32, u64, i32 = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32


class Grid(ctypes.Structure):
    _fields_ = [("table", ctypes.c_void_p), ("h_offsets", ctypes.c_void_p), ("D", u32), ("C", u32), ("L", u32), ("H", u32)]


class Mlp(ctypes.Structure):
    _fields_ = [("n_layers", u32), ("in_dim", u32), ("hidden", u32), ("out_dim", u32), ("skip_mask", u32), ("head", u32), ("arith", u32),
                ("W", ctypes.c_void_p * NAFB_MAX_LAYERS), ("b", ctypes.c_void_p * NAFB_MAX_LAYERS)]


class LossTail(ctypes.Structure):
    _fields_ = [("target", ctypes.c_void_p), ("mask", ctypes.c_void_p), ("chunk", u32), ("gscale", ctypes.c_float), ("loss_out", ctypes.c_void_p),
                ("dacc", ctypes.c_void_p), ("zero_pred", i32), ("ticket", ctypes.c_void_p), ("done_flag", ctypes.c_void_p),
                ("step_state", ctypes.c_void_p)]


class PixelSource(ctypes.Structure):
    _fields_ = [("projs", ctypes.c_void_p), ("mask", ctypes.c_void_p), ("valid", ctypes.c_void_p), ("n_valid", ctypes.c_void_p),
                ("order", ctypes.c_void_p), ("n_proj", u32), ("H", u32), ("W", u32), ("n_order", u32)]


class MlpGrads(ctypes.Structure):
    _fields_ = [("gW", ctypes.c_void_p * NAFB_MAX_LAYERS), ("gb", ctypes.c_void_p * NAFB_MAX_LAYERS)]


class Sampler(ctypes.Structure):
    _fields_ = [("pts", ctypes.c_void_p), ("n_points", u64), ("rays", ctypes.c_void_p), ("t_rand", ctypes.c_void_p),
                ("rng_state", ctypes.c_void_p),
                ("n_rays", u32), ("n_samples", u32), ("perturb", i32),
                ("poses", ctypes.c_void_p), ("pixels", ctypes.c_void_p), ("det_w", u32), ("det_h", u32),
                ("det_du", ctypes.c_float), ("det_dv", ctypes.c_float), ("det_u0", ctypes.c_float), ("det_v0", ctypes.c_float),
                ("det_dsd", ctypes.c_float), ("det_near", ctypes.c_float), ("det_far", ctypes.c_float), ("det_parallel", i32),
                ("n1", u32), ("n2", u32), ("n3", u32), ("i0", u32), ("i1", u32),
                ("s1", ctypes.c_double), ("s2", ctypes.c_double), ("s3", ctypes.c_double),
                ("bound", ctypes.c_float), ("clamp", ctypes.c_float)]


NAFB_MAX_RANKS = 8
STATE_STEP, STATE_SEED_LO, STATE_SEED_HI, STATE_TICKET, STATE_TICKET_FWD, STATE_LR, STATE_WORDS = 0, 1, 2, 4, 5, 6, 8   # LR: a double in words 6..7
XFLAG_ARRIVE, XFLAG_DONE, XFLAG_ERROR, XFLAG_TICKET, XFLAG_TICKET2, XFLAG_WORDS = 0, 8, 16, 17, 18, 32


class Exchange(ctypes.Structure):
    _fields_ = [("world", u32), ("rank", u32), ("param", ctypes.c_void_p * NAFB_MAX_RANKS), ("grad", ctypes.c_void_p * NAFB_MAX_RANKS),
                ("flags", ctypes.c_void_p * NAFB_MAX_RANKS), ("grad_zero", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p),
                ("exp_avg_sq", ctypes.c_void_p), ("n", u64), ("state", ctypes.c_void_p), ("mc_param", ctypes.c_void_p),
                ("mc_grad", ctypes.c_void_p), ("stage", ctypes.c_void_p * NAFB_MAX_RANKS), ("stage_slot", u64)]


_SIGNATURES = {
    "nafb_peer_alloc": (ctypes.c_int, [u64, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "nafb_peer_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "nafb_peer_close": (ctypes.c_int, [ctypes.c_void_p]),
    "nafb_peer_free": (ctypes.c_int, [ctypes.c_void_p]),
    "nafb_exchange_slice": (ctypes.c_int, [u64, u32, u32, ctypes.POINTER(u64), ctypes.POINTER(u64)]),
    "nafb_adam_exchange_step": (ctypes.c_int, [ctypes.POINTER(Exchange), ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, u32,
                                               ctypes.c_float, ctypes.c_void_p]),
    "nafb_abi_version": (ctypes.c_int, []),
    "nafb_last_error": (ctypes.c_char_p, []),
    "nafb_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 3),
    "nafb_hash_encode_forward": (ctypes.c_int, [ctypes.POINTER(Grid), c_f32p, c_f32p, u32, ctypes.c_int, ctypes.c_int, c_f32p, ctypes.c_void_p]),
    "nafb_hash_encode_forward_dtype": (ctypes.c_int, [ctypes.POINTER(Grid), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, u32,
                                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "nafb_hash_encode_backward_dtype": (ctypes.c_int, [ctypes.POINTER(Grid), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, u32,
                                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nafb_hash_encode_backward": (ctypes.c_int, [ctypes.POINTER(Grid), c_f32p, c_f32p, c_f32p, u32, ctypes.c_int, ctypes.c_int, c_f32p, c_f32p, ctypes.c_void_p]),
    "nafb_minmax": (ctypes.c_int, [c_f32p, u64, c_f32p, ctypes.c_void_p]),
    "nafb_density_stash_bytes": (u64, [ctypes.POINTER(Grid), ctypes.POINTER(Mlp), u64]),
    "nafb_density_forward": (ctypes.c_int, [ctypes.POINTER(Grid), ctypes.POINTER(Mlp), ctypes.POINTER(Sampler), ctypes.c_int, c_f32p, c_f32p, c_f32p, c_f32p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nafb_density_backward_workspace_bytes": (u64, [ctypes.POINTER(Mlp)]),
    "nafb_density_backward": (ctypes.c_int, [ctypes.POINTER(Grid), ctypes.POINTER(Mlp), ctypes.POINTER(Sampler), ctypes.c_int, c_f32p, c_f32p, ctypes.POINTER(MlpGrads), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nafb_sample_points": (ctypes.c_int, [ctypes.POINTER(Sampler), c_f32p, c_f32p, c_f32p, ctypes.c_void_p]),
    "nafb_sample_fine": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, u32, u32, u32, u32, ctypes.c_float, c_f32p, c_f32p, c_f32p, ctypes.c_void_p]),
    "nafb_generate_rays": (ctypes.c_int, [ctypes.POINTER(Sampler), c_f32p, ctypes.c_void_p]),
    "nafb_ray_integral_forward": (ctypes.c_int, [c_f32p, u32, c_f32p, c_f32p, c_f32p, c_f32p, u32, u32, ctypes.c_void_p]),
    "nafb_ray_integral_backward": (ctypes.c_int, [c_f32p, u32, c_f32p, c_f32p, c_f32p, u32, u32, ctypes.c_void_p]),
    "nafb_mse_loss": (ctypes.c_int, [c_f32p, c_f32p, ctypes.c_void_p, u32, u32, ctypes.c_float, c_f32p, c_f32p, ctypes.c_int, ctypes.c_void_p]),
    "nafb_sqdiff_f64": (ctypes.c_int, [c_f32p, c_f32p, u64, ctypes.c_void_p, u32, ctypes.c_void_p]),
    "nafb_ssim3d_f64": (ctypes.c_int, [c_f32p, c_f32p, u32, u32, u32, u32, ctypes.c_double, ctypes.c_void_p, u32, ctypes.c_void_p]),
    "nafb_ptycho_mask": (ctypes.c_int, [c_f32p, u32, u32, u32, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "nafb_draw_pixels": (ctypes.c_int, [ctypes.POINTER(PixelSource), u32, ctypes.c_void_p, c_f32p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nafb_density_forward_loss": (ctypes.c_int, [ctypes.POINTER(Grid), ctypes.POINTER(Mlp), ctypes.POINTER(Sampler), c_f32p, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.POINTER(LossTail), ctypes.c_void_p]),
    "nafb_adam_step_dev": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, u64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_float,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "nafb_adam_step": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, u64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, u32, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]),
}

_DIAG_SIGNATURES = {
    "nafb_microbench": (ctypes.c_int, [ctypes.c_int, c_f32p, u32, ctypes.c_int, c_f32p, ctypes.POINTER(u64), ctypes.c_void_p]),
    "nafb_selftest_umma": (ctypes.c_int, [c_f32p] * 6 + [ctypes.c_void_p]),
    "nafb_diag_last_error": (ctypes.c_char_p, []),
}

_lib = None
_diag = None


def diag_lib():
    """libnafb200_diag.so: the tcgen05 known-answer kernel and the L2 micro-benchmarks (kept out of the product library)."""
    global _diag
    if _diag is None:
        if not os.path.exists(DIAG_LIB_PATH):
            raise RuntimeError(f"{DIAG_LIB_PATH} is missing: run `python -m neuralvolumetricreconstructionformedicalimages_b200.build`")
        L = ctypes.CDLL(DIAG_LIB_PATH)
        for name, (res, args) in _DIAG_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _diag = L
    return _diag


def exported_symbols():
    """Names every build of the library must export (== the declarations of include/nafb200.h)."""
    return sorted(_SIGNATURES)


def lib():
    """Load libnafb200.so once. Raises RuntimeError when it is absent -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built. Run "
                "`python -m neuralvolumetricreconstructionformedicalimages_b200.build` (needs nvcc). "
                "This package has no CPU or eager-PyTorch fallback.")
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"could not load {LIB_PATH}: {e}") from e
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.nafb_abi_version() != ABI_VERSION:
            raise RuntimeError("libnafb200.so ABI version mismatch; rebuild the extension")
        _lib = L
    return _lib


def check(rc: int, value_error: bool = False):
    if rc != OK:
        msg = lib().nafb_last_error().decode()
        if rc == ERR_UNSUPPORTED and not value_error:
            raise RuntimeError(msg)
        raise (ValueError if value_error else RuntimeError)(msg)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32):
    """Same conditions as the reference's CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_* (hashencoder.cu:17-20)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        kind = "an int" if dtype == torch.int32 else "a float32"
        raise RuntimeError(f"{name} must be {kind} tensor (got {t.dtype})")
    return t


DTYPES = {torch.float32: 0, torch.float16: 1, torch.float64: 2}   # enum nafb_dtype


def require_floating(t: torch.Tensor, name: str, like: torch.Tensor = None) -> int:
    """CHECK_CUDA / CHECK_CONTIGUOUS / CHECK_IS_FLOATING (hashencoder.cu:17-20) for the dtype-dispatched op; with `like`, the
    dtype must also be `like`'s -- the reference dispatches on one tensor and `data_ptr<scalar_t>()` of the others raises
    RuntimeError on a mismatch (hashencoder.cu:392-394).  Returns the nafb_dtype code."""
    require_cuda(t, name, dtype=None)
    if t.dtype not in DTYPES:
        raise RuntimeError(f"{name} must be a floating tensor")
    if like is not None and t.dtype != like.dtype:
        raise RuntimeError(f"expected scalar type {like.dtype} but found {t.dtype} ({name})")
    return DTYPES[t.dtype]


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class RawCudaBuffer:
    """A device allocation owned by the library (nafb_peer_alloc), exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, handle: bytes = b"", owner: bool = True):
        self.ptr, self.nbytes, self.handle, self.owner = int(ptr), int(nbytes), handle, owner

    def tensor(self, byte_offset: int, numel: int, dtype, device):
        item = torch.empty(0, dtype=dtype).element_size()
        assert byte_offset % item == 0 and byte_offset + numel * item <= self.nbytes
        typestr = {torch.float32: "<f4", torch.uint32: "<u4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]

        class _View:
            __cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (self.ptr + byte_offset, False), "version": 2}

        t = torch.as_tensor(_View(), device=device)
        t._nafb_keepalive = self
        return t

    def release(self):
        if self.ptr:
            (lib().nafb_peer_free if self.owner else lib().nafb_peer_close)(ctypes.c_void_p(self.ptr))
            self.ptr = 0


def peer_alloc(nbytes: int) -> RawCudaBuffer:
    p = ctypes.c_void_p()
    h = ctypes.create_string_buffer(64)
    check(lib().nafb_peer_alloc(int(nbytes), ctypes.byref(p), h))
    return RawCudaBuffer(p.value, nbytes, h.raw, owner=True)


def peer_open(handle: bytes, nbytes: int) -> RawCudaBuffer:
    p = ctypes.c_void_p()
    check(lib().nafb_peer_open(ctypes.c_char_p(handle), ctypes.byref(p)))
    return RawCudaBuffer(p.value, nbytes, handle, owner=False)


def exchange_slice(n: int, rank: int, world: int):
    i0, i1 = u64(), u64()
    check(lib().nafb_exchange_slice(int(n), int(rank), int(world), ctypes.byref(i0), ctypes.byref(i1)))
    return int(i0.value), int(i1.value)


def lr_words(lr: float):
    """The two int32 words (low, high) of the double `lr` as the device step state stores it (NAFB_STATE_LR)."""
    w = np.array([float(lr)], dtype=np.float64).view(np.int32)
    return int(w[0]), int(w[1])


def make_grid(table: torch.Tensor, offsets_np: np.ndarray, D: int, C: int, H: int) -> Grid:
    """offsets_np must stay alive while the returned struct is used (it holds a host pointer)."""
    assert offsets_np.dtype == np.int32 and offsets_np.flags["C_CONTIGUOUS"]
    return Grid(table.data_ptr(), offsets_np.ctypes.data, D, C, offsets_np.shape[0] - 1, H)


def make_mlp(weights, biases, in_dim, hidden, out_dim, skips, head: str, arith: int = ARITH_TC) -> Mlp:
    n = len(weights)
    if n > NAFB_MAX_LAYERS:
        raise RuntimeError(f"num_layers={n} exceeds NAFB_MAX_LAYERS={NAFB_MAX_LAYERS}")
    m = Mlp()
    m.n_layers, m.in_dim, m.hidden, m.out_dim = n, in_dim, hidden, out_dim
    m.skip_mask = sum(1 << int(s) for s in set(skips) if 0 <= int(s) < 32)
    m.head = ACT[head]
    m.arith = int(arith)
    for i, (w, b) in enumerate(zip(weights, biases)):
        m.W[i] = w.data_ptr()
        m.b[i] = b.data_ptr()
    return m


def make_mlp_grads(gws, gbs) -> MlpGrads:
    g = MlpGrads()
    for i, (w, b) in enumerate(zip(gws, gbs)):
        g.gW[i] = w.data_ptr() if w is not None else None
        g.gb[i] = b.data_ptr() if b is not None else None
    return g
