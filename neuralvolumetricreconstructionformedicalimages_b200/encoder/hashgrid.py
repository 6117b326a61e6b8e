"""HashEncoder -- drop-in for the reference's src/encoder/hashencoder/hashgrid.py (same class,
constructor, attributes, state_dict key ``embeddings`` and call convention ``enc(x, size)``),
running on the sm_100a kernels of libnafb200.so through the C ABI.  No JIT build at import,
no CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib


_OFFSETS_CACHE = {}


def _offsets_host(offsets: torch.Tensor) -> np.ndarray:
    """The level offsets as a host int32 array.  The reference's op receives them as a device tensor (hashgrid.py:28); reading
    them back costs a device-to-host copy + a stream synchronisation, so the copy is made once per tensor VERSION."""
    key = (offsets.data_ptr(), offsets._version, offsets.numel(), str(offsets.device))
    hit = _OFFSETS_CACHE.get(key)
    if hit is None:
        if len(_OFFSETS_CACHE) > 64:
            _OFFSETS_CACHE.clear()
        hit = np.ascontiguousarray(offsets.detach().cpu().numpy(), dtype=np.int32)
        hit.setflags(write=False)
        _OFFSETS_CACHE[key] = hit
    return hit


class _hash_encode(Function):
    """Same contract as the reference autograd.Function (hashgrid.py:10-71):
    inputs [B, D] in [0, 1], embeddings [sO, C], offsets [L+1] int32 -> [B, L*C].
    The [L,B,C] -> [B,L*C] permute of hashgrid.py:40 is folded into the kernel's store."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.half)   # hashgrid.py:12: fp16 op under autocast, else the tensors' own type
    def forward(ctx, inputs, embeddings, offsets, base_resolution, calc_grad_inputs=False):
        L_ = _lib.lib()
        inputs = inputs.contiguous()
        embeddings = embeddings.contiguous()
        dtype = _lib.require_floating(inputs, "inputs")
        _lib.require_floating(embeddings, "embeddings", like=inputs)
        if offsets.dtype != torch.int32:
            raise RuntimeError("offsets must be an int tensor")
        offsets_np = _offsets_host(offsets)
        B, D = inputs.shape
        L = offsets_np.shape[0] - 1
        C = embeddings.shape[1]
        H = int(base_resolution)
        outputs = torch.empty(B, L * C, device=inputs.device, dtype=inputs.dtype)
        if calc_grad_inputs:
            dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=inputs.dtype)
        else:
            dy_dx = torch.zeros(1, device=inputs.device, dtype=inputs.dtype)
        grid = _lib.make_grid(embeddings, offsets_np, D, C, H)
        with torch.cuda.device(inputs.device):
            _lib.check(L_.nafb_hash_encode_forward_dtype(ctypes.byref(grid), dtype, _lib.ptr(embeddings), _lib.ptr(inputs), _lib.ptr(outputs), B,
                                                         _lib.LAYOUT_BLC, int(bool(calc_grad_inputs)), _lib.ptr(dy_dx), _lib.stream_ptr()))
        ctx.save_for_backward(inputs, embeddings, dy_dx)
        ctx.dtype = dtype
        ctx.offsets_np = offsets_np
        ctx.dims = [B, D, C, L, H]
        ctx.calc_grad_inputs = calc_grad_inputs
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        L_ = _lib.lib()
        inputs, embeddings, dy_dx = ctx.saved_tensors
        grad = grad.contiguous()
        _lib.require_floating(grad, "grad", like=inputs)
        B, D, C, L, H = ctx.dims
        calc_grad_inputs = ctx.calc_grad_inputs
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs) if calc_grad_inputs else None
        grid = _lib.make_grid(embeddings, ctx.offsets_np, D, C, H)
        with torch.cuda.device(inputs.device):
            _lib.check(L_.nafb_hash_encode_backward_dtype(ctypes.byref(grid), ctx.dtype, _lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(grad_embeddings),
                                                          B, _lib.LAYOUT_BLC, int(bool(calc_grad_inputs)), _lib.ptr(dy_dx),
                                                          _lib.ptr(grad_inputs), _lib.stream_ptr()))
        return grad_inputs, grad_embeddings, None, None, None


hash_encode = _hash_encode.apply


def minmax(x: torch.Tensor):
    """(min, max) of a CUDA fp32 tensor in one kernel + one 8-byte read back."""
    L_ = _lib.lib()
    x = x.contiguous()
    _lib.require_cuda(x, "inputs")
    out = torch.empty(2, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(L_.nafb_minmax(_lib.ptr(x), x.numel(), _lib.ptr(out), _lib.stream_ptr()))
    lo, hi = out.tolist()
    return lo, hi


class HashEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19):
        super().__init__()
        self.input_dim = input_dim  # coord dims, 2 or 3
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim

        if level_dim % 2 != 0:
            print("[WARN] detected HashGrid level_dim % 2 != 0: the gradient scatter cannot use paired float2 reductions")

        # entry offsets per level (hashgrid.py:92-102): T_l = min(2^log2, (H * 2^l + 1)^D)
        self.max_params = 2 ** log2_hashmap_size
        offs = [0]
        for i in range(num_levels):
            resolution = base_resolution * 2 ** i
            offs.append(offs[-1] + min(self.max_params, (resolution + 1) ** input_dim))
        self._offsets_np = np.asarray(offs, dtype=np.int32)
        self.offsets = torch.from_numpy(self._offsets_np)  # plain attribute (not a buffer), as in the reference
        self.n_params = self.offsets[-1] * level_dim

        self.embeddings = nn.Parameter(torch.zeros(offs[-1], level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"HashEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"H={self.base_resolution} params={self.embeddings.shape}")

    def forward(self, inputs, size=1):
        # inputs: [..., input_dim] in [-size, size]  ->  [..., num_levels * level_dim]
        lo, hi = minmax(inputs)
        if lo < -size or hi > size:
            raise ValueError(f"HashGrid encoder: inputs range [{lo}, {hi}] not in [{-size}, {size}]!")
        inputs = (inputs + size) / (2 * size)  # map to [0, 1]; same eager expression as hashgrid.py:125
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = hash_encode(inputs, self.embeddings, self.offsets, self.base_resolution, inputs.requires_grad)
        return outputs.view(prefix_shape + [self.output_dim])
