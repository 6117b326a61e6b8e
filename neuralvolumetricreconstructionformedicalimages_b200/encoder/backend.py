"""`_backend` -- the reference's native FFI, same two entry points with the same 11 / 13 positional arguments
(src/encoder/hashencoder/src/bindings.cpp:5-8, hashencoder.h:13-14), served by libnafb200.so.

A maintainer of the reference replaces src/encoder/hashencoder/backend.py (which JIT-compiles the CUDA extension at
import time, backend.py:6-16) by ``from neuralvolumetricreconstructionformedicalimages_b200.encoder.backend import _backend``;
src/encoder/hashencoder/hashgrid.py:37,66 keeps calling ``_backend.hash_encode_forward(...)`` /
``_backend.hash_encode_backward(...)`` unchanged.  tests/test_gpu_parity.py runs this shim and the reference's own compiled
extension at the same call site on the same GPU.

Ownership and semantics are the reference's: the caller allocates every tensor; ``outputs`` [L,B,C] and ``dy_dx`` are
overwritten; ``grad_embeddings`` (pre-zeroed by hashgrid.py:59) and ``grad_inputs`` are accumulated into; wrong device /
layout / dtype raise RuntimeError like the TORCH_CHECKs of hashencoder.cu:17-20; unsupported C or D raise with the
reference's message (hashencoder.cu:310,324).  fp32, fp16 and fp64 like AT_DISPATCH_FLOATING_TYPES_AND_HALF (:392,423):
every tensor of a call must have the dtype of ``inputs`` (forward) / ``grad`` (backward).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib


def _grid(embeddings, offsets, D, C, L, H):
    if offsets.dtype != torch.int32:
        raise RuntimeError("offsets must be an int tensor")
    offs = np.ascontiguousarray(offsets.detach().cpu().numpy(), dtype=np.int32)   # hashgrid.py:22 moves it to the GPU; the ABI wants it on the host
    if offs.shape[0] != L + 1:
        raise RuntimeError(f"offsets has {offs.shape[0]} entries, expected L + 1 = {L + 1}")
    return _lib.make_grid(embeddings, offs, D, C, H), offs


class _backend:
    @staticmethod
    def hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, H, calc_grad_inputs, dy_dx):
        dtype = _lib.require_floating(inputs, "inputs")
        for t, name in ((embeddings, "embeddings"), (outputs, "outputs"), (dy_dx, "dy_dx")):   # dy_dx is checked even when unused (:385)
            _lib.require_floating(t, name, like=inputs)
        g, keep = _grid(embeddings, offsets, D, C, L, H)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.lib().nafb_hash_encode_forward_dtype(ctypes.byref(g), dtype, _lib.ptr(embeddings), _lib.ptr(inputs), _lib.ptr(outputs),
                                                                 int(B), _lib.LAYOUT_LBC, int(bool(calc_grad_inputs)),
                                                                 _lib.ptr(dy_dx) if calc_grad_inputs else None, _lib.stream_ptr()))

    @staticmethod
    def hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, H, calc_grad_inputs, dy_dx, grad_inputs):
        dtype = _lib.require_floating(grad, "grad")
        for t, name in ((inputs, "inputs"), (embeddings, "embeddings"), (grad_embeddings, "grad_embeddings")):
            _lib.require_floating(t, name, like=grad)
        g, keep = _grid(grad_embeddings, offsets, D, C, L, H)
        with torch.cuda.device(inputs.device):
            _lib.check(_lib.lib().nafb_hash_encode_backward_dtype(ctypes.byref(g), dtype, _lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(grad_embeddings),
                                                                  int(B), _lib.LAYOUT_BLC, int(bool(calc_grad_inputs)),
                                                                  _lib.ptr(dy_dx) if calc_grad_inputs else None,
                                                                  _lib.ptr(grad_inputs) if calc_grad_inputs else None, _lib.stream_ptr()))


__all__ = ["_backend"]
