"""Import the staged, unmodified reference (baseline/_ref/src, see baseline/stage_ref.sh) for the baseline arms.

Two back ends for the reference's one native op (src/encoder/hashencoder):
  cuda : the reference's own CUDA extension, pre-built for sm_100a under baseline/_ref/build  (baseline/ref_cuda_bench.py)
  host : the reference's own kernel text compiled for the host (oracle/_ref/libref_hashgrid.so, built by oracle/build_ref.sh)
         -- what `bench.py --impl reference` and bench.py's cpu_baseline time on the box's cores.
Nothing of this repository's product package is on these paths.
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available(backend: str) -> bool:
    if not os.path.isdir(os.path.join(REF, "src")):
        return False
    if backend == "cuda":
        bd = os.path.join(REF, "build")
        return os.path.isdir(bd) and any(f.endswith(".so") for f in os.listdir(bd))
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_hashgrid.so"))


class _HostBackend:
    """hash_encode_forward / hash_encode_backward (bindings.cpp:5-8) over the host build of the reference's kernels."""

    @staticmethod
    def hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, H, calc_grad_inputs, dy_dx):
        from oracle import hashgrid as oh
        out, dy = oh.ref_hash_forward(inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(), H, bool(calc_grad_inputs))
        outputs.copy_(torch.from_numpy(out))
        if calc_grad_inputs:
            dy_dx.copy_(torch.from_numpy(dy))

    @staticmethod
    def hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, H, calc_grad_inputs, dy_dx, grad_inputs):
        from oracle import hashgrid as oh
        gg = oh.ref_hash_backward(grad.detach().numpy(), inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(), H, ordered=False)
        grad_embeddings.add_(torch.from_numpy(gg))          # the kernel accumulates into the caller's zeroed buffer (hashgrid.py:59)
        if calc_grad_inputs:
            raise NotImplementedError("inputs never require grad in NAF (hashgrid.py:132)")


def import_reference(backend: str):
    """-> (get_encoder, get_network, render, calc_mse_loss) of the reference."""
    if not available(backend):
        raise RuntimeError(f"reference ({backend}) not staged: run baseline/stage_ref.sh / oracle/build_ref.sh where /root/reference is mounted")
    for name in ["matplotlib", "matplotlib.pyplot", "open3d", "skimage", "skimage.metrics", "imageio", "imageio.v2"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.metrics"].structural_similarity = lambda *a, **k: 0.0
    mod = types.ModuleType("src.encoder.hashencoder.backend")
    if backend == "cuda":
        bd = os.path.join(REF, "build")
        so = [f for f in os.listdir(bd) if f.endswith(".so")][0]
        spec = importlib.util.spec_from_file_location("_hash_encoder", os.path.join(bd, so))
        ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ext)
        mod._backend = ext
    else:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        mod._backend = _HostBackend
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    sys.modules["src.encoder.hashencoder.backend"] = mod
    from src.encoder import get_encoder
    from src.loss import calc_mse_loss
    from src.network import get_network
    from src.render import render
    return get_encoder, get_network, render, calc_mse_loss
