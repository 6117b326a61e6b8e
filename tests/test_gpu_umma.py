"""Known-answer test of the tcgen05 plumbing (umma.cuh): K-major and MN-major operand
descriptors over the canonical no-swizzle layout, bf16x3 split precision, TMEM accumulators."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_umma_selftest():
    from neuralvolumetricreconstructionformedicalimages_b200 import _lib
    L_ = _lib.diag_lib()   # diagnostics live outside the product library (include/nafb200_diag.h)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(128, 128)).astype(np.float32)
    X = rng.normal(size=(128, 32)).astype(np.float32)
    W = rng.normal(size=(32, 64)).astype(np.float32)
    dA, dX, dW = [torch.from_numpy(a).cuda() for a in (A, X, W)]
    D1 = torch.full((128, 32), float("nan"), device="cuda")
    D2 = torch.full((128, 64), float("nan"), device="cuda")
    D3 = torch.full((128, 32), float("nan"), device="cuda")
    assert 0 == (L_.nafb_selftest_umma(_lib.ptr(dA), _lib.ptr(dX), _lib.ptr(dW), _lib.ptr(D1), _lib.ptr(D2), _lib.ptr(D3), _lib.stream_ptr()))
    torch.cuda.synchronize()
    A64, X64, W64 = A.astype(np.float64), X.astype(np.float64), W.astype(np.float64)
    R1 = A64[:, :32] @ W64[:, :32].T
    R2 = A64[:, :32] @ W64
    R3 = A64.T @ X64
    for name, got, ref in (("D1", D1, R1), ("D2", D2, R2), ("D3", D3, R3)):
        got = got.cpu().numpy()
        err = np.abs(got - ref).max() / np.abs(ref).max()
        print(name, "max rel err", err)
        assert err < 1e-4, (name, err)   # bf16x3: ~2^-16 per product; plain bf16 would be ~4e-3
