"""Pin nafb_oracle.c against the reference's own kernel text (oracle/_ref), on fresh random
inputs.  Skipped where oracle/_ref was not built (it is built in the container that mounts
/root/reference and travels to the GPU box as a git-ignored binary)."""
import numpy as np
import pytest

from helpers import formula_table, make_rays  # noqa: F401
from oracle import hashgrid as oh

pytestmark = pytest.mark.skipif(not oh.have_ref(), reason="oracle/_ref not built")


@pytest.mark.parametrize("D,C,L,H,log2T", [(3, 2, 16, 16, 19), (2, 2, 16, 16, 19), (3, 4, 6, 8, 14), (2, 1, 10, 4, 8), (3, 8, 3, 32, 16)])
def test_fwd_bwd_bit_identical(D, C, L, H, log2T):
    rng = np.random.default_rng(D * 100 + C)
    offs = oh.level_offsets(L, H, log2T, D)
    tab = formula_table(int(offs[-1]), C, 1.0)
    B = 3000
    x = rng.uniform(0, 1, (B, D)).astype(np.float32)
    # values that straddle a floor boundary at some level: k / scale (+- 1 ulp)
    for i, lvl in enumerate(range(L)):
        s = np.float32(2.0 ** lvl * H - 1)
        v = np.float32((rng.integers(1, int(s)) - 0.5) / s) if s > 1 else np.float32(0.5)
        x[10 + 3 * i] = v
        x[11 + 3 * i] = np.nextafter(v, np.float32(0))
        x[12 + 3 * i] = np.nextafter(v, np.float32(1))
    a, da = oh.oracle_hash_forward(x, tab, offs, H, calc_grad_inputs=True)
    b, db = oh.ref_hash_forward(x, tab, offs, H, calc_grad_inputs=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))  # also proves g++ contracted the FMAs
    da, db = da.reshape(B, L, D, C), db.reshape(B, L, D, C)
    assert np.array_equal(da[:, :, D - 1], db[:, :, D - 1])  # only the well-defined axis (see nafb_oracle.c note)
    g = rng.normal(size=(B, L * C)).astype(np.float32)
    ga = oh.oracle_hash_backward(g, x, offs, tab.shape[0], C, H)
    gb = oh.ref_hash_backward(g, x, tab, offs, H, ordered=True)
    assert np.array_equal(ga.view(np.uint32), gb.view(np.uint32))
    # unordered (omp atomic) reference run: same sums up to fp32 reassociation
    gc = oh.ref_hash_backward(g, x, tab, offs, H, ordered=False)
    np.testing.assert_allclose(gc, ga, rtol=1e-4, atol=1e-5)


def test_index_random_lattice():
    rng = np.random.default_rng(5)
    offs = oh.level_offsets(16, 16, 19, 3)
    for lvl in range(16):
        T = int(offs[lvl + 1] - offs[lvl])
        res = 16 * 2 ** lvl
        for _ in range(200):
            p = [int(v) for v in rng.integers(0, res + 2, 3)]
            assert oh.oracle_grid_index(3, 2, T, res, p) == oh.ref_grid_index_3(2, T, res, p)
