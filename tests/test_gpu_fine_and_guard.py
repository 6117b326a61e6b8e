"""The fine pass's sampling stage (reference src/render/render.py:113-126 with sample_pdf :215-247) as one kernel,
nafb_sample_fine, against the same stages written with torch operators (this package's sample_pdf / sort / clamp, which are
A/B-tested against the reference's own render in test_gpu_parity.py) on the same uniforms."""
import numpy as np
import pytest
import torch

import importlib

R = importlib.import_module("neuralvolumetricreconstructionformedicalimages_b200.render.render")   # the module (the package re-exports the function `render`)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(N, S, seed, weights="random"):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(N, 3, generator=g) * 0.2
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1)
    near = torch.rand(N, 1, generator=g) * 0.2 + 0.1
    far = near + torch.rand(N, 1, generator=g) * 0.8 + 0.2
    rays = torch.cat([o, d, near, far], -1).to(DEV)
    t = torch.sort(torch.rand(N, S, generator=g), -1).values
    z = (near + (far - near) * t).to(DEV)
    if weights == "random":
        w = torch.rand(N, S, generator=g)
    elif weights == "peaked":            # almost all mass in one bin: many denominators under the 1e-5 threshold elsewhere
        w = torch.zeros(N, S)
        w[torch.arange(N), torch.randint(1, S - 1, (N,), generator=g)] = 1.0
    else:                                # no signal at all: the + 1e-5 makes the pdf uniform
        w = torch.zeros(N, S)
    return rays, z, w.to(DEV)


def _torch_stages(rays, z_vals, weights, n_fine, det, bound):
    z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
    z_samples = R.sample_pdf(z_mid, weights[..., 1:-1], n_fine, det=det)
    z_all, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
    b = bound - 1e-6
    pts = (rays[..., None, :3] + rays[..., None, 3:6] * z_all[..., :, None]).clamp(-b, b)
    return z_all, pts, R.compute_tv_regularization(pts)


@pytest.mark.parametrize("N,S,NF,kind", [(257, 192, 192, "random"), (64, 64, 32, "random"), (33, 3, 5, "random"), (50, 100, 77, "peaked"),
                                        (20, 48, 16, "flat"), (5, 512, 512, "random")])
@pytest.mark.parametrize("det", [False, True])
def test_sample_fine_kernel_vs_torch_stages(N, S, NF, kind, det):
    rays, z, w = _inputs(N, S, seed=N + S + NF, weights=kind)
    bound = 0.3
    torch.manual_seed(1234)
    z_ref, pts_ref, tv_ref = _torch_stages(rays, z, w, NF, det, bound)
    torch.manual_seed(1234)                                      # the kernel path draws the same CPU uniforms
    z_k, pts_k, tv_k = R.sample_fine(rays, z, w, NF, det, bound)
    assert z_k.shape == (N, S + NF) and pts_k.shape == (N, S + NF, 3) and tv_k.shape == (N,)
    assert bool((z_k[:, 1:] >= z_k[:, :-1]).all()), "depths must come out sorted"
    # every coarse depth is in the output, bit for bit (merge, not resampling)
    zk, zc = z_k.cpu().numpy(), z.cpu().numpy()
    for r in range(0, N, max(1, N // 7)):
        assert np.isin(zc[r].view(np.uint32), zk[r].view(np.uint32)).all()
    # same depths as the torch stages up to the summation order of the cdf (a sample may differ in the last bits)
    # (t = (u - cdf_lo) / denom amplifies a 1-ulp difference of the cdf when denom is close to its 1e-5 floor, and a u next to a
    #  cdf value can pick the neighbouring bin: such samples move by a fraction of a bin, everything else agrees to the last bits)
    span = float((z.max() - z.min()))
    dz = np.abs(zk - z_ref.cpu().numpy())
    assert (dz > 2e-6 * span).mean() < 0.01, np.quantile(dz, [0.5, 0.99, 1.0])
    assert dz.max() < 2.5 * float((z[:, 1:] - z[:, :-1]).max()), dz.max()
    # positions: bit-identical function of the kernel's own depths (mul, add, clamp -- no FMA, like the eager expression)
    b = bound - 1e-6
    pts_from_zk = (rays[..., None, :3] + rays[..., None, 3:6] * z_k[..., :, None]).clamp(-b, b)
    assert torch.equal(pts_k, pts_from_zk)
    np.testing.assert_allclose(float(tv_k.sum()), float(tv_ref), rtol=1e-4)


def test_sample_fine_contract():
    rays, z, w = _inputs(8, 600, seed=3)
    with pytest.raises(RuntimeError, match="n_samples \\+ n_fine <= 1024"):
        R.sample_fine(rays, z, w, 600, True, 0.3)
    with pytest.raises(RuntimeError):
        R.sample_fine(rays.cpu(), z, w, 8, True, 0.3)             # no CPU path


def test_render_fine_pass_kernel_vs_operator_path():
    """render(..., n_fine > 0) end to end: the kernel path and the operator path (FUSED_FINE_SAMPLING = False) on the same seeds."""
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network
    rays, _, _ = _inputs(96, 8, seed=11)
    rays[:, 6:7], rays[:, 7:8] = 0.05, 0.55

    def net(seed):
        torch.manual_seed(seed)
        enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
        return get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(DEV)

    coarse, fine = net(1), net(2)
    out = {}
    for flag in (True, False):
        R.FUSED_FINE_SAMPLING = flag
        try:
            for perturb in (False, True):
                torch.manual_seed(7)
                out[(flag, perturb)] = R.render(rays, coarse, fine, 64, 48, perturb, 409600, 0.0)
        finally:
            R.FUSED_FINE_SAMPLING = True
    for perturb in (False, True):
        a, b = out[(True, perturb)], out[(False, perturb)]
        assert set(a) == set(b)
        d = (a["pts"] - b["pts"]).abs().max(-1).values
        assert float((d > 1e-5).float().mean()) < 0.01            # a sample next to a bin edge may hop; everything else agrees
        np.testing.assert_allclose(a["acc"].detach().cpu().numpy(), b["acc"].detach().cpu().numpy(), rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(float(a["tv_loss"]), float(b["tv_loss"]), rtol=1e-4)


def test_numerics_guard_positive_cases():
    """The guard the reference runs after every render chunk (render.py:141-144: `torch.isnan / isinf` on the outputs) and the range
    check of hashgrid.py:122 are ONE flag word OR-ed by the fused kernel: bit 0 = a position outside [-bound, bound], bit 1 = a
    non-finite activation.  The negative case (flags == 0) is asserted by the parity tests; here both bits must RISE, in the
    tensor-core and in the fp32 SIMT arithmetic, and the drop-in modules must turn them into the reference's behaviour."""
    from neuralvolumetricreconstructionformedicalimages_b200 import _lib, fused
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network
    torch.manual_seed(0)
    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(DEV)
    meta = net.fused_meta()
    pts = (torch.rand(1000, 3, device=DEV) - 0.5) * 0.58
    for arith in (_lib.ARITH_TC, _lib.ARITH_SIMT):
        meta.arith = arith
        flags = torch.zeros(1, dtype=torch.int32, device=DEV)
        fused.density_forward(meta, net.encoder.embeddings.detach(), [p.detach() for p in net.flat_params()], pts=pts, flags=flags)
        assert int(flags.item()) == 0
        bad = pts.clone()
        bad[137, 1] = 0.31                                       # outside the [-0.3, 0.3] box
        fused.density_forward(meta, net.encoder.embeddings.detach(), [p.detach() for p in net.flat_params()], pts=bad, flags=flags)
        assert int(flags.item()) == 1
        flags.zero_()
        table = net.encoder.embeddings.detach().clone()
        table[: 17 ** 3] = float("nan")                          # level 0 (dense, 17^3 entries): every point reads it
        out = fused.density_forward(meta, table, [p.detach() for p in net.flat_params()], pts=pts, flags=flags)
        assert int(flags.item()) == 2 and bool(torch.isnan(out["sigma"]).any())
    # the module raises what the reference's encoder raises (hashgrid.py:122-123)
    bad = pts.clone()
    bad[5, 0] = -0.4
    with pytest.raises(ValueError, match="not in"):
        net(bad)
