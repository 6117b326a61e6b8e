"""The CPU oracle (oracle/) against the committed reference outputs (tests/golden/*.npz).

The fixtures were produced by tests/golden/generate_golden.py from the reference itself
(Python modules imported from /root/reference; CUDA op = the reference kernel text compiled
for the host).  Integer / index results must be bit-exact; fp32 results are compared bit for
bit where the oracle restates the same operation order, otherwise with the tolerance written
next to the check.
"""
import numpy as np
import pytest
import torch

from helpers import formula_table, make_rays  # noqa: F401
from oracle import hashgrid as oh
from oracle import naf


def test_level_offsets_chest():
    offs = oh.level_offsets(16, 16, 19, 3)
    assert offs[:5].tolist() == [0, 4913, 40850, 315475, 839763]  # hashgrid.py:92-104, SURVEY 8c
    assert int(offs[-1]) == 7131219
    assert all(int(offs[i + 1] - offs[i]) == 524288 for i in range(3, 16))


def test_hash_index_kat(golden):
    kat = golden("hash_kat.npz")["kat"]
    for lvl, res, T, x, y, z, entry in kat.tolist():
        assert oh.oracle_grid_index(3, 2, T, res, (x, y, z)) == entry, (lvl, x, y, z)


def test_hash_level_modes():
    """L0-2 dense, L3-11 hash, L12/13 linear-overflow, L14/15 hash."""
    offs = oh.level_offsets(16, 16, 19, 3)
    for lvl in range(16):
        T = int(offs[lvl + 1] - offs[lvl])
        res = 16 * 2 ** lvl
        p = (3, 5, 2)
        got = oh.oracle_grid_index(3, 2, T, res, p)
        s1 = (res + 1) % 2 ** 32
        s2 = (s1 * (res + 1)) % 2 ** 32
        linear = (p[0] + p[1] * s1 + p[2] * s2) % 2 ** 32 % T
        hashed = ((p[0] * 1) ^ (p[1] * 19349663 % 2 ** 32) ^ (p[2] * 83492791 % 2 ** 32)) % T
        expect = linear if lvl in (0, 1, 2, 12, 13) else hashed
        assert got == expect, lvl


def test_position_kat():
    offs = oh.level_offsets(16, 16, 19, 3)
    x = np.float32([0.3333333, 0.3333333, 0.3333333])
    for lvl, g, f in [(0, 5, 0.49999952), (3, 42, 0.83333206), (15, 174762, 0.828125)]:
        _, _, pg, fr = oh.oracle_corners(x, offs, lvl, 2, 16)
        assert pg.tolist() == [g] * 3
        assert fr[0] == np.float32(f)


def test_hash_forward_backward_chest(golden, chest_table_unit):
    table, offs = chest_table_unit
    fx = golden("hash_chest.npz")
    out, dy_dx = oh.oracle_hash_forward(fx["x"], table, offs, 16, calc_grad_inputs=True)
    assert np.array_equal(out.view(np.uint32), fx["out_LBC"].view(np.uint32))
    B = fx["x"].shape[0]
    assert np.array_equal(dy_dx.reshape(B, 16, 3, 2)[:, :, 2, :], fx["dy_dx_last"])
    gg = oh.oracle_hash_backward(fx["grad"], fx["x"][128:256], offs, table.shape[0], 2, 16)
    rows = np.flatnonzero(np.any(gg != 0, axis=1))
    assert np.array_equal(rows, fx["grad_rows"])
    assert np.array_equal(gg[rows].view(np.uint32), fx["grad_vals"].view(np.uint32))


@pytest.mark.parametrize("tag", ["d2c4", "d3c1", "d3c8", "d2c2"])
def test_hash_small_configs(golden, tag):
    fx = golden("hash_small.npz")
    L, C, D, H, log2T = fx[f"{tag}_cfg"].tolist()
    offs = oh.level_offsets(L, H, log2T, D)
    tab = formula_table(int(offs[-1]), C, 1.0)
    out, _ = oh.oracle_hash_forward(fx[f"{tag}_x"], tab, offs, H)
    assert np.array_equal(out.view(np.uint32), fx[f"{tag}_out"].view(np.uint32))
    gg = oh.oracle_hash_backward(fx[f"{tag}_grad"], fx[f"{tag}_x"], offs, tab.shape[0], C, H)
    assert np.array_equal(gg.view(np.uint32), fx[f"{tag}_gtab"].view(np.uint32))


@pytest.mark.parametrize("steps", [1, 2, 7, 24, 40, 192, 320, 384, 576])
def test_linspace_restatement(steps):
    assert np.array_equal(naf.linspace01(steps).view(np.uint32), torch.linspace(0.0, 1.0, steps).numpy().view(np.uint32))


def _load_mlp(net, fx, prefix):
    with torch.no_grad():
        for i, lin in enumerate(net.layers):
            lin.weight.copy_(torch.from_numpy(fx[f"{prefix}W{i}"]))
            lin.bias.copy_(torch.from_numpy(fx[f"{prefix}b{i}"]))


def test_render_freq(golden):
    fx = golden("render.npz")
    net = naf.OracleDensityNetwork(naf.OracleFreqEncoder(3, 6), bound=0.3, num_layers=4, hidden_dim=32, skips=[2])
    _load_mlp(net, fx, "freq_")
    rays = torch.from_numpy(fx["freq_rays"])
    with torch.no_grad():
        r0 = naf.render(rays, net, 40, False)
        r1 = naf.render(rays, net, 40, True, t_rand=torch.from_numpy(fx["freq_t_rand"]))
    # sample positions: bit exact
    assert np.array_equal(r0["pts"].numpy().view(np.uint32), fx["freq_pts_noperturb"].view(np.uint32))
    assert np.array_equal(r1["pts"].numpy().view(np.uint32), fx["freq_pts_perturb"].view(np.uint32))
    # projections: same torch-CPU ops in the same order -> tight tolerance (sum order inside torch.sum may differ)
    np.testing.assert_allclose(r0["acc"].numpy(), fx["freq_acc_noperturb"], rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(r1["acc"].numpy(), fx["freq_acc_perturb"], rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(r1["tv_loss"].numpy(), fx["freq_tv_perturb"], rtol=1e-5)


def test_render_hierarchical_noise_two_channel(golden):
    """The oracle's restatement of the fine pass (render.py:113-126, sample_pdf :215-247), of raw_noise_std (:196-199) and of the
    two-channel weights (:207-208) against outputs of the reference's own render() (tests/golden/generate_golden.py fine), the
    random draws reproduced from the same generator seed in the reference's order."""
    fx = golden("render_fine.npz")
    S, NF = int(fx["n_samples"]), int(fx["n_fine"])
    rays = torch.from_numpy(fx["rays"])
    fine = naf.OracleDensityNetwork(naf.OracleFreqEncoder(3, 6), bound=0.3, num_layers=4, hidden_dim=32, skips=[2])
    _load_mlp(fine, fx, "fine_")
    for out_dim in (1, 2):
        net = naf.OracleDensityNetwork(naf.OracleFreqEncoder(3, 6), bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=out_dim)
        _load_mlp(net, fx, f"coarse{out_dim}_")
        for perturb, noise in ((False, 0.0), (True, 0.0), (True, 0.3)):
            tag = f"od{out_dim}_p{int(perturb)}_n{int(noise > 0)}_"
            torch.manual_seed(int(fx["seed"]))
            with torch.no_grad():
                r = naf.render_hierarchical(rays, net, fine, S, NF, perturb, noise)
            assert np.array_equal(r["pts0"].numpy().view(np.uint32), fx[tag + "pts0"].view(np.uint32)), tag
            np.testing.assert_allclose(r["acc0"].numpy(), fx[tag + "acc0"], rtol=2e-6, atol=1e-8, err_msg=tag)
            np.testing.assert_allclose(r["weights0"].numpy(), fx[tag + "weights0"], rtol=1e-5, atol=1e-7, err_msg=tag)
            # same torch-CPU ops on the same values: the fine positions agree to the last bits of the CDF arithmetic
            np.testing.assert_allclose(r["pts"].numpy(), fx[tag + "pts"], rtol=0, atol=2e-7, err_msg=tag)
            np.testing.assert_allclose(r["acc"].numpy(), fx[tag + "acc"], rtol=5e-6, atol=1e-8, err_msg=tag)
            np.testing.assert_allclose(float(r["tv_loss"]), float(fx[tag + "tv_loss"]), rtol=1e-5, err_msg=tag)
    # the noise changes the integral but not the weights (it is added inside the sum only)
    assert np.array_equal(fx["od1_p1_n0_weights0"], fx["od1_p1_n1_weights0"]) and not np.allclose(fx["od1_p1_n0_acc0"], fx["od1_p1_n1_acc0"])


def _chest_net(fx, table_scale=0.5, **kw):
    enc = oh.OracleHashEncoder(3, 16, 2, 16, 19, use_ref=False, normalise="mul_recip")  # as the fixtures (CUDA evaluation)
    with torch.no_grad():
        enc.embeddings.copy_(torch.from_numpy(formula_table(enc.embeddings.shape[0], 2, table_scale)))
    cfg = dict(bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
    cfg.update(kw)
    return naf.OracleDensityNetwork(enc, **cfg)


def test_render_chest_forward_backward(golden):
    fx = golden("render.npz")
    net = _chest_net(fx)
    _load_mlp(net, fx, "chest_")
    rays = torch.from_numpy(fx["chest_rays"])
    ret = naf.render(rays, net, 24, True, t_rand=torch.from_numpy(fx["chest_t_rand"]))
    assert np.array_equal(ret["pts"].detach().numpy().view(np.uint32), fx["chest_pts"].view(np.uint32))
    np.testing.assert_allclose(ret["raw"].detach().numpy().reshape(-1, 1), fx["chest_sigma"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(ret["acc"].detach().numpy(), fx["chest_acc"], rtol=2e-6, atol=1e-8)
    loss = naf.chunked_masked_mse(ret["acc"], torch.from_numpy(fx["chest_projs"]), None, None)
    np.testing.assert_allclose(loss.detach().numpy(), fx["chest_loss"], rtol=1e-6)
    loss.backward()
    for i, lin in enumerate(net.layers):
        np.testing.assert_allclose(lin.weight.grad.numpy(), fx[f"chest_gW{i}"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(lin.bias.grad.numpy(), fx[f"chest_gb{i}"], rtol=1e-4, atol=1e-9)
    gt = net.encoder.embeddings.grad.numpy()
    rows = np.flatnonzero(np.any(gt != 0, axis=1))
    assert np.array_equal(rows, fx["chest_gtab_rows"])
    np.testing.assert_allclose(gt[rows], fx["chest_gtab_vals"], rtol=1e-4, atol=1e-10)


@pytest.mark.parametrize("head", ["relu", "tanh", "none"])
def test_heads(golden, head):
    fx = golden("render.npz")
    net = _chest_net(fx, last_activation=head)
    _load_mlp(net, fx, "chest_")
    with torch.no_grad():
        s = net(torch.from_numpy(fx["chest_pts"]).reshape(-1, 3))
    np.testing.assert_allclose(s.numpy(), fx[f"chest_sigma_{head}"], rtol=1e-6, atol=1e-7)


def test_deep_two_skips(golden):
    fx = golden("render.npz")
    net = _chest_net(fx, num_layers=6, skips=[2, 4])
    _load_mlp(net, fx, "deep_")
    with torch.no_grad():
        s = net(torch.from_numpy(fx["chest_pts"]).reshape(-1, 3))
    np.testing.assert_allclose(s.numpy(), fx["deep_sigma"], rtol=1e-6, atol=1e-7)


BASE_GEO = dict(DSD=1500.0, DSO=1000.0, nDetector=[10, 6], dDetector=[1.5, 2.0], nVoxel=[8, 6, 4], dVoxel=[1.0, 2.0, 1.5],
                offOrigin=[0, 0, 0], offDetector=[0.5, -1.0], accuracy=0.5, filter=None)


@pytest.mark.parametrize("mode,tilt", [("cone", 0), ("parallel", 29), ("parallel", 0), ("cone", 10)])
def test_geometry_rays(golden, mode, tilt):
    fx = golden("geometry.npz")
    geo = naf.Geometry(dict(BASE_GEO, mode=mode, tilt_angle=tilt))
    tag = f"{mode}_t{tilt}"
    poses = np.stack([naf.angle2pose(geo.DSO, a, tilt) for a in fx["angles"]])
    np.testing.assert_allclose(poses, fx[f"poses_{tag}"], rtol=0, atol=1e-15)
    rays = naf.get_rays(fx["angles"], geo).numpy()
    assert np.array_equal(rays.view(np.uint32), fx[f"rays_{tag}"].view(np.uint32))


def test_geometry_misc(golden):
    fx = golden("geometry.npz")
    geo = naf.Geometry(dict(BASE_GEO, mode="cone", tilt_angle=0))
    assert np.array_equal(np.asarray(naf.get_near_far(geo)), fx["near_far"])
    assert np.array_equal(naf.get_voxels(geo), fx["voxels"])
    chest = naf.Geometry(dict(DSD=1500.0, DSO=1000.0, nDetector=[256, 256], dDetector=[1.0, 1.0], nVoxel=[128] * 3, dVoxel=[1.0] * 3,
                              offOrigin=[0, 0, 0], offDetector=[0, 0], mode="cone"))
    assert np.array_equal(np.asarray(naf.get_near_far(chest)), fx["near_far_chest"])
    np.testing.assert_allclose(fx["near_far_chest"], [0.90449, 1.09551], atol=5e-6)
    m = naf.ptycho_mask(torch.from_numpy(fx["mask_in"]), 0.007).numpy()
    assert np.array_equal(m, fx["mask_out"])
    l = naf.chunked_masked_mse(torch.from_numpy(fx["mse_pred"]), torch.from_numpy(fx["mse_tgt"]), torch.from_numpy(fx["mse_mask"]), 20)
    np.testing.assert_allclose(l.numpy(), fx["mse_chunk20"], rtol=1e-6)
    assert abs(naf.psnr_3d(fx["psnr_a"], fx["psnr_b"]) - float(fx["psnr_3d"])) < 1e-9
