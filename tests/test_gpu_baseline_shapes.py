"""Parity at the shapes BASELINE.json names (not at toy sizes):

  * large batch        65 536 rays x 384 samples, one step, against the REFERENCE's own CUDA build on the same GPU;
  * 512^3 voxel query  1e5 random voxels + both sides of every slab seam of the 2- / 4- / 8-GPU sharding against the oracle;
  * lamino_chip        parallel beam, 29 degree tilt, ptychography mask, through NAFEngine.train_step(pixels=...) against the oracle;
  * chest_50           full shape (128^3, 50 x 256 x 256, 1024 x 192) for 2 000 iterations in PRODUCTION mode (pixel source,
                       in-kernel uniforms, CUDA graph, fused loss) next to the reference's CUDA build: PSNR-3D within 0.1 dB and
                       SSIM-3D within 0.005 (BASELINE.json north_star; reference train.py:220-258).

Reference citations: src/render/render.py:31-147, src/encoder/hashencoder/hashgrid.py:118-137, train.py:48-135 / :220-286,
src/dataset/tigre.py:354-382 / :388-400 / :463-528, src/utils/util.py:55-139 / :196-205.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if torch.cuda.is_available():
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import phantom as PH
    from neuralvolumetricreconstructionformedicalimages_b200.dataset.mask import PixelSampler, get_ptycho_mask
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.engine import NAFEngine
    from neuralvolumetricreconstructionformedicalimages_b200.loss import calc_mse_loss
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network
    from neuralvolumetricreconstructionformedicalimages_b200.render import render
    from neuralvolumetricreconstructionformedicalimages_b200.utils import get_psnr_3d, get_ssim_3d

from helpers import formula_table, make_rays
from oracle import hashgrid as oh
from oracle import naf

DEV = "cuda"


def _net(table_scale=None, seed=0):
    torch.manual_seed(seed)
    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
    if table_scale is not None:
        with torch.no_grad():
            enc.embeddings.copy_(torch.from_numpy(formula_table(enc.embeddings.shape[0], 2, table_scale)))
    return net.to(DEV)


def _oracle_of(net):
    enc = oh.OracleHashEncoder(normalise="mul_recip")
    o = naf.OracleDensityNetwork(enc, bound=net.bound, num_layers=len(net.layers), hidden_dim=32, skips=net.skips,
                                 last_activation=net.last_activation)
    with torch.no_grad():
        enc.embeddings.copy_(net.encoder.embeddings.detach().cpu())
        for a, b in zip(o.layers, net.layers):
            a.weight.copy_(b.weight.detach().cpu())
            a.bias.copy_(b.bias.detach().cpu())
    return o


def _reference(backend="cuda"):
    from baseline import ref_loader
    if not ref_loader.available(backend):
        pytest.skip("the reference's CUDA build is not staged on this machine (baseline/stage_ref.sh)")
    return ref_loader.import_reference(backend)


def _reference_twin(net, r_get_encoder, r_get_network):
    r_enc = r_get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    r_net = r_get_network("mlp")(r_enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(DEV)
    with torch.no_grad():
        r_net.encoder.embeddings.copy_(net.encoder.embeddings)
        for a, b in zip(r_net.layers, net.layers):
            a.weight.copy_(b.weight)
            a.bias.copy_(b.bias)
    return r_net


# ----------------------------------------------------------------------------------------------------- large batch
def test_large_batch_step_vs_the_reference_cuda_build():
    """BASELINE config 4: 65 536 rays x 384 samples (25.2 M points) in ONE render call, forward + loss + backward, against the
    reference's render / DensityNetwork / HashEncoder (its CUDA extension) / calc_mse_loss with the same parameters, rays and
    uniforms.  Sample positions bit-exact (all 25.2 M), loss rtol 1e-4, table gradient relative L2 < 2e-4."""
    r_get_encoder, r_get_network, r_render, r_calc_mse_loss = _reference()
    rng = np.random.default_rng(41)
    N, S = 65536, 384
    net = _net(table_scale=0.3)
    r_net = _reference_twin(net, r_get_encoder, r_get_network)
    rays = torch.from_numpy(make_rays(N, rng)).to(DEV)
    projs = torch.from_numpy(rng.uniform(0, 0.05, N).astype(np.float32)).to(DEV)
    t_rand = torch.rand(N, S, device=DEV, generator=torch.Generator(device=DEV).manual_seed(41))
    real = torch.rand
    torch.rand = lambda *a, **k: t_rand.clone()
    try:
        r_ret = r_render(rays, r_net, None, S, 0, True, 409600, 0.0)
        ret = render(rays, net, None, S, 0, True, 409600, 0.0)
    finally:
        torch.rand = real
    assert ret["pts"].shape == (N, S, 3)
    assert torch.equal(ret["pts"].view(torch.int32), r_ret["pts"].detach().view(torch.int32))      # bit-exact, every point
    sub = torch.from_numpy(rng.choice(N, 4096, replace=False)).to(DEV)
    np.testing.assert_allclose(ret["acc"][sub].detach().cpu().numpy(), r_ret["acc"][sub].detach().cpu().numpy(), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(ret["acc"].detach(), r_ret["acc"].detach(), rtol=1e-4, atol=1e-7)
    r_loss, loss = {"loss": 0.0}, {"loss": 0.0}
    r_calc_mse_loss(r_loss, projs, r_ret["acc"])
    calc_mse_loss(loss, projs, ret["acc"])
    np.testing.assert_allclose(float(loss["loss"]), float(r_loss["loss"]), rtol=1e-4)
    r_loss["loss"].backward()
    del r_ret
    loss["loss"].backward()
    ga, gb = net.encoder.embeddings.grad, r_net.encoder.embeddings.grad
    rel = float((ga - gb).double().norm() / gb.double().norm())
    assert rel < 2e-4, rel
    assert torch.equal(ga != 0, gb != 0)                                                            # the same set of touched entries
    for a, b in zip(net.layers, r_net.layers):
        for pa, pb in ((a.weight, b.weight), (a.bias, b.bias)):
            # MLP gradients: sums over 25.2 M points with heavy cancellation; bf16x3 products here, fp32 cuBLAS (another
            # summation order) there -- the mixed-precision bound of DESIGN.md section 2 (5e-3), as a relative L2 norm
            relp = float((pa.grad - pb.grad).double().norm() / pb.grad.double().norm())
            assert relp < 5e-3, relp


# ----------------------------------------------------------------------------------------------------- 512^3 voxel query
def test_voxel_query_512_random_voxels_and_slab_seams_vs_oracle():
    """BASELINE config 5: the forward-only fused query of the 512^3 lattice (tigre.py:388-400 + train.py:246-250).  The whole
    volume in one call and as 8 outermost-index slabs; 1e5 random voxels and both sides of every slab seam of the 2- / 4- / 8-rank
    sharding are compared with the oracle evaluated on the reference's float64-linspace voxel centres (rtol 2e-5)."""
    n = 512
    net = _net(table_scale=0.3)
    eng = NAFEngine(net, n_samples=8, use_cuda_graph=False)
    s_half = (n * 0.001) / 2 - 0.001 / 2
    vol = eng.voxel_query((n, n, n), (s_half,) * 3)
    assert vol.shape == (n, n, n) and bool(torch.isfinite(vol).all())
    # slabs tile the volume bit for bit
    from neuralvolumetricreconstructionformedicalimages_b200 import parallel
    for world in (2, 8):
        for rank in range(world):
            i0, i1 = parallel.shard_range(n, rank, world)
            part = eng.voxel_query((n, n, n), (s_half,) * 3, slab=(i0, i1))
            assert torch.equal(part, vol[i0:i1]), (world, rank)
            del part
    rng = np.random.default_rng(7)
    idx = [rng.integers(0, n, (100000, 3))]
    for world in (2, 4, 8):
        for rank in range(1, world):
            i0, _ = parallel.shard_range(n, rank, world)
            jk = rng.integers(0, n, (64, 2))
            idx += [np.concatenate([np.full((64, 1), i0 - 1), jk], 1), np.concatenate([np.full((64, 1), i0), jk], 1)]
    corners = np.array([[a, b, c] for a in (0, n - 1) for b in (0, n - 1) for c in (0, n - 1)])
    idx = np.concatenate(idx + [corners], 0)
    lin = np.linspace(-s_half, s_half, n)                       # float64, then the fp32 cast of tigre.py:277
    xyz = np.stack([lin[idx[:, 0]], lin[idx[:, 1]], lin[idx[:, 2]]], -1).astype(np.float32)
    o = _oracle_of(net)
    with torch.no_grad():
        ref = naf.run_network(torch.from_numpy(xyz), o, 409600).squeeze(-1).numpy()
    got = vol[torch.from_numpy(idx[:, 0]).to(DEV), torch.from_numpy(idx[:, 1]).to(DEV), torch.from_numpy(idx[:, 2]).to(DEV)].cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)


# ----------------------------------------------------------------------------------------------------- lamino_chip
def _lamino_geometry(n_proj=12, det_w=356, det_h=256):
    """config/lamino_chip.yaml-like (format_data.py:25-58): parallel beam, tilt 29 degrees, 256 x 356 detector, the angle grid of
    data/angles_real.npy (0.72 + 0.96 k degrees)."""
    return dict(DSD=1500.0, DSO=1000.0, nDetector=[det_w, det_h], dDetector=[1.0, 1.0], nVoxel=[356, 356, 70], dVoxel=[1.0, 1.0, 1.0],
                offOrigin=[0, 0, 0], offDetector=[0, 0], accuracy=0.5, mode="parallel", filter=None, tilt_angle=29,
                angles=np.deg2rad(0.72 + 0.96 * 15 * np.arange(n_proj)))


def test_lamino_train_steps_with_ptycho_mask_vs_oracle():
    """BASELINE config 3 through the engine: detector PIXELS of a tilted parallel-beam scan (rays generated in-kernel,
    tigre.py:463-528), projections = phase of a complex full_proj, mask = get_ptycho_mask(full_proj, 0.007) (util.py:196-205,
    train.py:59-60,93-95), pixels drawn on the device among the non-zero ones (tigre.py:354-382); five optimisation steps
    (graph replay) against five oracle steps (CPU autograd + torch Adam) on the rays the reference's dataset code generates."""
    rng = np.random.default_rng(12)
    data = _lamino_geometry()
    geo = G.ConeGeometry(data)
    P, H, W = len(data["angles"]), int(geo.nDetector[1]), int(geo.nDetector[0])
    rays_all = G.rays_with_near_far(data["angles"], geo, "cpu")                        # [P,H,W,8], the reference's generator restated
    assert rays_all.shape == (P, H, W, 8)
    # synthetic complex projections: phase = smooth blobs, amplitude with a low-signal frame and speckles so that the mask bites
    yy, xx = np.mgrid[0:H, 0:W]
    phase = np.stack([0.04 * np.exp(-(((xx - W / 2 - 20 * np.cos(a)) / 70.0) ** 2 + ((yy - H / 2) / 50.0) ** 2)) for a in data["angles"]])
    phase[:, :3, :] = 0.0                                                               # zero projection values are never drawn
    amp = np.where((np.abs(xx - W / 2) < 150) & (np.abs(yy - H / 2) < 110), 1.0, 0.004) * np.ones((P, 1, 1))
    amp = amp * np.where(rng.uniform(size=(P, H, W)) < 0.03, 0.003, 1.0)
    full_proj = torch.from_numpy((amp * np.exp(1j * phase)).astype(np.complex64)).to(DEV)
    projs = torch.from_numpy(phase.astype(np.float32)).to(DEV)
    sampler = PixelSampler(projs, full_proj, threshold=0.007)
    # the mask is the oracle's (and the reference's) mask, bit for bit
    for p in (0, P - 1):
        assert np.array_equal(sampler.mask[p].cpu().numpy().astype(bool), naf.ptycho_mask(full_proj[p].cpu(), 0.007).numpy())
    kept = float(sampler.mask.float().mean())
    assert 0.3 < kept < 0.9, kept

    N, S = 1024, 192
    net = _net(table_scale=0.05)
    o = _oracle_of(net)
    init = [p.detach().cpu().clone() for p in net.parameters()]
    opt = torch.optim.Adam(o.parameters(), lr=1e-3, betas=(0.9, 0.999))
    eng = NAFEngine(net, lr=1e-3, n_samples=S, perturb=True, loss_chunk=200, use_cuda_graph=True)
    eng.set_geometry(data["angles"], geo)
    gen = torch.Generator(device=DEV).manual_seed(3)
    for it in range(5):
        proj = int(rng.integers(0, P))
        pix, pj, mk = sampler.draw(proj, N, generator=gen)
        assert bool((pj != 0).all()) and int(mk.sum()) < N                              # valid pixels only; some are masked out
        t_rand = torch.from_numpy(rng.uniform(0, 1, (N, S)).astype(np.float32))
        l = eng.train_step(None, pj, mk, t_rand.to(DEV), pixels=pix)
        pc = pix.cpu().long()
        rays = rays_all[pc[:, 0], pc[:, 1], pc[:, 2]]
        lo = naf.train_step(o, opt, rays, pj.cpu(), S, True, mask=mk.cpu().bool(), chunk=200, t_rand=t_rand)
        np.testing.assert_allclose(l.item(), lo.item(), rtol=2e-4)
    # Parameters after five Adam steps.  Adam's first updates are lr * sign-like (m / sqrt(v) ~ +-1), so an entry whose gradient is
    # rounding noise (rays clamped onto the volume boundary cancel almost exactly in this geometry) may move the other way: bound
    # the FRACTION of such entries and the relative L2 error of the update instead of the worst entry.
    for a, b, a0 in zip(net.parameters(), o.parameters(), init):
        ua, ub = a.detach().cpu() - a0, b.detach() - a0
        assert float(((ua - ub).abs() > 2e-4).float().mean()) < 2e-3
        assert float((ua - ub).double().norm() / ub.double().norm()) < 0.05
    eng.check_health()


# ----------------------------------------------------------------------------------------------------- chest_50 quality
QUALITY_ITERS = 2000


def test_chest50_full_shape_quality_vs_the_reference_cuda_build():
    """BASELINE config 2 + north_star's quality bar.  config/chest_50.yaml at full shape (128^3 volume, 50 projections of
    256 x 256, 1024 rays x 192 samples, lr 1e-3, 16 x 2 hash grid with 2^19 tables, 4 x 32 MLP) on an analytic phantom:
    the engine in PRODUCTION mode -- detector pixels drawn on the device, rays generated in-kernel, in-kernel uniforms, CUDA graph,
    loss in the forward launch, dense Adam kernel -- and the reference's CUDA build (render + HashEncoder extension +
    calc_mse_loss + torch.optim.Adam, its own torch.rand) train for 2 000 iterations from the same initialisation on the same
    pixel batches.  PSNR-3D within 0.1 dB, SSIM-3D within 0.005 (train.py:220-258, util.py:55-139)."""
    r_get_encoder, r_get_network, r_render, r_calc_mse_loss = _reference()
    data = G.chest50_like(128, 256, 50)
    geo = G.ConeGeometry(data)
    ells = PH.default_ellipsoids(float(geo.sVoxel[0]) / 2)
    vol_gt = torch.from_numpy(PH.phantom_volume(geo, ells)).to(DEV)
    rays_all = G.rays_with_near_far(data["angles"], geo, DEV)                           # [50,256,256,8] (the reference keeps this resident)
    projs_all = PH.phantom_projections(rays_all, ells)
    P, H, W = projs_all.shape
    N, S = 1024, 192
    net = _net(seed=3)                                                                  # reference initialisation: U(-1e-4, 1e-4) tables
    r_net = _reference_twin(net, r_get_encoder, r_get_network)
    r_opt = torch.optim.Adam(r_net.parameters(), lr=1e-3, betas=(0.9, 0.999))
    eng = NAFEngine(net, lr=1e-3, n_samples=S, perturb=True, loss_chunk=None, use_cuda_graph=True, seed=17)
    eng.set_geometry(data["angles"], geo)
    sampler = PixelSampler(projs_all)
    gen = torch.Generator(device=DEV).manual_seed(5)
    torch.manual_seed(23)                                                               # the reference's torch.rand stream
    rng = np.random.default_rng(5)
    first = last = None
    for it in range(QUALITY_ITERS):
        proj = int(rng.integers(0, P))
        pix, pj, mk = sampler.draw(proj, N, generator=gen)
        lc = eng.train_step(None, pj, None, None, pixels=pix)                           # production path
        pl = pix.long()
        rays = rays_all[pl[:, 0], pl[:, 1], pl[:, 2]]
        r_opt.zero_grad()
        ret = r_render(rays, r_net, None, S, 0, True, 409600, 0.0)
        loss = {"loss": 0.0}
        r_calc_mse_loss(loss, pj, ret["acc"])
        loss["loss"].backward()
        r_opt.step()
        if it == 0:
            first = (float(lc.item()), float(loss["loss"].item()))
    last = (float(lc.item()), float(loss["loss"].item()))
    eng.check_health()
    assert last[0] < 0.05 * first[0] and last[1] < 0.05 * first[1], (first, last)      # both train
    n = [int(v) for v in geo.nVoxel]
    vol_c = eng.voxel_query(n, G.voxel_half_extent(geo))
    with torch.no_grad():
        vox = torch.from_numpy(G.get_voxels(geo).astype(np.float32)).to(DEV)
        vol_r = r_net(vox.reshape(-1, 3)).reshape(n)
    psnr_c, psnr_r = get_psnr_3d(vol_c, vol_gt), get_psnr_3d(vol_r, vol_gt)
    ssim_c, ssim_r = get_ssim_3d(vol_c, vol_gt), get_ssim_3d(vol_r, vol_gt)
    print(f"chest_50 full shape, {QUALITY_ITERS} iterations: engine PSNR-3D {psnr_c:.3f} dB SSIM-3D {ssim_c:.4f} | "
          f"reference CUDA build PSNR-3D {psnr_r:.3f} dB SSIM-3D {ssim_r:.4f} | losses {last}")
    assert psnr_c > 20.0 and psnr_r > 20.0, (psnr_c, psnr_r)
    assert abs(psnr_c - psnr_r) <= 0.1, (psnr_c, psnr_r)
    assert abs(ssim_c - ssim_r) <= 0.005, (ssim_c, ssim_r)
