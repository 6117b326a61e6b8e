"""The dataset's per-iteration work on the device (csrc/select.cu): get_ptycho_mask (reference src/utils/util.py:196-205) and
TIGREDataset.__getitem__'s ray selection (src/dataset/tigre.py:354-382: non-zero pixels only, n_rays without replacement,
np.random.choice(replace=False), projection values gathered) against the reference's fixture, the oracle and the statistics of
uniform sampling without replacement."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import naf

DEV = "cuda"

if torch.cuda.is_available():
    from neuralvolumetricreconstructionformedicalimages_b200.dataset.mask import PixelSampler, get_ptycho_mask


def test_ptycho_mask_vs_reference_fixture_and_oracle(golden):
    fx = golden("geometry.npz")
    m = get_ptycho_mask(torch.from_numpy(fx["mask_in"].copy()).to(DEV), 0.007).cpu().numpy()
    assert np.array_equal(m, fx["mask_out"])                       # the reference's own output (tests/golden/generate_golden.py)
    rng = np.random.default_rng(0)
    for H, W in ((256, 356), (1, 9), (7, 1), (33, 64)):
        amp = np.where(rng.uniform(size=(5, H, W)) < 0.4, rng.uniform(0, 0.007, (5, H, W)), rng.uniform(0.007, 1.0, (5, H, W)))
        hr = torch.from_numpy((amp * np.exp(1j * rng.uniform(-3, 3, amp.shape))).astype(np.complex64))
        got = get_ptycho_mask(hr.to(DEV), 0.007).cpu().numpy()
        ref = np.stack([naf.ptycho_mask(hr[p].clone(), 0.007).numpy() for p in range(5)])
        assert np.array_equal(got, ref), (H, W)
        # magnitudes within an ulp of the threshold: the comparison depends on how |z| is rounded, so the yardstick is the
        # reference's own statement sequence (util.py:196-205) executed by torch ON THE DEVICE, where the reference runs it
        amp[rng.uniform(size=amp.shape) < 0.05] = 0.007
        hr = torch.from_numpy((amp * np.exp(1j * rng.uniform(-3, 3, amp.shape))).astype(np.complex64)).to(DEV)
        got = get_ptycho_mask(hr, 0.007)
        for p in range(5):
            m = torch.abs(hr[p]) < 0.007
            m[1:, :] &= (m[1:, :] == m[:-1, :])
            m[:, 1:] &= (m[:, 1:] == m[:, :-1])
            assert torch.equal(got[p], ~m), (H, W, p)


def test_draw_pixels_contract_and_statistics():
    rng = np.random.default_rng(4)
    P, H, W = 3, 40, 50
    projs = torch.from_numpy(rng.uniform(0.1, 1, (P, H, W)).astype(np.float32))
    projs[1, :10] = 0.0                                            # tigre.py:356: zero pixels are never drawn
    projs[2].view(-1)[::3] = 0.0
    full = torch.from_numpy((rng.uniform(0, 0.02, (P, H, W)) * np.exp(1j * rng.uniform(-3, 3, (P, H, W)))).astype(np.complex64))
    ps = PixelSampler(projs.to(DEV), full.to(DEV), 0.007, seed=11)
    assert ps.n_valid.tolist() == [(projs[p] != 0).sum().item() for p in range(P)]
    for p in range(P):
        assert torch.equal(ps.valid[p].cpu().long(), torch.nonzero(projs[p].reshape(-1) != 0).reshape(-1))
    g = torch.Generator(device=DEV).manual_seed(0)
    N = 500
    pix, pr, mk = ps.draw(1, N, g)
    assert pix.shape == (N, 3) and pix.dtype == torch.int32 and bool((pix[:, 0] == 1).all())
    flat = (pix[:, 1].long() * W + pix[:, 2].long()).cpu()
    assert len(torch.unique(flat)) == N and int(pix[:, 1].min()) >= 10        # no replacement, zero rows excluded
    assert torch.equal(pr.cpu(), projs[1].reshape(-1)[flat]) and torch.equal(mk.cpu(), ps.mask[1].reshape(-1).cpu()[flat])
    with pytest.raises(ValueError):
        ps.draw(1, H * W, g)
    # drawing EVERY valid pixel returns each exactly once
    M = int(ps.n_valid[2])
    pa, _, _ = ps.draw(2, M, g)
    assert torch.equal(torch.sort(pa[:, 1].long() * W + pa[:, 2].long()).values.cpu(), ps.valid[2].cpu().long())
    # the graph-capturable entry: the counter advances on the device, projections follow the DataLoader order k % P
    pixels = torch.empty(64, 3, dtype=torch.int32, device=DEV)
    vals = torch.empty(64, device=DEV)
    msk = torch.empty(64, dtype=torch.uint8, device=DEV)
    seen = []
    for k in range(7):
        ps.draw_into(64, pixels, vals, msk)
        seen.append(int(pixels[0, 0]))
        assert bool((pixels[:, 0] == k % P).all()) and bool((vals != 0).all())
    assert seen == [0, 1, 2, 0, 1, 2, 0] and ps.draws_done() == 7
    ps.check()
    # same seed + same counter -> same draw; another counter -> another draw
    ps2 = PixelSampler(projs.to(DEV), full.to(DEV), 0.007, seed=11)
    a = torch.empty(64, 3, dtype=torch.int32, device=DEV)
    ps2.draw_into(64, a, vals, msk)
    ps.set_draw(0)
    ps.draw_into(64, pixels, vals, msk)
    assert torch.equal(a, pixels)
    ps.draw_into(64, pixels, vals, msk)
    assert not torch.equal(a[:, 1:], pixels[:, 1:])
    # too few valid pixels: the error word is raised on the device and surfaces in check()
    small = PixelSampler(projs[:, :4, :4].contiguous().to(DEV))
    small.draw_into(64, pixels, vals, msk)
    with pytest.raises(ValueError):
        small.check()


def test_draw_pixels_is_uniform_without_replacement():
    """Statistics of np.random.choice(n_valid, n_rays, replace=False) (tigre.py:358): every valid pixel is included with
    probability n/M, every POSITION of the batch is uniform over the valid pixels (the chunked loss of train.py:69-127 weighs
    positions differently, so the order must be random too), pairs of pixels are (very slightly negatively) uncorrelated."""
    rng = np.random.default_rng(1)
    H, W, N, K = 24, 32, 96, 4000
    projs = torch.from_numpy(rng.uniform(0.1, 1, (1, H, W)).astype(np.float32))
    projs[0, ::4] = 0.0
    ps = PixelSampler(projs.to(DEV), seed=5)
    M = int(ps.n_valid[0])
    pixels = torch.empty(K, N, 3, dtype=torch.int32, device=DEV)
    vals = torch.empty(N, device=DEV)
    for k in range(K):
        ps.draw_into(N, pixels[k], vals, None)
    flat = (pixels[..., 1].long() * W + pixels[..., 2].long()).cpu().numpy()            # [K, N]
    valid = set(ps.valid[0].cpu().tolist())
    assert set(np.unique(flat).tolist()) <= valid and all(len(np.unique(r)) == N for r in flat[:200])
    counts = np.bincount(flat.ravel(), minlength=H * W)[sorted(valid)].astype(np.float64)
    p = N / M
    z = (counts - K * p) / np.sqrt(K * p * (1 - p))
    assert abs(z.mean()) < 0.2 and 0.8 < z.std() < 1.2 and np.abs(z).max() < 5.0, (z.mean(), z.std(), np.abs(z).max())
    # position 0 and the last position are uniform over the valid pixels (chi-square against the uniform law)
    for pos in (0, N - 1):
        c = np.bincount(flat[:, pos], minlength=H * W)[sorted(valid)].astype(np.float64)
        chi2 = ((c - K / M) ** 2 / (K / M)).sum()
        assert abs(chi2 - (M - 1)) < 5 * np.sqrt(2 * (M - 1)), (pos, chi2, M)
    # joint inclusion of two fixed pixels: P(both) = n(n-1) / (M(M-1))
    a, b = sorted(valid)[3], sorted(valid)[-7]
    both = np.mean([(a in r) and (b in r) for r in (set(x) for x in flat.tolist())])
    expect = N * (N - 1) / (M * (M - 1))
    assert abs(both - expect) < 5 * np.sqrt(expect * (1 - expect) / K), (both, expect)


def test_metric_kernels_vs_oracle():
    """get_psnr_3d / get_ssim_3d on fp32 CUDA volumes (csrc/metrics.cu; reference src/utils/util.py:55-139, eval_step train.py:253-258)
    against the oracle's float64 numpy / scipy restatements (themselves pinned to the reference's PSNR fixture and to a brute-force
    evaluation of the published SSIM definition in tests/test_host_logic.py)."""
    from neuralvolumetricreconstructionformedicalimages_b200.utils import get_psnr_3d, get_ssim_3d
    rng = np.random.default_rng(3)
    for shape in ((20, 17, 23), (7, 7, 7), (64, 64, 64), (9, 40, 8)):
        v1 = rng.uniform(0, 1, shape).astype(np.float32)
        v2 = np.clip(v1 + rng.normal(0, 0.05, shape), 0, 1).astype(np.float32)
        a, b = torch.from_numpy(v1).to(DEV), torch.from_numpy(v2).to(DEV)
        assert abs(get_psnr_3d(a, b) - naf.psnr_3d(v1, v2)) < 1e-9
        assert abs(get_ssim_3d(a, b) - naf.ssim_3d(v1, v2)) < 1e-10, shape
        assert get_psnr_3d(a, a) == 100.0 and abs(get_ssim_3d(a, a) - 1.0) < 1e-12
        # the torch-op path (what host-side float64 inputs take) agrees with the kernels
        assert abs(get_ssim_3d(a.double(), b.double()) - get_ssim_3d(a, b)) < 1e-10
    with pytest.raises(ValueError):
        get_ssim_3d(torch.zeros(6, 9, 9, device=DEV), torch.zeros(6, 9, 9, device=DEV))


def test_train_step_sampled_equals_draw_then_train_step():
    """NAFEngine.train_step_sampled (draw kernel + forward/loss + backward + optimizer in ONE CUDA graph, no per-step host input)
    performs exactly the steps of `sampler.draw_into(...)` followed by `train_step(pixels=...)` on a twin engine with the same seeds:
    same pixels, same in-kernel uniforms, so losses agree to the order of the float atomics."""
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import phantom as PH
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.engine import NAFEngine
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network
    data = G.chest50_like(n_voxel=32, n_detector=64, n_proj=5)
    geo = G.ConeGeometry(data)
    ells = PH.default_ellipsoids(float(geo.sVoxel[0]) / 2)
    projs = PH.phantom_projections(G.rays_with_near_far(data["angles"], geo, "cpu"), ells).to(DEV)
    # 8x8 BLOCKS of low amplitude: the ptycho rule compares finite differences, an isolated low pixel is not masked
    low = (torch.rand(projs.shape[0], projs.shape[1] // 8, projs.shape[2] // 8, device=DEV) < 0.3)
    low = low.repeat_interleave(8, 1).repeat_interleave(8, 2)
    full = torch.polar(torch.where(low, 0.003, 1.0), projs)

    def engine():
        torch.manual_seed(0)
        enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
        net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid").to(DEV)
        eng = NAFEngine(net, lr=1e-3, n_samples=64, perturb=True, loss_chunk=100, use_cuda_graph=True, seed=7)
        eng.set_geometry(data["angles"], geo)
        return eng

    N = 512
    e1, e2 = engine(), engine()
    s1, s2 = PixelSampler(projs, full, 0.007, seed=3), PixelSampler(projs, full, 0.007, seed=3)
    pix = torch.empty(N, 3, dtype=torch.int32, device=DEV)
    val = torch.empty(N, device=DEV)
    msk = torch.empty(N, dtype=torch.uint8, device=DEV)
    for k in range(7):                                    # eager first, then captured + replayed
        l1 = float(e1.train_step_sampled(s1, N))
        s2.draw_into(N, pix, val, msk)
        assert bool((pix[:, 0] == k % 5).all()) and 0 < int(msk.sum()) < N
        l2 = float(e2.train_step(None, val.clone(), msk.clone(), pixels=pix.clone()))
        np.testing.assert_allclose(l1, l2, rtol=1e-5)
    assert s1.draws_done() == 8 and e1.step_count == 7          # one draw ahead: step k draws the batch of step k + 1 beside its optimizer
    # somebody else moves the sampler: the prefetched batch is dropped, the next step draws afresh -- here draw 3 again, on both sides
    s1.set_draw(3)
    s2.set_draw(3)
    l1 = float(e1.train_step_sampled(s1, N))
    s2.draw_into(N, pix, val, msk)
    assert bool((pix[:, 0] == 3).all())
    l2 = float(e2.train_step(None, val.clone(), msk.clone(), pixels=pix.clone()))
    np.testing.assert_allclose(l1, l2, rtol=1e-5)
    assert s1.draws_done() == 5
    s1.check()
    e1.check_health()
    assert float((e1.flat_param - e2.flat_param).abs().max()) < 5e-5
