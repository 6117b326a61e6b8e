"""Generate the committed golden fixtures (tests/golden/*.npz) by running the REFERENCE.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/generate_golden.py

What "the reference" means here:
  * Python half -- the unmodified modules under /root/reference/src (render, network,
    encoder.freqencoder, loss, dataset.tigre, utils.util) imported with sys.modules stubs
    for the packages this image lacks (matplotlib, open3d, skimage, imageio).
  * CUDA op -- the reference's kernel text (hashencoder.cu:30-298) compiled for the host
    by oracle/build_ref.sh (oracle/_ref/libref_hashgrid.so).  src.encoder.hashencoder is
    replaced by a module whose HashEncoder calls that library, because the reference's
    own backend.py JIT-loads a CUDA extension that cannot run without a GPU.

The fixtures are inputs + reference outputs; big hash tables are NOT stored, they are a
closed-form function of the entry index (formula_table) so tests rebuild them exactly.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

sys.path.insert(0, os.path.dirname(HERE))
from helpers import formula_table, make_rays  # noqa: E402
from oracle import hashgrid as oh  # noqa: E402

REF = "/root/reference"


def install_reference():
    """Import the reference's python packages with stubs for absent third-party modules."""
    for name in ["matplotlib", "matplotlib.pyplot", "open3d", "skimage", "skimage.metrics", "imageio", "imageio.v2"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage.metrics"].structural_similarity = lambda *a, **k: float("nan")
    sys.modules["skimage"].metrics = sys.modules["skimage.metrics"]
    sys.modules["imageio"].v2 = sys.modules["imageio.v2"]

    # src.encoder.hashencoder -> backed by the reference kernel text on the host
    stub = types.ModuleType("src.encoder.hashencoder")

    class HashEncoder(oh.OracleHashEncoder):
        def __init__(self, input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19):
            # hashgrid.py:125 `(inputs + size) / (2 * size)` is evaluated by ATen's CUDA div kernel as a
            # multiply by fl(1 / fl(2*size)) (python-scalar divisor); the reference only ever runs this
            # module on CUDA, so the fixtures follow that evaluation (checked against torch on the
            # GPU by tests/test_gpu_parity.py::test_aten_cuda_assumptions).
            super().__init__(input_dim, num_levels, level_dim, base_resolution, log2_hashmap_size, use_ref=True, normalise="mul_recip")

    stub.HashEncoder = HashEncoder
    sys.modules["src.encoder.hashencoder"] = stub
    sys.path.insert(0, REF)
    import src.encoder  # noqa: F401
    import src.loss  # noqa: F401
    import src.network  # noqa: F401
    import src.render  # noqa: F401


def gen_hash_fixtures():
    assert oh.have_ref(), "run oracle/build_ref.sh first"
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- KAT: index table of SURVEY.md 8c re-derived from the reference text
    kat_pts = [(0, (0, 0, 0)), (0, (16, 16, 16)), (0, (3, 5, 7)), (1, (32, 32, 32)), (2, (10, 20, 30)), (3, (1, 2, 3)),
               (3, (128, 128, 128)), (7, (1000, 2000, 2047)), (11, (32768, 1, 2)), (12, (1, 2, 3)),
               (12, (65536, 65536, 65536)), (12, (40000, 50000, 60000)), (13, (1, 2, 3)), (13, (131072, 131071, 77777)),
               (14, (5, 6, 7)), (15, (524287, 524288, 1)), (15, (262144, 131072, 65536))]
    offs = oh.level_offsets(16, 16, 19, 3)
    rows = []
    for lvl, p in kat_pts:
        T = int(offs[lvl + 1] - offs[lvl])
        res = 16 * 2 ** lvl
        rows.append([lvl, res, T, p[0], p[1], p[2], oh.ref_grid_index_3(2, T, res, p)])
    # plus 400 random lattice points per level
    for lvl in range(16):
        T = int(offs[lvl + 1] - offs[lvl])
        res = 16 * 2 ** lvl
        for _ in range(25):
            p = [int(v) for v in rng.integers(0, res + 1, 3)]
            rows.append([lvl, res, T, p[0], p[1], p[2], oh.ref_grid_index_3(2, T, res, p)])
    out["kat"] = np.asarray(rows, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "hash_kat.npz"), **out)

    # ---- full chest config: 16 x 2, 2^19, base 16, D = 3
    L, C, D, H = 16, 2, 3, 16
    table = formula_table(int(offs[-1]), C, 1.0)
    B = 384
    x = rng.uniform(0, 1, (B, D)).astype(np.float32)
    x[0] = 0.0
    x[1] = 1.0
    x[2] = [0.3333333, 0.5, 1.0]
    x[3] = [np.nextafter(np.float32(1), np.float32(0)), 0.0, 0.5]
    # a coherent ray segment (consecutive samples share coarse cells)
    t = np.linspace(0, 1, 128, dtype=np.float32)[:, None]
    x[128:256] = (np.float32([0.1, 0.2, 0.3]) * (1 - t) + np.float32([0.9, 0.7, 0.35]) * t).astype(np.float32)
    y, dy_dx = oh.ref_hash_forward(x, table, offs, H, calc_grad_inputs=True)
    Bb = 128
    g = rng.normal(size=(Bb, L * C)).astype(np.float32)
    gg = oh.ref_hash_backward(g, x[128:256], table, offs, H, ordered=True)
    nz = np.flatnonzero(np.any(gg != 0, axis=1))
    np.savez_compressed(os.path.join(HERE, "hash_chest.npz"), x=x, out_LBC=y,
                        dy_dx_last=dy_dx.reshape(B, L, D, C)[:, :, D - 1, :].copy(),
                        grad=g, grad_rows=nz.astype(np.int64), grad_vals=gg[nz], cfg=np.asarray([L, C, D, H, 19]))

    # ---- small configs covering D=2 and every C
    small = {}
    for tag, (D, C, L, H, log2T) in {"d2c4": (2, 4, 8, 8, 10), "d3c1": (3, 1, 5, 4, 9), "d3c8": (3, 8, 4, 16, 12), "d2c2": (2, 2, 12, 16, 15)}.items():
        offs_s = oh.level_offsets(L, H, log2T, D)
        tab = formula_table(int(offs_s[-1]), C, 1.0)
        B = 200
        xs = rng.uniform(0, 1, (B, D)).astype(np.float32)
        xs[0] = 0
        xs[1] = 1
        ys, _ = oh.ref_hash_forward(xs, tab, offs_s, H)
        gs = rng.normal(size=(B, L * C)).astype(np.float32)
        ggs = oh.ref_hash_backward(gs, xs, tab, offs_s, H, ordered=True)
        small[f"{tag}_cfg"] = np.asarray([L, C, D, H, log2T])
        small[f"{tag}_x"] = xs
        small[f"{tag}_out"] = ys
        small[f"{tag}_grad"] = gs
        small[f"{tag}_gtab"] = ggs
    np.savez_compressed(os.path.join(HERE, "hash_small.npz"), **small)


def _seed_mlp(net, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for lin in net.layers:
            lin.weight.copy_((torch.rand(lin.weight.shape, generator=g) - 0.5) * (2.0 / np.sqrt(lin.weight.shape[1])))
            lin.bias.copy_((torch.rand(lin.bias.shape, generator=g) - 0.5) * 0.2)


def _mlp_arrays(net, prefix=""):
    d = {}
    for i, lin in enumerate(net.layers):
        d[f"{prefix}W{i}"] = lin.weight.detach().numpy().copy()
        d[f"{prefix}b{i}"] = lin.bias.detach().numpy().copy()
    return d


def gen_render_fixtures():
    import src.render  # noqa: F401  (the package attribute `render` is the function; take the module)
    rr = sys.modules["src.render.render"]
    from src.encoder import get_encoder
    from src.loss import calc_mse_loss
    from src.network import get_network

    rng = np.random.default_rng(7)
    out = {}
    # ---- (a) frequency encoder network, reference render(), perturb False / True
    enc = get_encoder("frequency", multires=6)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
    _seed_mlp(net, 1)
    N, S = 48, 40
    rays = torch.from_numpy(make_rays(N, rng))
    t_rand = torch.from_numpy(rng.uniform(0, 1, (N, S)).astype(np.float32))
    ret0 = rr.render(rays, net, None, S, 0, False, 409600, 0.0)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: t_rand.clone()
    try:
        ret1 = rr.render(rays, net, None, S, 0, True, 409600, 0.0)
    finally:
        torch.rand = real_rand
    out.update(freq_rays=rays.numpy(), freq_t_rand=t_rand.numpy(), **_mlp_arrays(net, "freq_"))
    out.update(freq_acc_noperturb=ret0["acc"].detach().numpy(), freq_pts_noperturb=ret0["pts"].detach().numpy(),
               freq_tv_noperturb=ret0["tv_loss"].detach().numpy(),
               freq_acc_perturb=ret1["acc"].detach().numpy(), freq_pts_perturb=ret1["pts"].detach().numpy(),
               freq_tv_perturb=ret1["tv_loss"].detach().numpy())

    # ---- (b) chest config: hash 16x2 2^19 + 4x32 skips [2] sigmoid, bound 0.3 (config/chest_50.yaml)
    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
    _seed_mlp(net, 2)
    with torch.no_grad():
        net.encoder.embeddings.copy_(torch.from_numpy(formula_table(net.encoder.embeddings.shape[0], 2, 0.5)))
    N, S = 40, 24
    rays = torch.from_numpy(make_rays(N, rng))
    projs = torch.from_numpy(rng.uniform(0, 0.05, N).astype(np.float32))
    t_rand = torch.from_numpy(rng.uniform(0, 1, (N, S)).astype(np.float32))
    torch.rand = lambda *a, **k: t_rand.clone()
    try:
        ret = rr.render(rays, net, None, S, 0, True, 409600, 0.0)
    finally:
        torch.rand = real_rand
    loss = {"loss": 0.0}
    calc_mse_loss(loss, projs, ret["acc"])
    loss["loss"].backward()
    gt = net.encoder.embeddings.grad.numpy()
    nz = np.flatnonzero(np.any(gt != 0, axis=1))
    out.update(chest_rays=rays.numpy(), chest_projs=projs.numpy(), chest_t_rand=t_rand.numpy(), **_mlp_arrays(net, "chest_"))
    out.update(chest_acc=ret["acc"].detach().numpy(), chest_pts=ret["pts"].detach().numpy(), chest_loss=loss["loss"].detach().numpy(),
               chest_gtab_rows=nz.astype(np.int64), chest_gtab_vals=gt[nz])
    for i, lin in enumerate(net.layers):
        out[f"chest_gW{i}"] = lin.weight.grad.numpy().copy()
        out[f"chest_gb{i}"] = lin.bias.grad.numpy().copy()
    # raw densities + the other heads on the same points (forward only)
    with torch.no_grad():
        pts = ret["pts"].reshape(-1, 3)
        out["chest_sigma"] = net(pts).numpy()
        for head in ["relu", "tanh", "none"]:
            n2 = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation=head)
            n2.load_state_dict(net.state_dict())
            out[f"chest_sigma_{head}"] = n2(pts).numpy()
        # a deeper net with two skips exercises the generic layer loop (network.py:40-51)
        n3 = get_network("mlp")(enc, bound=0.3, num_layers=6, hidden_dim=32, skips=[2, 4], out_dim=1, last_activation="sigmoid")
        _seed_mlp(n3, 3)
        out.update(_mlp_arrays(n3, "deep_"))
        out["deep_sigma"] = n3(pts).numpy()
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out)


def gen_geometry_fixtures():
    from src.dataset.tigre import ConeGeometry, TIGREDataset
    from src.loss import calc_mse_loss
    from src.utils.util import get_psnr_3d, get_ptycho_mask

    rng = np.random.default_rng(11)
    out = {}
    ds = object.__new__(TIGREDataset)
    base = dict(DSD=1500.0, DSO=1000.0, nDetector=[10, 6], dDetector=[1.5, 2.0], nVoxel=[8, 6, 4], dVoxel=[1.0, 2.0, 1.5],
                offOrigin=[0, 0, 0], offDetector=[0.5, -1.0], accuracy=0.5, filter=None)
    angles = np.asarray([0.0, 0.3, 1.7, 3.0, np.deg2rad(0.72), np.deg2rad(179.28)])
    out["angles"] = angles
    for mode, tilt in [("cone", 0), ("parallel", 29), ("parallel", 0), ("cone", 10)]:
        geo = ConeGeometry(dict(base, mode=mode, tilt_angle=tilt))
        tag = f"{mode}_t{tilt}"
        out[f"poses_{tag}"] = np.stack([ds.angle2pose(geo.DSO, a, tilt) for a in angles])
        out[f"rays_{tag}"] = ds.get_rays(angles, geo, "cpu").numpy()
        if mode == "parallel":
            r2 = ds.get_rays2(angles, geo, "cpu").numpy()
            assert np.array_equal(r2, out[f"rays_{tag}"]), "get_rays vs get_rays2 differ"
    geo = ConeGeometry(dict(base, mode="cone", tilt_angle=0))
    out["near_far"] = np.asarray(ds.get_near_far(geo))
    out["voxels"] = ds.get_voxels(geo)
    chest = ConeGeometry(dict(DSD=1500.0, DSO=1000.0, nDetector=[256, 256], dDetector=[1.0, 1.0], nVoxel=[128, 128, 128],
                              dVoxel=[1.0, 1.0, 1.0], offOrigin=[0, 0, 0], offDetector=[0, 0], accuracy=0.5, mode="cone", filter=None))
    out["near_far_chest"] = np.asarray(ds.get_near_far(chest))
    # ptycho mask
    hr = (rng.normal(size=(12, 17)) * 0.01 + 1j * rng.normal(size=(12, 17)) * 0.01).astype(np.complex64)
    hr[3:7, 4:9] = 0
    out["mask_in"] = hr
    out["mask_out"] = get_ptycho_mask(torch.from_numpy(hr), 0.007).numpy()
    # chunked masked MSE exactly as train.py:69-127 intends it
    n = 53
    pred = torch.from_numpy(rng.uniform(0, 1, n).astype(np.float32))
    tgt = torch.from_numpy(rng.uniform(0, 1, n).astype(np.float32))
    m = torch.from_numpy(rng.uniform(0, 1, n) > 0.3)
    loss = {"loss": 0.0}
    for i in range(0, n, 20):
        calc_mse_loss(loss, tgt[i:i + 20][m[i:i + 20]], pred[i:i + 20][m[i:i + 20]])
    out.update(mse_pred=pred.numpy(), mse_tgt=tgt.numpy(), mse_mask=m.numpy(), mse_chunk20=loss["loss"].numpy())
    a = rng.uniform(0, 1, (5, 6, 7)).astype(np.float32)
    b = (a + rng.normal(size=a.shape) * 0.05).astype(np.float32)
    out.update(psnr_a=a, psnr_b=b, psnr_3d=np.asarray(get_psnr_3d(a, b)))
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **out)


def gen_fine_fixtures():
    """render() with a fine network (n_fine > 0), raw_noise_std > 0 and a two-channel coarse head (render.py:113-126, :196-199,
    :207-208, sample_pdf :215-247) on the frequency-encoder network: inputs, generator seed and the reference's outputs."""
    import src.render  # noqa: F401
    rr = sys.modules["src.render.render"]
    from src.encoder import get_encoder
    from src.network import get_network

    rng = np.random.default_rng(17)
    N, S, NF = 24, 16, 8
    rays = torch.from_numpy(make_rays(N, rng))
    out = dict(rays=rays.numpy(), n_samples=S, n_fine=NF, seed=123)
    enc = get_encoder("frequency", multires=6)
    fine = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
    _seed_mlp(fine, 12)
    out.update(_mlp_arrays(fine, "fine_"))
    for out_dim in (1, 2):
        net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=out_dim, last_activation="sigmoid")
        _seed_mlp(net, 10 + out_dim)
        out.update(_mlp_arrays(net, f"coarse{out_dim}_"))
        for perturb, noise in ((False, 0.0), (True, 0.0), (True, 0.3)):
            torch.manual_seed(123)
            ret = rr.render(rays, net, fine, S, NF, perturb, 409600, noise)
            tag = f"od{out_dim}_p{int(perturb)}_n{int(noise > 0)}_"
            for k in ("acc", "pts", "tv_loss", "acc0", "weights0", "pts0"):
                out[tag + k] = ret[k].detach().numpy()
    np.savez_compressed(os.path.join(HERE, "render_fine.npz"), **out)


def gen_signatures():
    """The operator-API surface the drop-in must mirror (names, order, defaults)."""
    import inspect
    import json

    import src.render  # noqa: F401
    rr = sys.modules["src.render.render"]
    from src.encoder import get_encoder
    from src.encoder.freqencoder import FreqEncoder
    from src.loss import calc_mse_loss
    from src.network import get_network
    from src.network.network import DensityNetwork

    sig = {
        "render": str(inspect.signature(rr.render)),
        "run_network": str(inspect.signature(rr.run_network)),
        "raw2outputs": str(inspect.signature(rr.raw2outputs)),
        "sample_pdf": str(inspect.signature(rr.sample_pdf)),
        "get_encoder": str(inspect.signature(get_encoder)),
        "get_network": str(inspect.signature(get_network)),
        "DensityNetwork.__init__": str(inspect.signature(DensityNetwork.__init__)),
        "DensityNetwork.forward": str(inspect.signature(DensityNetwork.forward)),
        "FreqEncoder.__init__": str(inspect.signature(FreqEncoder.__init__)),
        "calc_mse_loss": str(inspect.signature(calc_mse_loss)),
        # hashgrid.py:78,118 (not importable without a GPU build; transcribed)
        "HashEncoder.__init__": "(self, input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)",
        "HashEncoder.forward": "(self, inputs, size=1)",
    }
    with open(os.path.join(HERE, "signatures.json"), "w") as f:
        json.dump(sig, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    oh.build_oracle()
    oh.build_ref()
    if sys.argv[1:] == ["fine"]:      # only the hierarchical-sampling fixture (added later; the others stay as committed)
        install_reference()
        gen_fine_fixtures()
        sys.exit(0)
    gen_hash_fixtures()
    install_reference()
    gen_render_fixtures()
    gen_fine_fixtures()
    gen_geometry_fixtures()
    gen_signatures()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
