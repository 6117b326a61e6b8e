"""CPU tests of the host-side logic of the product package (no kernel launches):
geometry vs the reference fixtures, API surface vs the reference signatures, C-ABI exports."""
import ctypes
import importlib.util
import inspect
import json
import os
import re

import numpy as np
import pytest
import torch

import neuralvolumetricreconstructionformedicalimages_b200 as pkg
from neuralvolumetricreconstructionformedicalimages_b200 import _lib
from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_GEO = dict(DSD=1500.0, DSO=1000.0, nDetector=[10, 6], dDetector=[1.5, 2.0], nVoxel=[8, 6, 4], dVoxel=[1.0, 2.0, 1.5],
                offOrigin=[0, 0, 0], offDetector=[0.5, -1.0], accuracy=0.5, filter=None)


@pytest.mark.parametrize("mode,tilt", [("cone", 0), ("parallel", 29), ("parallel", 0), ("cone", 10)])
def test_rays_bit_identical_to_reference(golden, mode, tilt):
    fx = golden("geometry.npz")
    geo = G.ConeGeometry(dict(BASE_GEO, mode=mode, tilt_angle=tilt))
    tag = f"{mode}_t{tilt}"
    poses = np.stack([G.angle2pose(geo.DSO, a, tilt) for a in fx["angles"]])
    np.testing.assert_allclose(poses, fx[f"poses_{tag}"], rtol=0, atol=1e-15)
    rays = G.get_rays(fx["angles"], geo).numpy()
    assert np.array_equal(rays.view(np.uint32), fx[f"rays_{tag}"].view(np.uint32))


def test_near_far_voxels(golden):
    fx = golden("geometry.npz")
    geo = G.ConeGeometry(dict(BASE_GEO, mode="cone"))
    assert np.array_equal(np.asarray(G.get_near_far(geo)), fx["near_far"])
    assert np.array_equal(G.get_voxels(geo), fx["voxels"])
    chest = G.ConeGeometry(G.chest50_like())
    assert np.array_equal(np.asarray(G.get_near_far(chest)), fx["near_far_chest"])
    r = G.rays_with_near_far([0.0, 1.0], G.ConeGeometry(dict(BASE_GEO, mode="cone")))
    assert r.shape == (2, 6, 10, 8)


def test_operator_signatures_match_reference():
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.encoder.freqencoder import FreqEncoder
    from neuralvolumetricreconstructionformedicalimages_b200.encoder.hashgrid import HashEncoder
    from neuralvolumetricreconstructionformedicalimages_b200.loss import calc_mse_loss
    from neuralvolumetricreconstructionformedicalimages_b200.network import DensityNetwork, get_network
    from neuralvolumetricreconstructionformedicalimages_b200.render import raw2outputs, render, run_network, sample_pdf

    with open(os.path.join(ROOT, "tests", "golden", "signatures.json")) as f:
        ref = json.load(f)
    ours = {"render": render, "run_network": run_network, "raw2outputs": raw2outputs, "sample_pdf": sample_pdf, "get_encoder": get_encoder,
            "get_network": get_network, "DensityNetwork.__init__": DensityNetwork.__init__, "DensityNetwork.forward": DensityNetwork.forward,
            "FreqEncoder.__init__": FreqEncoder.__init__, "calc_mse_loss": calc_mse_loss, "HashEncoder.__init__": HashEncoder.__init__,
            "HashEncoder.forward": HashEncoder.forward}
    strip = lambda s: re.sub(r" at 0x[0-9a-f]+", "", s)  # noqa: E731
    for name, fn in ours.items():
        assert strip(str(inspect.signature(fn))) == strip(ref[name]), name


def test_module_surface_and_state_dict():
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network

    enc = get_encoder("hashgrid", input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19)
    assert enc.output_dim == 32 and enc.offsets.dtype == torch.int32 and int(enc.offsets[-1]) == 7131219
    assert float(enc.embeddings.abs().max()) <= 1e-4
    net = get_network("mlp")(enc, bound=0.3, num_layers=4, hidden_dim=32, skips=[2], out_dim=1, last_activation="sigmoid")
    sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "encoder.embeddings": (7131219, 2), "layers.0.weight": (32, 32), "layers.0.bias": (32,), "layers.1.weight": (32, 32),
        "layers.1.bias": (32,), "layers.2.weight": (32, 64), "layers.2.bias": (32,), "layers.3.weight": (1, 32), "layers.3.bias": (1,)}
    assert net.fused_meta() is not None and net.in_dim == 32 and net.bound == 0.3
    fe = get_encoder("frequency", multires=6)
    assert fe.output_dim == 39
    assert get_network("mlp")(fe, num_layers=4, hidden_dim=32, skips=[2]).fused_meta() is None
    with pytest.raises(NotImplementedError):
        get_encoder("sphere")
    with pytest.raises(NotImplementedError):
        get_network("cnn")
    with pytest.raises(NotImplementedError):
        get_network("mlp")(fe, last_activation="softplus")
    ident, dim = get_encoder("None", input_dim=3)
    assert dim == 3 and ident(5) == 5


def test_no_cpu_fallback():
    """The product must fail loudly off-GPU instead of computing on the host."""
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.loss import calc_mse_loss
    from neuralvolumetricreconstructionformedicalimages_b200.render import render

    enc = get_encoder("hashgrid", num_levels=2, log2_hashmap_size=8)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        enc(torch.zeros(4, 3), 1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        calc_mse_loss({"loss": 0.0}, torch.zeros(3), torch.zeros(3))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        render(torch.zeros(4, 8), None, None, 8, 0, False, 1024, 0.0)
    # and nothing under the package imports the oracle
    for dirpath, _, files in os.walk(os.path.dirname(pkg.__file__)):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_c_abi_exports_every_declared_symbol():
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "nafb200.h")).read()
    declared = sorted(set(re.findall(r"\b(nafb_[a-z0-9_]+)\s*\(", header)))
    assert declared == _lib.exported_symbols()
    for name in declared:
        assert hasattr(L, name), name
    assert L.nafb_abi_version() == 10
    # the diagnostics library exports what include/nafb200_diag.h declares, and the product library does not carry them
    dh = open(os.path.join(ROOT, "include", "nafb200_diag.h")).read()
    D = _lib.diag_lib()
    for name in sorted(set(re.findall(r"\b(nafb_[a-z0-9_]+)\s*\(", dh))):
        assert hasattr(D, name), name
        assert not hasattr(L, name), name
    # struct layouts agree with the C header (sizes computed by the C compiler)
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "nafb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(nafb_grid), sizeof(nafb_mlp), sizeof(nafb_mlp_grads), sizeof(nafb_sampler), sizeof(nafb_exchange), sizeof(nafb_loss_tail));}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [ctypes.sizeof(_lib.Grid), ctypes.sizeof(_lib.Mlp), ctypes.sizeof(_lib.MlpGrads), ctypes.sizeof(_lib.Sampler), ctypes.sizeof(_lib.Exchange), ctypes.sizeof(_lib.LossTail)]
    # argument validation happens before any launch: safe to exercise without a GPU
    assert L.nafb_adam_step(None, None, None, None, 4, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, 1, None) == _lib.ERR_INVALID
    assert b"null pointer" in L.nafb_last_error()
    offs = np.asarray([0, 8, 16], np.int32)
    g = _lib.Grid(1, offs.ctypes.data, 3, 3, 2, 4)
    assert L.nafb_hash_encode_forward(ctypes.byref(g), 1, 1, 4, 0, 0, None, None) == _lib.ERR_UNSUPPORTED
    assert L.nafb_last_error() == b"GridEncoding: C must be 1, 2, 4, or 8."


def test_pixel_sampler_and_mask_need_a_gpu():
    """The mask / pixel selection run in CUDA kernels (csrc/select.cu); like every operator of the package they refuse CPU
    tensors instead of falling back.  (The oracle's restatement of the mask is pinned to the reference's output in
    test_oracle_golden.py; the kernels are compared with it in tests/test_gpu_select.py.)"""
    from neuralvolumetricreconstructionformedicalimages_b200.dataset.mask import PixelSampler, get_ptycho_mask
    with pytest.raises(RuntimeError):
        get_ptycho_mask(torch.zeros(4, 4, dtype=torch.complex64), 0.007)
    with pytest.raises(RuntimeError):
        PixelSampler(torch.ones(2, 4, 4))


def test_pose_table_and_detector_fields(golden):
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G
    fx = golden("geometry.npz")
    base = dict(DSD=1500.0, DSO=1000.0, nDetector=[10, 6], dDetector=[1.5, 2.0], nVoxel=[8, 6, 4], dVoxel=[1.0, 2.0, 1.5],
                offOrigin=[0, 0, 0], offDetector=[0.5, -1.0], accuracy=0.5, filter=None)
    geo = G.ConeGeometry(dict(base, mode="parallel", tilt_angle=29))
    T = G.pose_table(fx["angles"], geo).numpy()
    ref = fx["poses_parallel_t29"].astype(np.float32)
    assert np.array_equal(T[:, :9].reshape(-1, 3, 3), ref[:, :3, :3]) and np.array_equal(T[:, 9:], ref[:, :3, 3])
    d = G.detector_fields(geo)
    assert d["det_w"] == 10 and d["det_h"] == 6 and d["det_parallel"] == 1 and d["det_du"] == float(np.float32(0.0015))


def test_metrics_match_reference_fixtures_and_oracle(golden):
    """utils.metrics against the reference's get_psnr_3d output (fixture) and the scipy SSIM restatement of the oracle."""
    from neuralvolumetricreconstructionformedicalimages_b200.utils import get_mse, get_psnr, get_psnr_3d, get_ssim_3d
    from oracle import naf
    fx = golden("geometry.npz")
    a, b = torch.from_numpy(fx["psnr_a"]), torch.from_numpy(fx["psnr_b"])
    assert abs(get_psnr_3d(a, b) - float(fx["psnr_3d"])) < 1e-9
    assert get_psnr_3d(a, a) == 100.0
    rng = np.random.default_rng(3)
    v1 = rng.uniform(0, 1, (20, 17, 23)).astype(np.float32)
    v2 = np.clip(v1 + rng.normal(0, 0.05, v1.shape), 0, 1).astype(np.float32)
    s_ours, s_ref = get_ssim_3d(torch.from_numpy(v1), torch.from_numpy(v2)), naf.ssim_3d(v1, v2)
    assert abs(s_ours - s_ref) < 1e-12 and 0.5 < s_ours < 1.0
    assert abs(get_ssim_3d(torch.from_numpy(v1), torch.from_numpy(v1)) - 1.0) < 1e-12
    # A third, independent evaluation: explicit loops over every 7x7x7 window, straight from the definition that
    # skimage.metrics.structural_similarity documents (Wang et al. 2004, uniform window, SAMPLE covariance, K1 = 0.01, K2 = 0.03,
    # data_range = 2 for float images in skimage 0.19, mean over the windows that fit) -- no filter, no cropping logic to get wrong.
    # skimage itself is not installable in this image (absent from /opt/wheelhouse), so this is as far as the metric can be pinned:
    # "parity unpinned" against the library, pinned against its published algorithm by two implementations that share no code.
    w1, w2 = v1[:11, :9, :10].astype(np.float64), v2[:11, :9, :10].astype(np.float64)
    C1, C2 = (0.01 * 2.0) ** 2, (0.03 * 2.0) ** 2
    vals = []
    for i in range(w1.shape[0] - 6):
        for j in range(w1.shape[1] - 6):
            for k in range(w1.shape[2] - 6):
                p, q = w1[i:i + 7, j:j + 7, k:k + 7].ravel(), w2[i:i + 7, j:j + 7, k:k + 7].ravel()
                mp, mq = p.mean(), q.mean()
                vp, vq, cpq = p.var(ddof=1), q.var(ddof=1), ((p - mp) * (q - mq)).sum() / (p.size - 1)
                vals.append(((2 * mp * mq + C1) * (2 * cpq + C2)) / ((mp ** 2 + mq ** 2 + C1) * (vp + vq + C2)))
    brute = float(np.mean(vals))
    assert abs(naf.ssim_3d(w1, w2) - brute) < 1e-12
    assert abs(get_ssim_3d(torch.from_numpy(w1), torch.from_numpy(w2)) - brute) < 1e-12
    # known answers: a constant offset d between otherwise equal volumes leaves the structure term at 1 -> S = (2 m (m+d) + C1) / (m^2 + (m+d)^2 + C1)
    c = np.full((8, 8, 8), 0.4)
    assert abs(naf.ssim_3d(c, c + 0.2) - (2 * 0.4 * 0.6 + C1) / (0.4 ** 2 + 0.6 ** 2 + C1)) < 1e-12
    x = torch.from_numpy(rng.uniform(0, 1, (5, 6)).astype(np.float32))
    y = torch.from_numpy(rng.uniform(0, 1, (5, 6)).astype(np.float32))
    assert abs(float(get_mse(x, y)) - float(((x - y) ** 2).mean())) < 1e-12 and float(get_psnr(x, y)) > 0
    z = torch.complex(x, y)
    assert abs(float(get_mse(z, z * 0)) - float((x ** 2 + y ** 2).mean())) < 1e-6


def test_phantom_pickle_schema_roundtrip_and_reference_loader(tmp_path):
    """A phantom dataset written in the reference's pickle schema (format_data.py:25-58) loads back, its projections are the
    phantom's line integrals, and -- where the reference is mounted -- the reference's own TIGREDataset reads it and produces
    the rays our geometry code produces, bit for bit."""
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import geometry as G
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import phantom as PH
    # (the global near / far pair of tigre.py:575-586 ignores the tilt: the xy extent must be wide enough for the tilted rays
    #  to reach the object between near and far, as it is in the reference's laminography geometry)
    geom = dict(DSD=1500.0, DSO=1000.0, nDetector=[24, 16], dDetector=[2.0, 2.0], nVoxel=[128, 128, 16], dVoxel=[4.0, 4.0, 2.0],
                offOrigin=[0, 0, 0], offDetector=[0, 0], accuracy=0.5, mode="parallel", filter=None, tilt_angle=10)
    data = PH.make_dataset_dict(geom, n_train=5, n_val=2)
    path = str(tmp_path / "phantom.pickle")
    PH.save_pickle(data, path)
    back, geo = PH.load_pickle(path)
    assert back["train"]["projections"].shape == (5, 16, 24) and back["full_proj"].dtype == np.complex64
    assert back["image"].shape == (128, 128, 16) and geo.mode == "parallel" and geo.tilt_angle == 10
    assert np.allclose(np.angle(back["full_proj"]), back["train"]["projections"], atol=1e-6)
    assert float(back["train"]["projections"].max()) > 0
    ours = G.rays_with_near_far(back["train"]["angles"], geo, "cpu")
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference not mounted: schema checked against our own loader only")
    import sys
    import types
    for name in ["matplotlib", "matplotlib.pyplot", "open3d", "skimage", "skimage.metrics", "imageio", "imageio.v2"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_tigre", "/root/reference/src/dataset/tigre.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ds = mod.TIGREDataset(path, n_rays=32, type="train", device="cpu")
    assert np.array_equal(ds.rays.numpy().view(np.uint32), ours.numpy().view(np.uint32))
    assert np.array_equal(ds.projs.numpy(), back["train"]["projections"])


def test_host_staging_and_pending_loss_logic():
    """Host-side pieces of NAFEngine.train_step_host that need no GPU: the staging copy (memmove fast path and the converting
    fallback) and the PendingLoss handle (waits once, caches the value, survives the reuse of its slot)."""
    import numpy as np
    import torch
    from neuralvolumetricreconstructionformedicalimages_b200.engine import PendingLoss, _stage

    buf = torch.zeros(64, dtype=torch.uint8)
    pix = buf[:48].view(torch.int32).view(4, 3)
    src = torch.arange(12, dtype=torch.int32).view(4, 3)
    _stage(pix, src, torch.int32)                         # same layout: memmove
    assert torch.equal(pix, src)
    _stage(pix, (src + 100).to(torch.int64), torch.int32)  # other dtype: converting copy
    assert torch.equal(pix, src + 100)
    _stage(pix, (src + 7).t().contiguous().t(), torch.int32)   # non-contiguous view: copy_ keeps the logical order
    assert torch.equal(pix, src + 7)
    m = buf[48:52]
    _stage(m, torch.tensor([True, False, True, True]), None)   # bool mask into the uint8 staging bytes
    assert m.tolist() == [1, 0, 1, 1]
    _stage(m, torch.tensor([0, 1, 0, 1], dtype=torch.uint8), None)
    assert m.tolist() == [0, 1, 0, 1]

    class FakeEvent:
        def __init__(self):
            self.synced, self.ready = 0, False

        def query(self):
            return self.ready

        def synchronize(self):
            self.synced += 1
            self.ready = True

    ev, loss = FakeEvent(), np.array([0.25, 3.0], dtype=np.float32)
    p = PendingLoss(ev, loss)
    assert not p.done()
    assert p.result() == 0.25 and ev.synced == 1 and p.done()
    loss[0] = 9.0                                          # the slot is reused by a later step
    assert p.result() == 0.25 and ev.synced == 1           # cached: no second wait, no stale read
    # with a completion word (what the fused forward + loss launch stores after the loss): result() polls it, not the event
    ev2, buf2 = FakeEvent(), np.array([0.5, 7.0, 0.0, 0.0], dtype=np.float32)
    flag = buf2[2:3].view(np.uint32)
    p2 = PendingLoss(ev2, buf2, flag, expect=41)
    flag[0] = 40                                           # an older step's word: not done
    assert not p2.done()
    flag[0] = 41
    assert p2.done() and p2.result() == 0.5 and ev2.synced == 0


def test_level_offsets_are_read_back_once_per_tensor_version():
    """encoder/hashgrid.py::_offsets_host: the op receives the level offsets as a tensor (reference hashgrid.py:28); the host copy the
    C ABI needs is cached per (storage, version), and an in-place change of the tensor invalidates it."""
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import hashgrid as HG
    offs = torch.tensor([0, 10, 30, 70], dtype=torch.int32)
    a = HG._offsets_host(offs)
    assert a.dtype == np.int32 and a.tolist() == [0, 10, 30, 70] and not a.flags.writeable
    assert HG._offsets_host(offs) is a
    offs[3] = 90
    b = HG._offsets_host(offs)
    assert b is not a and b.tolist() == [0, 10, 30, 90]


def test_ctypes_structs_match_the_header(tmp_path):
    """ABI drift guard: every struct of include/nafb200.h has the size -- and its last field the offset -- that the ctypes mirror in
    _lib.py assumes (a C program compiled against the header prints them)."""
    import subprocess
    from neuralvolumetricreconstructionformedicalimages_b200 import _lib
    pairs = [("nafb_grid", _lib.Grid), ("nafb_mlp", _lib.Mlp), ("nafb_mlp_grads", _lib.MlpGrads), ("nafb_sampler", _lib.Sampler),
             ("nafb_loss_tail", _lib.LossTail), ("nafb_pixel_source", _lib.PixelSource), ("nafb_exchange", _lib.Exchange)]
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "nafb200.h"', "int main(void) {"]
    for cname, ct in pairs:
        last = ct._fields_[-1][0]
        src.append(f'  printf("{cname} %zu %zu\\n", sizeof({cname}), offsetof({cname}, {last}));')
    src += ['  printf("abi %d\\n", NAFB_ABI_VERSION);', "  return 0;", "}"]
    c_file, exe = tmp_path / "abi_sizes.c", tmp_path / "abi_sizes"
    c_file.write_text("\n".join(src))
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(c_file), "-o", str(exe)], check=True)
    out = dict((l.split()[0], [int(v) for v in l.split()[1:]]) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert out["abi"] == [_lib.ABI_VERSION]
    for cname, ct in pairs:
        last = ct._fields_[-1][0]
        assert out[cname] == [ctypes.sizeof(ct), getattr(ct, last).offset], (cname, out[cname], ctypes.sizeof(ct), getattr(ct, last).offset)
