"""The boundary BASELINE.json's north_star names: "keeps their Python call signatures (get_encoder, DensityNetwork.forward,
render(rays, net, ...)) so src/trainer.py and the YAML configs drive it unchanged".

The reference's OWN src/trainer.py and src/config/configloading.py (the staged, unmodified copy under baseline/_ref/src) are
imported as the package ``src`` whose sub-packages encoder / network / render / loss / dataset are THIS repository's modules.
config/chest_50.yaml is loaded with the reference's load_config; only the experiment paths, the number of epochs and the
eval / save periods are overridden (1 500 epochs would be the whole schedule).  A Trainer subclass supplies the two hooks the
reference's train.py supplies (compute_loss: train.py:48-135 with its chunk-slicing bug corrected -- SURVEY.md section 5.7;
eval_step: train.py:220-286 without the image files).  The run must train, evaluate, write ckpt.tar in the reference's format
and RESUME from it (trainer.py:60-70,114-126); the same checkpoint must also move into the fused engine and back.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
PKG = "neuralvolumetricreconstructionformedicalimages_b200"
DEV = "cuda"


def _alias_src():
    """``src`` = the reference's tree, with the hot-path sub-packages replaced by this repository's."""
    if not (os.path.exists(os.path.join(REF, "src", "trainer.py")) and os.path.exists(os.path.join(REF, "config", "chest_50.yaml"))):
        pytest.skip("the reference's trainer / YAMLs are not staged on this machine (baseline/stage_ref.sh)")
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    src = types.ModuleType("src")
    src.__path__ = [os.path.join(REF, "src")]          # trainer.py and config/ are found here, unmodified
    sys.modules["src"] = src
    for name in ("encoder", "network", "render", "loss", "dataset"):
        mod = importlib.import_module(f"{PKG}.{name}")
        sys.modules[f"src.{name}"] = mod
        setattr(src, name, mod)
    from src.config.configloading import load_config
    from src.trainer import Trainer
    assert os.path.samefile(sys.modules["src.trainer"].__file__, os.path.join(REF, "src", "trainer.py"))
    return Trainer, load_config


def _phantom_pickle(path, n_train=6, n_val=2):
    from neuralvolumetricreconstructionformedicalimages_b200.dataset import phantom as PH
    geometry = dict(DSD=1500.0, DSO=1000.0, nDetector=[128, 128], dDetector=[2.0, 2.0], nVoxel=[32, 32, 32], dVoxel=[4.0, 4.0, 4.0],
                    offOrigin=[0, 0, 0], offDetector=[0, 0], accuracy=0.5, mode="cone", filter=None)
    PH.save_pickle(PH.make_dataset_dict(geometry, n_train, n_val), path)


def _make_trainer_class(Trainer):
    from src.loss import calc_mse_loss
    from src.render import render, run_network
    from neuralvolumetricreconstructionformedicalimages_b200.utils import get_mse, get_psnr, get_psnr_3d, get_ssim_3d

    class BasicTrainer(Trainer):
        def compute_loss(self, data, global_step, idx_epoch):
            rays = data["rays"].reshape(-1, 8)
            projs = data["projs"].reshape(-1)
            mask = data["mask"].reshape(-1).bool() if "mask" in data else torch.ones_like(projs, dtype=torch.bool)
            ret = render(rays, self.net, self.net_fine, **self.conf["render"])         # one fused launch for the whole batch
            acc = ret["acc"].reshape(-1)
            loss = {"loss": 0.0}
            for i in range(0, rays.shape[0], 200):                                       # train.py:69-127 (chunk means added up)
                m = mask[i:i + 200]
                calc_mse_loss(loss, projs[i:i + 200][m], acc[i:i + 200][m])
            self.writer.add_scalar("train/loss", loss["loss"].item(), global_step)
            return loss["loss"]

        def eval_step(self, global_step, idx_epoch):
            sel = 0
            projs = self.eval_dset.projs[sel]
            rays = self.eval_dset.rays[sel].reshape(-1, 8)
            H, W = projs.shape
            pred = torch.cat([render(rays[i:i + self.n_rays], self.net, self.net_fine, **self.conf["render"])["acc"]
                              for i in range(0, rays.shape[0], self.n_rays)], 0).reshape(H, W)
            image_pred = run_network(self.eval_dset.voxels, self.net_fine if self.net_fine is not None else self.net, self.netchunk).squeeze()
            image = self.eval_dset.image
            loss = {"proj_mse": get_mse(pred, projs), "proj_psnr": get_psnr(pred, projs),
                    "psnr_3d": torch.tensor(get_psnr_3d(image_pred, image)), "ssim_3d": torch.tensor(get_ssim_3d(image_pred, image))}
            for k, v in loss.items():
                self.writer.add_scalar(f"eval/{k}", v, global_step)
            self.eval_history = getattr(self, "eval_history", []) + [{k: float(v) for k, v in loss.items()}]
            return loss

    return BasicTrainer


def test_reference_trainer_and_yaml_drive_the_package(tmp_path):
    Trainer, load_config = _alias_src()
    BasicTrainer = _make_trainer_class(Trainer)
    pickle_path = os.path.join(str(tmp_path), "phantom.pickle")
    _phantom_pickle(pickle_path)

    def config(epochs, resume):
        cfg = load_config(os.path.join(REF, "config", "chest_50.yaml"))                  # the shipped YAML, the reference's loader
        assert cfg["encoder"]["encoding"] == "hashgrid" and cfg["train"]["n_rays"] == 1024 and cfg["render"]["n_samples"] == 192
        cfg["exp"].update(expdir=os.path.join(str(tmp_path), "logs"), datadir=pickle_path)
        cfg["train"].update(epoch=epochs, resume=resume)
        cfg["log"].update(i_eval=1, i_save=1)
        return cfg

    torch.manual_seed(0)
    np.random.seed(0)
    tr = BasicTrainer(config(2, False), DEV)
    # the objects the reference trainer built are this package's
    assert type(tr.net).__module__.startswith(PKG) and type(tr.net.encoder).__module__.startswith(PKG)
    assert tr.net.encoder.embeddings.shape == (7131219, 2) and tr.net.fused_meta() is not None
    assert type(tr.train_dloader.dataset).__module__.startswith(PKG)
    tr.start()                                                                          # epochs 0..2: eval, 6 iterations, save
    hist = tr.eval_history
    assert len(hist) == 3 and hist[-1]["proj_mse"] < 0.98 * hist[0]["proj_mse"] and hist[-1]["psnr_3d"] > hist[0]["psnr_3d"], hist
    ckpt_path = os.path.join(str(tmp_path), "logs", "chest_50", "ckpt.tar")
    assert os.path.exists(ckpt_path) and os.path.exists(ckpt_path.replace("ckpt.tar", "ckpt_backup.tar"))
    ckpt = torch.load(ckpt_path, weights_only=False)
    assert ckpt["epoch"] == 2 and ckpt["network_fine"] is None
    assert sorted(ckpt["network"]) == sorted(["encoder.embeddings"] + [f"layers.{i}.{k}" for i in range(4) for k in ("weight", "bias")])
    assert ckpt["network"]["layers.2.weight"].shape == (32, 64)
    n_steps = 3 * len(tr.train_dloader)
    assert int(float(ckpt["optimizer"]["state"][0]["step"])) == n_steps
    params_after = {k: v.clone() for k, v in tr.net.state_dict().items()}

    # ---- resume with the reference trainer (trainer.py:60-70)
    tr2 = BasicTrainer(config(4, True), DEV)
    assert tr2.epoch_start == 3 and tr2.global_step == 3 * len(tr2.train_dloader)
    for k, v in tr2.net.state_dict().items():
        assert torch.equal(v, params_after[k]), k
    st = tr2.optimizer.state_dict()["state"]
    assert int(float(st[0]["step"])) == n_steps and torch.equal(st[0]["exp_avg"], ckpt["optimizer"]["state"][0]["exp_avg"].to(DEV))
    tr2.start()                                                                         # epochs 3, 4
    assert tr2.eval_history[-1]["proj_mse"] < hist[-1]["proj_mse"]
    ckpt = torch.load(ckpt_path, weights_only=False)                                    # now the state after epoch 4
    assert ckpt["epoch"] == 4
    n_steps = 5 * len(tr.train_dloader)

    # ---- the same checkpoint moves into the fused engine and back (NAFEngine.optimizer_state_dict: torch.optim.Adam layout)
    from neuralvolumetricreconstructionformedicalimages_b200.encoder import get_encoder
    from neuralvolumetricreconstructionformedicalimages_b200.engine import NAFEngine
    from neuralvolumetricreconstructionformedicalimages_b200.network import get_network
    cfg = config(2, False)
    cfg["network"].pop("net_type")
    net = get_network("mlp")(get_encoder(**cfg["encoder"]), **cfg["network"]).to(DEV)
    net.load_state_dict(ckpt["network"])
    eng = NAFEngine(net, lr=cfg["train"]["lrate"], n_samples=cfg["render"]["n_samples"], perturb=True, loss_chunk=200)
    eng.load_optimizer_state_dict(ckpt["optimizer"])
    assert eng.step_count == n_steps
    sd = eng.optimizer_state_dict()
    ref_sd = ckpt["optimizer"]
    assert sd["param_groups"][0]["params"] == ref_sd["param_groups"][0]["params"] and sd["param_groups"][0]["lr"] == ref_sd["param_groups"][0]["lr"]
    for pid in ref_sd["state"]:
        for k in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(sd["state"][pid][k], ref_sd["state"][pid][k].to(DEV)), (pid, k)
        assert float(sd["state"][pid]["step"]) == float(ref_sd["state"][pid]["step"])
    # a fused step from the restored state == a reference-trainer step from the same state (same batch, same uniforms)
    item = tr.train_dloader.dataset[1]
    S = cfg["render"]["n_samples"]
    t_rand = torch.rand(cfg["train"]["n_rays"], S, device=DEV)
    l_eng = eng.train_step(item["rays"], item["projs"], None, t_rand)
    tr3 = BasicTrainer(config(2, True), DEV)                                             # restores the same checkpoint
    real = torch.rand
    torch.rand = lambda *a, **k: t_rand.clone()
    try:
        l_ref = tr3.train_step({"rays": item["rays"][None], "projs": item["projs"][None]}, global_step=0, idx_epoch=0)
    finally:
        torch.rand = real
    np.testing.assert_allclose(float(l_eng), l_ref, rtol=2e-4)
    for (k, a), b in zip(net.state_dict().items(), tr3.net.state_dict().values()):
        frac = float(((a - b).abs() > 1e-5).float().mean())
        assert frac < 1e-3, (k, frac)
    # and back: what the engine writes is what torch.optim.Adam loads
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    opt.load_state_dict(eng.optimizer_state_dict())
    assert int(float(opt.state_dict()["state"][0]["step"])) == n_steps + 1
