"""TIGREDataset -- drop-in for the reference's src/dataset/tigre.py:222-382 (same constructor, attributes and item dict), so that
the reference's src/trainer.py (``from .dataset import TIGREDataset``) drives this package unchanged.

What differs from the reference, on purpose:
  * cone AND parallel (tilted) geometry through the package's ray generator (the reference's dataset only reaches its
    parallel-beam get_rays2 and raises NotImplementedError for cone data, tigre.py:247,512-513);
  * `full_proj` is optional (stock NAF pickles have none; the reference requires it, tigre.py:243-245);
  * projections live on the device and one item is drawn there (csrc/select.cu: non-zero pixels only, without replacement,
    np.random.choice(replace=False) semantics of tigre.py:356-358) -- the reference indexes a CPU tensor with CUDA indices
    (tigre.py:364), which raises on a GPU;
  * an item also carries what the fused engine consumes directly: "pixels" [n_rays, 3] int32 (projection, row, col) and, for
    laminography data, "mask" [n_rays] uint8 = get_ptycho_mask(full_proj[index], 0.007) at the drawn pixels (train.py:59-60,93-95),
    computed once per scan instead of once per iteration.
"""
from __future__ import annotations

import pickle

import torch
from torch.utils.data import Dataset

from . import geometry as G
from .mask import PixelSampler


class TIGREDataset(Dataset):
    def __init__(self, path, n_rays=1024, type="train", device="cuda", mask_threshold=0.007):
        super().__init__()
        with open(path, "rb") as handle:
            data = pickle.load(handle)
        self.geo = G.ConeGeometry(data)
        self.type = type
        self.n_rays = n_rays
        self.near, self.far = G.get_near_far(self.geo)
        split = data["train" if type == "train" else "val"]
        self.angles = split["angles"]
        self.projs = torch.tensor(split["projections"], dtype=torch.float32, device=device)
        self.rays = G.rays_with_near_far(self.angles, self.geo, device)                      # [P, H, W, 8] (tigre.py:247-255)
        self.n_samples = data["numTrain" if type == "train" else "numVal"]
        self.image = torch.tensor(data["image"], dtype=torch.float32, device=device)
        self.voxels = torch.tensor(G.get_voxels(self.geo), dtype=torch.float32, device=device if type == "val" else "cpu")
        self.full_proj = None
        if type == "train":
            if data.get("full_proj") is not None:
                self.full_proj = torch.tensor(data["full_proj"], dtype=torch.complex64, device=device)
            H, W = int(self.geo.nDetector[1]), int(self.geo.nDetector[0])
            rows, cols = torch.meshgrid(torch.arange(H, device=device), torch.arange(W, device=device), indexing="ij")
            self.coords = torch.stack([rows, cols], -1).reshape(-1, 2).to(torch.float32)     # tigre.py:257-275
            # the draw counter is not used by __getitem__ (it names its projection); it serves NAFEngine.train_step_sampled
            self.sampler = PixelSampler(self.projs, self.full_proj, mask_threshold)

    def __len__(self):
        return self.n_samples

    def __getitem__(self, index):
        if self.type == "train":
            pixels, projs, mask = self.sampler.draw(int(index), self.n_rays)
            pl = pixels.long()
            out = {"projs": projs, "rays": self.rays[index, pl[:, 1], pl[:, 2]], "coords": pl[:, 1:], "pixels": pixels}
            if self.full_proj is not None:
                out["full_proj"] = self.full_proj[index]
                out["mask"] = mask
            return out
        return {"projs": self.projs[index], "rays": self.rays[index]}
