"""Detector-pixel selection for one training step (the per-item work of the reference's TIGREDataset.__getitem__,
src/dataset/tigre.py:354-382, and the ptychography mask of src/utils/util.py:196-205), kept on the GPU.

The reference filters the non-zero pixels of a projection, draws ``n_rays`` of them without replacement with
``np.random.choice`` on the host, gathers rays / projections / coordinates with mixed-device indexing and recomputes the
mask of the whole complex projection every iteration.  Here the projections, the masks and the per-projection lists of
valid pixels live on the device; one step's batch is a [N,3] int32 tensor of (projection, row, col) that the fused
kernels turn into rays themselves (nafb_sampler.pixels).
"""
from __future__ import annotations

import torch


def get_ptycho_mask(hr: torch.Tensor, threshold: float) -> torch.Tensor:
    """util.py:196-205: pixels whose magnitude is below the threshold AND whose upper and left neighbours agree are
    masked out; returns True where the pixel is KEPT.  (The right-hand sides are evaluated before the in-place AND, as
    in the reference.)"""
    m = torch.abs(hr) < threshold
    m[1:, :] &= (m[1:, :] == m[:-1, :])
    m[:, 1:] &= (m[:, 1:] == m[:, :-1])
    return ~m


class PixelSampler:
    """Device-resident projections + masks + valid-pixel lists; draws the pixel batch of one step on the device."""

    def __init__(self, projs: torch.Tensor, full_proj: torch.Tensor | None = None, threshold: float = 0.007):
        """projs [P,H,W] real (what the loss compares against); full_proj [P,H,W] complex (lamino data) or None."""
        self.projs = projs.contiguous()
        P, H, W = projs.shape
        self.shape = (P, H, W)
        dev = projs.device
        if full_proj is not None:
            self.mask = torch.stack([get_ptycho_mask(full_proj[p].clone(), threshold) for p in range(P)]).to(torch.uint8)
        else:
            self.mask = torch.ones(P, H, W, dtype=torch.uint8, device=dev)
        # tigre.py:356: only pixels with a non-zero projection value are candidates
        self.valid = [torch.nonzero(projs[p].reshape(-1) != 0).reshape(-1).to(torch.int32) for p in range(P)]

    def draw_epoch(self, n_rays: int, generator=None, projections=None):
        """The batches of a whole epoch in a handful of batched device ops: for every projection in `projections` (default:
        all, in order) n_rays valid pixels without replacement.  Returns (pixels [K,N,3] int32, projs [K,N], mask [K,N] uint8).
        Uniform sampling without replacement = the n_rays smallest of i.i.d. uniform keys over the valid pixels (invalid
        pixels get a key that sorts last) -- the same distribution as one np.random.choice(replace=False) per item
        (tigre.py:358), without a host round trip or a launch sequence per iteration."""
        P, H, W = self.shape
        dev = self.projs.device
        sel_p = torch.arange(P, device=dev) if projections is None else torch.as_tensor(projections, device=dev, dtype=torch.long)
        if not hasattr(self, "_valid_mask"):
            self._valid_mask = (self.projs.reshape(P, -1) != 0)
            self._n_valid = self._valid_mask.sum(dim=1)
        if int(self._n_valid[sel_p].min()) < n_rays:
            raise ValueError(f"a projection has fewer than n_rays = {n_rays} valid pixels")
        keys = torch.rand(sel_p.numel(), H * W, device=dev, generator=generator)
        keys.masked_fill_(~self._valid_mask[sel_p], 2.0)
        sel = keys.topk(n_rays, dim=1, largest=False, sorted=False).indices                 # [K, N] flat pixel ids
        row, col = sel // W, sel % W
        pixels = torch.stack([sel_p[:, None].expand_as(sel), row, col], dim=2).to(torch.int32).contiguous()
        flat_p = self.projs.reshape(P, -1)[sel_p]
        flat_m = self.mask.reshape(P, -1)[sel_p]
        return pixels, torch.gather(flat_p, 1, sel), torch.gather(flat_m, 1, sel)

    def draw(self, proj: int, n_rays: int, generator=None):
        """(pixels [N,3] int32, projs [N] fp32, mask [N] uint8) for projection `proj`: n_rays valid pixels without replacement
        (uniform, as np.random.choice(replace=False) at tigre.py:358)."""
        P, H, W = self.shape
        cand = self.valid[proj]
        if cand.numel() < n_rays:
            raise ValueError(f"projection {proj} has only {cand.numel()} valid pixels (< n_rays = {n_rays})")
        sel = cand[torch.randperm(cand.numel(), device=cand.device, generator=generator)[:n_rays]].to(torch.int64)
        row, col = sel // W, sel % W
        pixels = torch.stack([torch.full_like(row, proj), row, col], dim=1).to(torch.int32).contiguous()
        return pixels, self.projs[proj].reshape(-1)[sel], self.mask[proj].reshape(-1)[sel]
