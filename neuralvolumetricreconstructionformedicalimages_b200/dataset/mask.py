"""Detector-pixel selection for one training step (the per-item work of the reference's TIGREDataset.__getitem__,
src/dataset/tigre.py:354-382, and the ptychography mask of src/utils/util.py:196-205), on the GPU.

The reference filters the non-zero pixels of a projection, draws ``n_rays`` of them without replacement with
``np.random.choice`` on the host, gathers rays / projections / coordinates with mixed-device indexing and recomputes the
mask of the whole complex projection every iteration.  Here the projections, the masks and the per-projection lists of
valid pixels live on the device; one kernel (csrc/select.cu, nafb_draw_pixels) draws a step's batch -- a [N,3] int32 tensor
of (projection, row, col) that the fused kernels turn into rays themselves (nafb_sampler.pixels), the projection values
and the mask bits -- straight into the buffers the training step reads.  The draw counter lives on the device, so the draw
is part of the step's CUDA graph (NAFEngine.train_step_sampled).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib


def get_ptycho_mask(hr: torch.Tensor, threshold: float) -> torch.Tensor:
    """util.py:196-205: pixels whose magnitude is below the threshold AND whose upper and left neighbours agree are
    masked out; returns True where the pixel is KEPT.  hr: complex (or real) [H,W] or [P,H,W] on a CUDA device
    (nafb_ptycho_mask; there is no CPU path)."""
    if not hr.is_cuda:
        raise RuntimeError("get_ptycho_mask: hr must be a CUDA tensor (this package has no CPU path)")
    shape = hr.shape
    if hr.dim() not in (2, 3):
        raise ValueError("get_ptycho_mask: hr must be [H,W] or [P,H,W]")
    c = hr.to(torch.complex64).reshape(-1, shape[-2], shape[-1]).contiguous()
    keep = torch.empty(c.shape, dtype=torch.uint8, device=hr.device)
    with torch.cuda.device(hr.device):
        _lib.check(_lib.lib().nafb_ptycho_mask(ctypes.c_void_p(torch.view_as_real(c).data_ptr()), c.shape[0], c.shape[1], c.shape[2],
                                               float(threshold), _lib.ptr(keep), _lib.stream_ptr()))
    return keep.reshape(shape).bool()


class PixelSampler:
    """Device-resident projections + masks + valid-pixel lists; draws the pixel batch of one step on the device."""

    def __init__(self, projs: torch.Tensor, full_proj: torch.Tensor | None = None, threshold: float = 0.007, seed: int | None = None,
                 order=None):
        """projs [P,H,W] real (what the loss compares against); full_proj [P,H,W] complex (lamino data) or None.
        order: the projection of draw k is order[k % len(order)] (default: k % P, the reference's un-shuffled DataLoader)."""
        if not projs.is_cuda:
            raise RuntimeError("PixelSampler: projections must live on a CUDA device (this package has no CPU path)")
        self.projs = projs.contiguous().to(torch.float32)
        P, H, W = projs.shape
        self.shape = (P, H, W)
        dev = projs.device
        self.device = dev
        self.mask = get_ptycho_mask(full_proj.to(dev), threshold).to(torch.uint8).contiguous() if full_proj is not None else None
        # tigre.py:356: only pixels with a non-zero projection value are candidates.  Front-packed index lists [P, H*W]:
        # a stable sort of the "is zero" flag keeps the valid pixels in raster order
        flat = self.projs.reshape(P, H * W)
        nz = flat != 0
        self.n_valid = nz.sum(dim=1).to(torch.int32).contiguous()
        self.valid_packed = torch.sort((~nz).to(torch.uint8), dim=1, stable=True).indices.to(torch.int32).contiguous()
        self.valid = [self.valid_packed[p, : int(self.n_valid[p])] for p in range(P)]
        seed = int(torch.initial_seed() if seed is None else seed) & 0xFFFFFFFFFFFFFFFF
        self.state = torch.from_numpy(np.array([0, seed & 0xFFFFFFFF, seed >> 32, 0], dtype=np.uint32).view(np.int32)).to(dev)
        self.order = None if order is None else torch.as_tensor(order, dtype=torch.int32, device=dev).contiguous()
        self.version = 0      # bumped by every python-level call that moves the draw counter (NAFEngine's prefetch checks it)
        self._src = _lib.PixelSource(projs=self.projs.data_ptr(), mask=self.mask.data_ptr() if self.mask is not None else None,
                                     valid=self.valid_packed.data_ptr(), n_valid=self.n_valid.data_ptr(),
                                     order=self.order.data_ptr() if self.order is not None else None, n_proj=P, H=H, W=W,
                                     n_order=0 if self.order is None else int(self.order.numel()))

    # ------------------------------------------------------------------ the kernel path
    def draw_into(self, n_rays: int, pixels: torch.Tensor, projs: torch.Tensor, mask: torch.Tensor | None):
        """Enqueue ONE draw (the next projection of the order) into caller-owned device buffers: pixels [N,3] int32, projs [N]
        fp32, mask [N] uint8 (or None).  Graph-capturable: the draw counter advances on the device."""
        self.version += 1
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nafb_draw_pixels(ctypes.byref(self._src), int(n_rays), _lib.ptr(pixels), _lib.ptr(projs), _lib.ptr(mask),
                                                   _lib.ptr(self.state), _lib.stream_ptr()))

    def set_draw(self, k: int):
        """Position of the draw counter (draw k takes projection order[k % len(order)])."""
        self.version += 1
        self.state[0:1] = torch.tensor([int(k)], dtype=torch.int32)

    def draws_done(self) -> int:
        """Draws made so far (synchronises).  NAFEngine.train_step_sampled keeps ONE draw ahead of the steps it has run."""
        return int(self.state[0].item())

    def check(self):
        """Raise what the reference's np.random.choice raises when a projection has fewer valid pixels than rays (synchronises)."""
        e = int(self.state[3].item())
        if e:
            raise ValueError(f"projection {e - 1} has fewer valid pixels than n_rays (cannot draw without replacement)")

    def draw(self, proj: int, n_rays: int, generator=None):
        """(pixels [N,3] int32, projs [N] fp32, mask [N] uint8) for projection `proj`: n_rays valid pixels without replacement,
        uniformly and in random order (np.random.choice(replace=False) at tigre.py:358).  `generator`: a torch.Generator whose
        next value seeds this one draw (reproducible streams for tests); default: the sampler's own counter-based stream."""
        P, H, W = self.shape
        if int(self.n_valid[proj]) < n_rays:
            raise ValueError(f"projection {proj} has only {int(self.n_valid[proj])} valid pixels (< n_rays = {n_rays})")
        dev = self.device
        pixels = torch.empty(n_rays, 3, dtype=torch.int32, device=dev)
        projs = torch.empty(n_rays, dtype=torch.float32, device=dev)
        mask = torch.empty(n_rays, dtype=torch.uint8, device=dev)
        if generator is not None:
            s = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device=generator.device).item())
            st = np.array([0, s & 0xFFFFFFFF, (s >> 32) & 0xFFFFFFFF, 0], dtype=np.uint32).view(np.int32)
        else:
            st = self.state.cpu().numpy().copy()
            self.state[0:1] += 1
            self.version += 1
            st[0] = (int(st[0]) * 2654435761) & 0x7FFFFFFF     # decorrelate from draw_into's counter, keep the explicit projection below
        one = torch.tensor([proj], dtype=torch.int32, device=dev)
        state = torch.from_numpy(st).to(dev)
        src = _lib.PixelSource(projs=self.projs.data_ptr(), mask=self.mask.data_ptr() if self.mask is not None else None,
                               valid=self.valid_packed.data_ptr(), n_valid=self.n_valid.data_ptr(), order=one.data_ptr(), n_proj=P, H=H, W=W,
                               n_order=1)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nafb_draw_pixels(ctypes.byref(src), int(n_rays), _lib.ptr(pixels), _lib.ptr(projs), _lib.ptr(mask),
                                                   _lib.ptr(state), _lib.stream_ptr()))
        return pixels, projs, mask

    def draw_epoch(self, n_rays: int, generator=None, projections=None):
        """The batches of a whole epoch: for every projection in `projections` (default: all, in order) n_rays valid pixels without
        replacement.  Returns (pixels [K,N,3] int32, projs [K,N], mask [K,N] uint8).  K launches of the draw kernel, no host
        synchronisation."""
        P, H, W = self.shape
        sel = list(range(P)) if projections is None else [int(p) for p in projections]
        if int(self.n_valid[torch.as_tensor(sel, device=self.device)].min()) < n_rays:
            raise ValueError(f"a projection has fewer than n_rays = {n_rays} valid pixels")
        out = [self.draw(p, n_rays, generator) for p in sel]
        return torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out]), torch.stack([o[2] for o in out])
