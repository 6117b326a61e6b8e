"""Autograd front-ends of the fused kernels (encode + MLP [+ sampling + ray integral]).

Nothing but the sample positions (or the rays) is saved for backward: the backward kernel
recomputes the gather and the activations tile by tile (the table is L2 resident), so no
[P, 32] activation ever goes to HBM.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch.autograd import Function

from . import _lib


class NetMeta:
    """Static description of a DensityNetwork for the kernels (shapes, skips, head, grid)."""

    def __init__(self, offsets_np, D, C, H, in_dim, hidden, out_dim, skips, head, bound, n_layers):
        self.offsets_np = np.ascontiguousarray(offsets_np, dtype=np.int32)
        self.D, self.C, self.H = int(D), int(C), int(H)
        self.in_dim, self.hidden, self.out_dim = int(in_dim), int(hidden), int(out_dim)
        self.skips = [int(s) for s in skips]
        self.head = head
        self.bound = float(bound)
        self.n_layers = int(n_layers)

    def fused_supported(self) -> bool:
        return (self.D == 3 and self.in_dim == 32 and self.hidden == 32 and self.out_dim == 1 and 2 <= self.n_layers <= _lib.NAFB_MAX_LAYERS
                and all(1 <= s <= self.n_layers - 2 for s in self.skips) and self.C in (1, 2, 4, 8))

    def grid(self, table):
        return _lib.make_grid(table, self.offsets_np, self.D, self.C, self.H)

    def mlp(self, params):
        ws, bs = params[0::2], params[1::2]
        return _lib.make_mlp(ws, bs, self.in_dim, self.hidden, self.out_dim, self.skips, self.head)

    def sampler(self, **kw):
        s = _lib.Sampler()
        s.bound = self.bound
        s.clamp = float(np.float32(self.bound - 1e-6))  # render.py:104: python double, cast by clamp()
        for k, v in kw.items():
            setattr(s, k, v)
        return s


def _prep(t, name):
    return _lib.require_cuda(t.contiguous(), name)


_ws_cache = {}


def _workspace(mlp_struct, device):
    n = int(_lib.lib().nafb_density_backward_workspace_bytes(ctypes.byref(mlp_struct)))
    key = (device, n)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.empty(n, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def density_forward(meta: NetMeta, table, params, *, pts=None, rays=None, t_rand=None, n_samples=0, perturb=False,
                    voxels=None, want_acc=False, want_pts=False, want_z=False, want_sigma=True, flags=None):
    """Low-level launcher shared by the autograd functions and the engine. Returns dict of outputs."""
    L_ = _lib.lib()
    dev = table.device
    grid = meta.grid(table)
    mlp = meta.mlp(params)
    out = {}
    if pts is not None:
        P = pts.shape[0]
        smp = meta.sampler(pts=pts.data_ptr(), n_points=P)
        src = _lib.SRC_POINTS
    elif rays is not None:
        N = rays.shape[0]
        P = N * n_samples
        smp = meta.sampler(rays=rays.data_ptr(), t_rand=t_rand.data_ptr() if (perturb and t_rand is not None) else None,
                           n_rays=N, n_samples=n_samples, perturb=int(bool(perturb)))
        src = _lib.SRC_RAYS
    else:
        n1, n2, n3, i0, i1, s1, s2, s3 = voxels
        P = (i1 - i0) * n2 * n3
        smp = meta.sampler(n1=n1, n2=n2, n3=n3, i0=i0, i1=i1, s1=s1, s2=s2, s3=s3)
        src = _lib.SRC_VOXELS
    sigma = torch.empty(P, device=dev, dtype=torch.float32) if want_sigma else None
    acc = torch.zeros(rays.shape[0], device=dev, dtype=torch.float32) if want_acc else None
    z = torch.empty(rays.shape[0], n_samples, device=dev, dtype=torch.float32) if want_z else None
    po = torch.empty(rays.shape[0], n_samples, 3, device=dev, dtype=torch.float32) if want_pts else None
    with torch.cuda.device(dev):
        _lib.check(L_.nafb_density_forward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), src, _lib.ptr(sigma), _lib.ptr(acc),
                                           _lib.ptr(z), _lib.ptr(po), _lib.ptr(flags), _lib.stream_ptr()))
    out.update(sigma=sigma, acc=acc, z_vals=z, pts=po)
    return out


def density_backward(meta: NetMeta, table, params, dsig_or_dacc, grad_table, grad_params, *, pts=None, rays=None, t_rand=None,
                     n_samples=0, perturb=False):
    """Accumulates into grad_table / grad_params (list aligned with params; entries may be None)."""
    L_ = _lib.lib()
    dev = table.device
    grid = meta.grid(table)
    mlp = meta.mlp(params)
    grads = _lib.make_mlp_grads(grad_params[0::2], grad_params[1::2])
    if pts is not None:
        smp = meta.sampler(pts=pts.data_ptr(), n_points=pts.shape[0])
        src = _lib.SRC_POINTS
    else:
        smp = meta.sampler(rays=rays.data_ptr(), t_rand=t_rand.data_ptr() if (perturb and t_rand is not None) else None,
                           n_rays=rays.shape[0], n_samples=n_samples, perturb=int(bool(perturb)))
        src = _lib.SRC_RAYS
    ws = _workspace(mlp, dev)
    with torch.cuda.device(dev):
        _lib.check(L_.nafb_density_backward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), src, _lib.ptr(dsig_or_dacc),
                                            _lib.ptr(grad_table), ctypes.byref(grads), _lib.ptr(ws), _lib.stream_ptr()))


class DensityFn(Function):
    """sigma = DensityNetwork(points)   (reference network.py:34-58), points [P,3]."""

    @staticmethod
    def forward(ctx, pts, table, meta, flags, *params):
        pts = _prep(pts.detach(), "inputs")
        table_c = _prep(table.detach(), "embeddings")
        ps = [_prep(p.detach(), "weight") for p in params]
        out = density_forward(meta, table_c, ps, pts=pts, flags=flags)
        ctx.save_for_backward(pts, table_c, *ps)
        ctx.meta = meta
        return out["sigma"].view(-1, 1)

    @staticmethod
    def backward(ctx, dsigma):
        pts, table, *ps = ctx.saved_tensors
        meta = ctx.meta
        dsigma = _prep(dsigma.reshape(-1), "grad")
        need_table = ctx.needs_input_grad[1]
        grad_table = torch.zeros_like(table) if need_table else None
        grad_params = [torch.zeros_like(p) for p in ps]
        density_backward(meta, table, ps, dsigma, grad_table, grad_params, pts=pts)
        return (None, grad_table, None, None, *grad_params)


class RenderFn(Function):
    """acc, pts, z_vals = fused render of one ray chunk (render.py:82-131 with n_fine == 0)."""

    @staticmethod
    def forward(ctx, rays, t_rand, table, meta, n_samples, perturb, *params):
        rays = _prep(rays.detach(), "rays")
        table_c = _prep(table.detach(), "embeddings")
        ps = [_prep(p.detach(), "weight") for p in params]
        tr = _prep(t_rand.detach(), "t_rand") if (perturb and t_rand is not None) else None
        out = density_forward(meta, table_c, ps, rays=rays, t_rand=tr, n_samples=n_samples, perturb=perturb, want_acc=True,
                              want_pts=True, want_z=True, want_sigma=False)
        ctx.save_for_backward(rays, table_c, *( [tr] if tr is not None else [] ), *ps)
        ctx.has_tr = tr is not None
        ctx.meta, ctx.n_samples, ctx.perturb = meta, n_samples, perturb
        ctx.mark_non_differentiable(out["pts"], out["z_vals"])
        return out["acc"], out["pts"], out["z_vals"]

    @staticmethod
    def backward(ctx, dacc, _dpts, _dz):
        saved = list(ctx.saved_tensors)
        rays, table = saved[0], saved[1]
        tr = saved[2] if ctx.has_tr else None
        ps = saved[3:] if ctx.has_tr else saved[2:]
        dacc = _prep(dacc, "grad")
        grad_table = torch.zeros_like(table) if ctx.needs_input_grad[2] else None
        grad_params = [torch.zeros_like(p) for p in ps]
        density_backward(ctx.meta, table, ps, dacc, grad_table, grad_params, rays=rays, t_rand=tr, n_samples=ctx.n_samples,
                         perturb=ctx.perturb)
        return (None, None, grad_table, None, None, None, *grad_params)
