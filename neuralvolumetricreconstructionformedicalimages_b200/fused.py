"""Autograd front-ends of the fused kernels (encode + MLP [+ sampling + ray integral]).

What is saved for backward: the sample positions (or the rays) and -- in the tensor-core mode --
the "stash" the forward kernel leaves behind: the bf16 (hi|lo) images of the encodings, 128 B per
point, already in the operand layout of the backward kernel's MMAs.  The backward kernel reloads
them with full-line loads (no second table gather: the gather is bound by the SM's one
address-divergent wavefront per clock, not by bytes) and recomputes the MLP activations tile by
tile.  Without a stash (fp32 SIMT mode, inference) everything is recomputed from the points.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch.autograd import Function

from . import _lib


# Arithmetic of the fused density kernels for networks that do not choose one themselves (NetMeta.arith is None):
# _lib.ARITH_TC = tcgen05 tensor cores (bf16x3 split products, fp32 accumulation) where the configuration allows,
# _lib.ARITH_SIMT = fp32 FMAs everywhere.  The choice travels WITH EVERY CALL (nafb_mlp.arith): the library keeps no mode.
DEFAULT_ARITH = _lib.ARITH_TC


def set_default_arithmetic(arith: int) -> int:
    """Set the package-wide default (returns the previous one).  Per network: `net.fused_meta().arith = ...`."""
    global DEFAULT_ARITH
    if arith not in (_lib.ARITH_TC, _lib.ARITH_SIMT):
        raise ValueError("arithmetic must be _lib.ARITH_TC (0) or _lib.ARITH_SIMT (1)")
    prev, DEFAULT_ARITH = DEFAULT_ARITH, int(arith)
    return prev


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
        self.arith = None   # None: follow fused.DEFAULT_ARITH at call time

    def fused_supported(self) -> bool:
        return (self.D == 3 and self.in_dim == 32 and self.hidden == 32 and self.out_dim == 1 and 2 <= self.n_layers <= _lib.NAFB_MAX_LAYERS
                and all(1 <= s <= self.n_layers - 2 for s in self.skips) and self.C in (1, 2, 4, 8))

    def grid(self, table):
        return _lib.make_grid(table, self.offsets_np, self.D, self.C, self.H)

    def mlp(self, params):
        ws, bs = params[0::2], params[1::2]
        return _lib.make_mlp(ws, bs, self.in_dim, self.hidden, self.out_dim, self.skips, self.head,
                             DEFAULT_ARITH if self.arith is None else self.arith)

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
    """Workspace of the autograd front-ends: one per (device, CUDA stream) -- launches on one stream are ordered, launches on
    different streams must not share the grid-barrier words."""
    n = int(_lib.lib().nafb_density_backward_workspace_bytes(ctypes.byref(mlp_struct)))
    key = (device, n, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.zeros(n, dtype=torch.uint8, device=device)   # zero-filled: it holds the words of the backward kernel's grid barrier
        _ws_cache[key] = ws
    return ws


def new_workspace(mlp_struct, device):
    """A private backward workspace (zero-filled): for callers that may have several backward launches in flight."""
    n = int(_lib.lib().nafb_density_backward_workspace_bytes(ctypes.byref(mlp_struct)))
    return torch.zeros(n, dtype=torch.uint8, device=device)


def stash_bytes(meta: NetMeta, table, params, n_points: int) -> int:
    """Size of the encoding stash for n_points in the current arithmetic mode (0: not used)."""
    grid = meta.grid(table)
    mlp = meta.mlp(params)
    return int(_lib.lib().nafb_density_stash_bytes(ctypes.byref(grid), ctypes.byref(mlp), int(n_points)))


def density_forward(meta: NetMeta, table, params, *, pts=None, rays=None, t_rand=None, n_samples=0, perturb=False,
                    voxels=None, want_acc=False, want_pts=False, want_z=False, want_sigma=True, flags=None, want_stash=False,
                    sigma_out=None):
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
    if sigma_out is not None:
        if sigma_out.numel() != P or sigma_out.dtype != torch.float32 or not sigma_out.is_contiguous() or sigma_out.device != dev:
            raise RuntimeError("sigma_out must be a contiguous float32 tensor of P elements on the table's device")
        sigma = sigma_out.view(-1)
    else:
        sigma = torch.empty(P, device=dev, dtype=torch.float32) if want_sigma else None
    acc = torch.zeros(rays.shape[0], device=dev, dtype=torch.float32) if want_acc else None
    z = torch.empty(rays.shape[0], n_samples, device=dev, dtype=torch.float32) if want_z else None
    po = torch.empty(rays.shape[0], n_samples, 3, device=dev, dtype=torch.float32) if want_pts else None
    stash = None
    if want_stash:
        nb = int(L_.nafb_density_stash_bytes(ctypes.byref(grid), ctypes.byref(mlp), P))
        stash = torch.empty(nb, dtype=torch.uint8, device=dev) if nb else None
    with torch.cuda.device(dev):
        _lib.check(L_.nafb_density_forward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), src, _lib.ptr(sigma), _lib.ptr(acc),
                                           _lib.ptr(z), _lib.ptr(po), _lib.ptr(flags), _lib.ptr(stash), _lib.stream_ptr()))
    out.update(sigma=sigma, acc=acc, z_vals=z, pts=po, stash=stash)
    return out


def density_backward(meta: NetMeta, table, params, dsig_or_dacc, grad_table, grad_params, *, pts=None, rays=None, t_rand=None,
                     n_samples=0, perturb=False, stash=None, rng_state=None, sampler=None, n_points=None, workspace=None):
    """Accumulates into grad_table / grad_params (list aligned with params; entries may be None).
    `stash` is what density_forward(want_stash=True) returned for the same points, or None.
    `sampler`: a ready nafb_sampler of the RAYS source (the engine passes the one its forward used) with `n_points`."""
    L_ = _lib.lib()
    dev = table.device
    grid = meta.grid(table)
    mlp = meta.mlp(params)
    grads = _lib.make_mlp_grads(grad_params[0::2], grad_params[1::2])
    if sampler is not None:
        smp, src = sampler, _lib.SRC_RAYS
    elif pts is not None:
        smp = meta.sampler(pts=pts.data_ptr(), n_points=pts.shape[0])
        src = _lib.SRC_POINTS
    else:
        smp = meta.sampler(rays=rays.data_ptr(), t_rand=t_rand.data_ptr() if (perturb and t_rand is not None) else None,
                           rng_state=rng_state.data_ptr() if (perturb and t_rand is None and rng_state is not None) else None,
                           n_rays=rays.shape[0], n_samples=n_samples, perturb=int(bool(perturb)))
        src = _lib.SRC_RAYS
    ws = workspace if workspace is not None else _workspace(mlp, dev)
    if stash is not None:
        P = n_points if sampler is not None else (pts.shape[0] if pts is not None else rays.shape[0] * n_samples)
        if stash.numel() != int(L_.nafb_density_stash_bytes(ctypes.byref(grid), ctypes.byref(mlp), P)):
            stash = None  # the arithmetic mode changed between forward and backward: recompute
    with torch.cuda.device(dev):
        _lib.check(L_.nafb_density_backward(ctypes.byref(grid), ctypes.byref(mlp), ctypes.byref(smp), src, _lib.ptr(dsig_or_dacc),
                                            _lib.ptr(grad_table), ctypes.byref(grads), _lib.ptr(ws), _lib.ptr(stash), _lib.stream_ptr()))


class DensityFn(Function):
    """sigma = DensityNetwork(points)   (reference network.py:34-58), points [P,3]."""

    @staticmethod
    def forward(ctx, pts, table, meta, flags, *params):
        pts = _prep(pts.detach(), "inputs")
        table_c = _prep(table.detach(), "embeddings")
        ps = [_prep(p.detach(), "weight") for p in params]
        need_grad = any(ctx.needs_input_grad[i] for i in (1, *range(4, 4 + len(params))))
        out = density_forward(meta, table_c, ps, pts=pts, flags=flags, want_stash=need_grad)
        ctx.save_for_backward(pts, table_c, *ps)
        ctx.meta, ctx.stash = meta, out["stash"]
        return out["sigma"].view(-1, 1)

    @staticmethod
    def backward(ctx, dsigma):
        pts, table, *ps = ctx.saved_tensors
        meta = ctx.meta
        dsigma = _prep(dsigma.reshape(-1), "grad")
        need_table = ctx.needs_input_grad[1]
        grad_table = torch.zeros_like(table) if need_table else None
        grad_params = [torch.zeros_like(p) for p in ps]
        density_backward(meta, table, ps, dsigma, grad_table, grad_params, pts=pts, stash=ctx.stash)
        ctx.stash = None
        return (None, grad_table, None, None, *grad_params)


class RenderFn(Function):
    """acc, pts, z_vals = fused render of one ray chunk (render.py:82-131 with n_fine == 0)."""

    @staticmethod
    def forward(ctx, rays, t_rand, table, meta, n_samples, perturb, *params):
        rays = _prep(rays.detach(), "rays")
        table_c = _prep(table.detach(), "embeddings")
        ps = [_prep(p.detach(), "weight") for p in params]
        tr = _prep(t_rand.detach(), "t_rand") if (perturb and t_rand is not None) else None
        need_grad = any(ctx.needs_input_grad[i] for i in (2, *range(6, 6 + len(params))))
        out = density_forward(meta, table_c, ps, rays=rays, t_rand=tr, n_samples=n_samples, perturb=perturb, want_acc=True,
                              want_pts=True, want_z=True, want_sigma=False, want_stash=need_grad)
        ctx.save_for_backward(rays, table_c, *( [tr] if tr is not None else [] ), *ps)
        ctx.has_tr, ctx.stash = tr is not None, out["stash"]
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
                         perturb=ctx.perturb, stash=ctx.stash)
        ctx.stash = None
        return (None, None, grad_table, None, None, None, *grad_params)
