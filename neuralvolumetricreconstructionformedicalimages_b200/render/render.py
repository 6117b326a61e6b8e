"""render / run_network / raw2outputs / sample_pdf -- drop-in for the reference's
src/render/render.py (same names, positional order, defaults and returned dict keys).

Fast path: when ``net`` is this package's DensityNetwork in a fused-capable configuration and no
fine network is requested, a whole ray chunk (sampling -> gather -> MLP -> sum sigma*delta) is one
kernel forward and one kernel backward.  Anything else (foreign ``net`` callables, hierarchical
sampling) goes through the same stages as separate operators.
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function

from .. import _lib
from ..fused import RenderFn
from ..network.network import DensityNetwork

# The reference prints a warning when an output contains NaN/Inf (render.py:141-144), which costs
# several host synchronisations per chunk.  Opt in with CHECK_NUMERICS = True.
CHECK_NUMERICS = False
# Hierarchical sampling (n_fine > 0): sample_pdf + merge + sort + positions + TV as one kernel (nafb_sample_fine).  False: the
# same stages as torch operators (kept for A/B tests and for shapes the kernel does not take).
FUSED_FINE_SAMPLING = True


def compute_tv_regularization(pts):
    """0-th order TV of the sample positions (render.py:16-28): sum |pts[:,1:] - pts[:,:-1]|."""
    return torch.sum(torch.abs(pts[:, 1:, :] - pts[:, :-1, :]))


def _sampler_struct(rays, t_rand, n_samples, perturb, bound):
    import numpy as np
    s = _lib.Sampler()
    s.rays = rays.data_ptr()
    s.t_rand = t_rand.data_ptr() if (perturb and t_rand is not None) else None
    s.n_rays, s.n_samples, s.perturb = rays.shape[0], int(n_samples), int(bool(perturb))
    s.bound = float(bound)
    s.clamp = float(np.float32(bound - 1e-6))
    return s


def sample_points(rays, n_samples, perturb, t_rand, bound, want_tv=False):
    """render.py:88-105 in one kernel -> z_vals [N,S], pts [N,S,3] (+ per-ray TV partial sums)."""
    L_ = _lib.lib()
    rays = _lib.require_cuda(rays.contiguous(), "rays")
    N = rays.shape[0]
    z = torch.empty(N, n_samples, device=rays.device, dtype=torch.float32)
    pts = torch.empty(N, n_samples, 3, device=rays.device, dtype=torch.float32)
    tv = torch.empty(N, device=rays.device, dtype=torch.float32) if want_tv else None
    if perturb:
        t_rand = _lib.require_cuda(t_rand.contiguous(), "t_rand")
    s = _sampler_struct(rays, t_rand, n_samples, perturb, bound)
    with torch.cuda.device(rays.device):
        _lib.check(L_.nafb_sample_points(ctypes.byref(s), _lib.ptr(z), _lib.ptr(pts), _lib.ptr(tv), _lib.stream_ptr()))
    return z, pts, tv


def _tv_partial(rays, n_samples, perturb, t_rand, bound):
    L_ = _lib.lib()
    tv = torch.empty(rays.shape[0], device=rays.device, dtype=torch.float32)
    s = _sampler_struct(rays, t_rand, n_samples, perturb, bound)
    with torch.cuda.device(rays.device):
        _lib.check(L_.nafb_sample_points(ctypes.byref(s), None, None, _lib.ptr(tv), _lib.stream_ptr()))
    return tv


class _RayIntegral(Function):
    """acc = sum_i raw[..., 0] * dists  (render.py:192-201) with its gradient wrt raw."""

    @staticmethod
    def forward(ctx, raw, z_vals, rays):
        L_ = _lib.lib()
        raw = _lib.require_cuda(raw.contiguous(), "raw")
        z_vals = _lib.require_cuda(z_vals.contiguous(), "z_vals")
        rays = _lib.require_cuda(rays.contiguous(), "rays")
        N, S, out_dim = raw.shape
        acc = torch.empty(N, device=raw.device, dtype=torch.float32)
        absdiff = torch.empty(N, S, device=raw.device, dtype=torch.float32)
        with torch.cuda.device(raw.device):
            _lib.check(L_.nafb_ray_integral_forward(_lib.ptr(raw), out_dim, _lib.ptr(z_vals), _lib.ptr(rays), _lib.ptr(acc),
                                                    _lib.ptr(absdiff), N, S, _lib.stream_ptr()))
        ctx.save_for_backward(z_vals, rays)
        ctx.shape = (N, S, out_dim)
        ctx.mark_non_differentiable(absdiff)
        return acc, absdiff

    @staticmethod
    def backward(ctx, dacc, _dabs):
        L_ = _lib.lib()
        z_vals, rays = ctx.saved_tensors
        N, S, out_dim = ctx.shape
        dacc = _lib.require_cuda(dacc.contiguous(), "grad")
        draw = torch.empty(N, S, out_dim, device=dacc.device, dtype=torch.float32)
        with torch.cuda.device(dacc.device):
            _lib.check(L_.nafb_ray_integral_backward(_lib.ptr(dacc), out_dim, _lib.ptr(z_vals), _lib.ptr(rays), _lib.ptr(draw), N, S,
                                                     _lib.stream_ptr()))
        return draw, None, None


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0.):
    """Beer-Lambert line integral of the predicted attenuation (render.py:178-212).

    raw [N,S,out_dim], z_vals [N,S], rays_d [N,3] -> acc [N], weights [N,S]."""
    if raw.shape[-1] not in (1, 2):
        raise NotImplementedError("Wrong raw shape")
    clean = raw
    if raw_noise_std > 0.:
        # the noise enters the line integral only (render.py:196-201); the weights below are taken from the un-noised prediction
        noise = (torch.randn(raw[..., 0].shape) * raw_noise_std).to(raw.device)  # render.py:197-199 (CPU generator)
        raw = torch.cat([(raw[..., :1] + noise[..., None]), raw[..., 1:]], -1)
    rays = torch.zeros(rays_d.shape[0], 8, device=rays_d.device, dtype=torch.float32)
    rays[:, 3:6] = rays_d
    acc, absdiff = _RayIntegral.apply(raw, z_vals, rays)
    if clean.shape[-1] == 1:
        if clean is not raw:
            absdiff = torch.cat([torch.full_like(clean[:, :1, -1], 1e-10), torch.abs(clean[:, 1:, -1] - clean[:, :-1, -1])], dim=-1).detach()
        weights = absdiff / torch.max(absdiff)
    else:  # with jac
        weights = clean[..., 1] / torch.max(clean[..., 1])
    return acc, weights


def run_network(inputs, fn, netchunk):
    """Evaluate ``fn`` on [..., 3] points in chunks of ``netchunk`` (render.py:148-156)."""
    flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    if isinstance(fn, DensityNetwork) and fn.fused_meta() is not None and flat.is_cuda:
        out_flat = fn(flat)  # one fused launch; chunking exists only to bound eager-mode memory
    else:
        out_flat = torch.cat([fn(flat[i:i + netchunk]) for i in range(0, flat.shape[0], netchunk)], 0)
    return out_flat.reshape(list(inputs.shape[:-1]) + [out_flat.shape[-1]])


def sample_pdf(bins, weights, N_samples, det=False):
    """Inverse-CDF sampling for the fine pass (render.py:215-247)."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    shape = list(cdf.shape[:-1]) + [N_samples]
    if det:
        u = torch.linspace(0., 1., steps=N_samples).expand(shape)
    else:
        u = torch.rand(shape)
    u = u.contiguous().to(cdf.device)
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_lo, bin_hi = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_lo) / denom
    return bin_lo + t * (bin_hi - bin_lo)


def sample_fine(rays, z_vals, weights, n_fine, det, bound, want_tv=True):
    """The fine pass's sampling stage as ONE kernel (render.py:113-126): sample_pdf over the midpoints of z_vals with
    weights[..., 1:-1], merge + sort with the coarse depths, clamped sample positions, per-ray TV partial sums.
    The uniforms are drawn exactly as the reference draws them (torch.rand on the CPU generator, or linspace when det)."""
    import numpy as np
    L_ = _lib.lib()
    rays = _lib.require_cuda(rays.contiguous(), "rays")
    z_vals = _lib.require_cuda(z_vals.contiguous(), "z_vals")
    weights = _lib.require_cuda(weights.detach().contiguous(), "weights")
    N, S = z_vals.shape
    dev = rays.device
    if det:
        u, u_stride = torch.linspace(0., 1., steps=n_fine).to(dev), 0                    # render.py:228-229
    else:
        u, u_stride = torch.rand([N, n_fine]).contiguous().to(dev), n_fine               # render.py:230-233 (CPU generator)
    M = S + n_fine
    z = torch.empty(N, M, device=dev, dtype=torch.float32)
    pts = torch.empty(N, M, 3, device=dev, dtype=torch.float32)
    tv = torch.empty(N, device=dev, dtype=torch.float32) if want_tv else None
    with torch.cuda.device(dev):
        _lib.check(L_.nafb_sample_fine(_lib.ptr(rays), _lib.ptr(z_vals), _lib.ptr(weights), _lib.ptr(u), u_stride, N, S, int(n_fine),
                                       float(np.float32(bound - 1e-6)), _lib.ptr(z), _lib.ptr(pts), _lib.ptr(tv), _lib.stream_ptr()))
    return z, pts, tv


def render(rays, net, net_fine, n_samples, n_fine, perturb, netchunk, raw_noise_std, chunk_size=None):
    """Render projections for ``rays`` [N, 8] (origin, direction, near, far).

    Returns {"acc": [N], "pts": [N, S, 3], "tv_loss": scalar} (un-chunked) or, when the rays are
    processed in ``chunk_size`` pieces, {"acc", "pts"} (+ "acc0", "weights0", "pts0" with a fine
    network) -- exactly the keys of the reference (render.py:31-79)."""
    n_rays = rays.shape[0]
    if chunk_size is None or chunk_size >= n_rays:
        return render_chunk(rays, net, net_fine, n_samples, n_fine, perturb, netchunk, raw_noise_std)
    parts = [render_chunk(rays[i:i + chunk_size], net, net_fine, n_samples, n_fine, perturb, netchunk, raw_noise_std)
             for i in range(0, n_rays, chunk_size)]
    keys = ["acc", "pts"] + (["acc0", "weights0", "pts0"] if "acc0" in parts[0] else [])
    return {k: torch.cat([p[k] for p in parts], dim=0) for k in keys}


def render_chunk(rays, net, net_fine, n_samples, n_fine, perturb, netchunk, raw_noise_std):
    if not rays.is_cuda:
        raise RuntimeError("rays must be a CUDA tensor")
    n_rays = rays.shape[0]
    rays = rays.contiguous()
    use_fine = net_fine is not None and n_fine > 0
    t_rand = torch.rand([n_rays, n_samples], device=rays.device) if perturb else None  # render.py:99

    meta = net.fused_meta() if isinstance(net, DensityNetwork) else None
    if meta is not None and not use_fine and not raw_noise_std > 0. and rays.dtype == torch.float32:
        acc, pts, z_vals = RenderFn.apply(rays, t_rand, net.encoder.embeddings, meta, int(n_samples), bool(perturb), *net.flat_params())
        tv_loss = _tv_partial(rays, n_samples, perturb, t_rand, net.bound).sum() * 0.1
        ret = {"acc": acc, "pts": pts, "tv_loss": tv_loss}
    else:
        z_vals, pts, tv = sample_points(rays, n_samples, perturb, t_rand, net.bound, want_tv=not use_fine)
        raw = run_network(pts, net, netchunk)
        acc, weights = raw2outputs(raw, z_vals, rays[..., 3:6], raw_noise_std)
        ret = {}
        if use_fine:
            ret.update(acc0=acc, weights0=weights, pts0=pts)
            if FUSED_FINE_SAMPLING and rays.dtype == torch.float32 and 3 <= n_samples and n_samples + n_fine <= 1024:
                z_vals, pts, tv = sample_fine(rays, z_vals, weights, n_fine, perturb == 0., net.bound)   # one kernel
                tv_loss = tv.sum() * 0.1
            else:
                z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
                z_samples = sample_pdf(z_mid, weights[..., 1:-1], n_fine, det=(perturb == 0.)).detach()
                z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
                bound = net.bound - 1e-6
                pts = (rays[..., None, :3] + rays[..., None, 3:6] * z_vals[..., :, None]).clamp(-bound, bound)
                tv_loss = compute_tv_regularization(pts) * 0.1
            raw = run_network(pts, net_fine, netchunk)
            acc, _ = raw2outputs(raw, z_vals, rays[..., 3:6], raw_noise_std)
        else:
            tv_loss = tv.sum() * 0.1
        ret.update(acc=acc, pts=pts, tv_loss=tv_loss)

    if CHECK_NUMERICS:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret
