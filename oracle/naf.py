"""CPU restatement (torch-CPU / numpy, fp32, one rounding per elementary op) of the
Python half of the NAF hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference files restated (paths relative to /root/reference):
  src/render/render.py        sampling :88-105, raw2outputs :178-212, render :31-146, run_network :148-156,
                              hierarchical pass :113-126, sample_pdf :215-247, raw_noise_std :196-199
  src/network/network.py      DensityNetwork :5-58
  src/encoder/freqencoder.py  FreqEncoder :29-43
  src/loss/loss.py            calc_mse_loss :26-46
  src/dataset/tigre.py        ConeGeometry :183-217, get_voxels :388-400, get_rays :402-456,
                              get_rays2 :463-528, angle2pose :530-572, get_near_far :575-586
  src/utils/util.py           get_psnr_3d :55-84, get_ptycho_mask :196-205
  train.py                    compute_loss :48-135 (chunked masked MSE, with the :93 slice corrected)
"""
from __future__ import annotations

import math

import numpy as np
import torch

# --------------------------------------------------------------------------- sampling


def linspace01(steps: int) -> np.ndarray:
    """torch.linspace(0, 1, steps) as fp32 bit pattern, restated without torch.

    ATen evaluates the two-sided formula with a fused multiply-add:
        i <  steps//2 :  fma(step, i, 0)           == fl(step * i)
        i >= steps//2 :  fma(-step, steps-1-i, 1)  == fl(1 - step*(steps-1-i))   (one rounding)
    with step = fl(1 / (steps-1)).  The products are exact in float64, so a float64
    evaluation followed by one rounding to fp32 reproduces the FMA.
    """
    if steps == 1:
        return np.zeros(1, np.float32)
    step = np.float32(1.0) / np.float32(steps - 1)
    i = np.arange(steps, dtype=np.int64)
    lo = (np.float64(step) * i.astype(np.float64)).astype(np.float32)
    hi = (1.0 - np.float64(step) * (steps - 1 - i).astype(np.float64)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


def sample_points(rays: torch.Tensor, n_samples: int, perturb: bool, t_rand: torch.Tensor | None, bound: float):
    """render.py:88-105.  rays [N,8] = (o, d, near, far).  Returns z_vals [N,S], pts [N,S,3].

    Every product/sum is its own fp32 rounding (the reference is eager PyTorch):
        z   = near*(1-t) + far*t
        mid = .5*(z[i+1]+z[i]);  upper=[mid, z_last]; lower=[z_first, mid]
        z   = lower + (upper-lower)*t_rand
        pts = clamp(o + d*z, -(bound-1e-6), bound-1e-6)
    """
    rays = rays.to(torch.float32)
    n_rays = rays.shape[0]
    o, d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8]
    t = torch.from_numpy(linspace01(n_samples))
    z = near * (1.0 - t) + far * t
    z = z.expand(n_rays, n_samples)
    if perturb:
        mids = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    b = bound - 1e-6  # python double, cast to fp32 by clamp (render.py:104-105)
    pts = pts.clamp(-b, b)
    return z.contiguous(), pts.contiguous()


def ray_integral(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor):
    """render.py:192-212 with raw_noise_std == 0.  raw [N,S,out_dim].  -> acc [N], weights [N,S]."""
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    dists = torch.cat([dists, torch.full_like(dists[:, :1], 1e-10)], -1)
    dists = dists * torch.norm(rays_d[:, None, :], dim=-1)
    acc = torch.sum(raw[..., 0] * dists, dim=-1)
    if raw.shape[-1] == 1:
        eps = torch.ones_like(raw[:, :1, -1]) * 1e-10
        w = torch.cat([eps, torch.abs(raw[:, 1:, -1] - raw[:, :-1, -1])], dim=-1)
        w = w / torch.max(w)
    elif raw.shape[-1] == 2:
        w = raw[..., 1] / torch.max(raw[..., 1])
    else:
        raise NotImplementedError("Wrong raw shape")
    return acc, w


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, det: bool) -> torch.Tensor:
    """render.py:215-247: inverse-CDF sampling of the fine pass.  bins [N,B], weights [N,B-1] -> [N,n_samples].
    det: u = linspace(0,1); else u is drawn from torch's GLOBAL CPU generator, as the reference does."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    shape = list(cdf.shape[:-1]) + [n_samples]
    u = torch.linspace(0., 1., steps=n_samples).expand(shape) if det else torch.rand(shape)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_lo, bin_hi = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)


def ray_integral_noisy(raw, z_vals, rays_d, raw_noise_std):
    """render.py:192-212 with raw_noise_std > 0: the noise (global CPU generator) enters the line integral only, the weights are
    taken from the un-noised prediction."""
    noise = torch.randn(raw[..., 0].shape) * raw_noise_std if raw_noise_std > 0. else None
    acc, w = ray_integral(raw, z_vals, rays_d)
    if noise is not None:
        noisy = torch.cat([raw[..., :1] + noise[..., None], raw[..., 1:]], -1)
        acc, _ = ray_integral(noisy, z_vals, rays_d)
    return acc, w


def render_hierarchical(rays, net, net_fine, n_samples, n_fine, perturb, raw_noise_std=0., netchunk=409600):
    """render.py:82-146 with a fine network (n_fine > 0), single chunk.  Random draws come from torch's global generators in the
    reference's order: t_rand (:99), coarse noise (:197), u of sample_pdf (:226), fine noise (:197)."""
    t_rand = torch.rand([rays.shape[0], n_samples]) if perturb else None
    z, pts = sample_points(rays, n_samples, bool(perturb), t_rand, net.bound)
    raw = run_network(pts, net, netchunk)
    acc0, w0 = ray_integral_noisy(raw, z, rays[:, 3:6], raw_noise_std)
    z_mid = .5 * (z[:, 1:] + z[:, :-1])
    z_samples = sample_pdf(z_mid, w0[:, 1:-1], n_fine, det=(perturb == 0.)).detach()
    z_all, _ = torch.sort(torch.cat([z, z_samples], -1), -1)
    b = net.bound - 1e-6
    pts_all = (rays[:, None, 0:3] + rays[:, None, 3:6] * z_all[:, :, None]).clamp(-b, b)
    raw_f = run_network(pts_all, net_fine, netchunk)
    acc, _ = ray_integral_noisy(raw_f, z_all, rays[:, 3:6], raw_noise_std)
    return {"acc": acc, "pts": pts_all, "tv_loss": tv_of_points(pts_all), "acc0": acc0, "weights0": w0, "pts0": pts}


def tv_of_points(pts: torch.Tensor) -> torch.Tensor:
    """render.py:16-28 and :130-131: 0.1 * sum |pts[:,1:] - pts[:,:-1]|."""
    return torch.sum(torch.abs(pts[:, 1:, :] - pts[:, :-1, :])) * 0.1


def run_network(inputs, fn, netchunk):
    """render.py:148-156."""
    flat = inputs.reshape(-1, inputs.shape[-1])
    outs = [fn(flat[i:i + netchunk]) for i in range(0, flat.shape[0], netchunk)]
    out = torch.cat(outs, 0)
    return out.reshape(list(inputs.shape[:-1]) + [out.shape[-1]])


def render(rays, net, n_samples, perturb, netchunk=409600, t_rand=None):
    """render.py:31-146 for n_fine == 0, raw_noise_std == 0, single chunk."""
    z, pts = sample_points(rays, n_samples, perturb, t_rand, net.bound)
    raw = run_network(pts, net, netchunk)
    acc, w = ray_integral(raw, z, rays[:, 3:6])
    return {"acc": acc, "pts": pts, "tv_loss": tv_of_points(pts), "z_vals": z, "raw": raw, "weights": w}


# --------------------------------------------------------------------------- network


class OracleFreqEncoder(torch.nn.Module):
    """freqencoder.py:5-43: [x, sin(x*2^k), cos(x*2^k)] for k = 0..N_freqs-1; ``bound`` ignored."""

    def __init__(self, input_dim=3, multires=6):
        super().__init__()
        self.input_dim = input_dim
        self.freqs = (2.0 ** torch.linspace(0.0, multires - 1, multires)).tolist()
        self.output_dim = input_dim * (1 + 2 * multires)

    def forward(self, x, bound=None):
        parts = [x]
        for f in self.freqs:
            parts += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(parts, dim=-1)


class OracleDensityNetwork(torch.nn.Module):
    """network.py:5-58.  layers[i]: Linear; skip layers take cat([encoding, h]); LeakyReLU(0.01)
    after every hidden layer; head in {sigmoid, relu(=LeakyReLU), tanh, none}."""

    def __init__(self, encoder, bound=0.2, num_layers=8, hidden_dim=256, skips=(4,), out_dim=1, last_activation="sigmoid"):
        super().__init__()
        self.encoder = encoder
        self.in_dim = encoder.output_dim
        self.bound = bound
        self.skips = list(skips)
        dims_in = [self.in_dim] + [hidden_dim + (self.in_dim if i in self.skips else 0) for i in range(1, num_layers - 1)]
        self.layers = torch.nn.ModuleList([torch.nn.Linear(k, hidden_dim) for k in dims_in] + [torch.nn.Linear(hidden_dim, out_dim)])
        if last_activation not in ("sigmoid", "relu", "tanh", "none"):
            raise NotImplementedError("Unknown last activation")
        self.last_activation = last_activation

    def head(self, x):
        return {"sigmoid": torch.sigmoid, "relu": lambda v: torch.nn.functional.leaky_relu(v, 0.01),
                "tanh": torch.tanh, "none": lambda v: v}[self.last_activation](x)

    def forward(self, x):
        enc = self.encoder(x, self.bound)
        h = enc
        n = len(self.layers)
        for i, lin in enumerate(self.layers):
            if i in self.skips:
                h = torch.cat([enc, h], -1)
            h = lin(h)
            h = self.head(h) if i == n - 1 else torch.nn.functional.leaky_relu(h, 0.01)
        return h


# --------------------------------------------------------------------------- loss


def calc_mse_loss(loss: dict, x, y, tv_loss=None):
    """loss.py:26-46."""
    mse = torch.mean((x - y) ** 2)
    loss["loss"] = loss["loss"] + mse
    loss["loss_mse"] = mse
    if tv_loss is not None:
        loss["loss"] = loss["loss"] + tv_loss
        loss["tv_loss"] = tv_loss
    return loss


def chunked_masked_mse(pred, target, mask=None, chunk=200):
    """train.py:56-127 as intended (the :93 slice taken along the ray axis): the total loss is
    the SUM over ray chunks of the MEAN over the masked-in rays of that chunk."""
    total = torch.zeros((), dtype=pred.dtype)
    n = pred.shape[0]
    chunk = n if chunk is None else chunk
    for i in range(0, n, chunk):
        p, t = pred[i:i + chunk], target[i:i + chunk]
        if mask is not None:
            m = mask[i:i + chunk].bool()
            p, t = p[m], t[m]
        total = total + torch.mean((t - p) ** 2)
    return total


def ptycho_mask(hr: torch.Tensor, threshold=0.007) -> torch.Tensor:
    """util.py:196-205.  NB the in-place ANDs read the already-updated mask:
    row pass first (m[1:] &= m[1:] == m[:-1] evaluated on the pre-pass copy of the RHS),
    then the column pass on the result.  Returns ~m (True = keep the pixel)."""
    m = torch.abs(hr) < threshold
    m = m.clone()
    rhs = (m[1:] == m[:-1]).clone()
    m[1:] &= rhs
    rhs = (m[:, 1:] == m[:, :-1]).clone()
    m[:, 1:] &= rhs
    return ~m


# --------------------------------------------------------------------------- geometry


class Geometry:
    """tigre.py:183-217 (mm -> m)."""

    def __init__(self, data: dict):
        self.DSD = data["DSD"] / 1000.0
        self.DSO = data["DSO"] / 1000.0
        self.nDetector = np.array(data["nDetector"])
        self.dDetector = np.array(data["dDetector"]) / 1000.0
        self.sDetector = self.nDetector * self.dDetector
        self.nVoxel = np.array(data["nVoxel"])
        self.dVoxel = np.array(data["dVoxel"]) / 1000.0
        self.sVoxel = self.nVoxel * self.dVoxel
        self.offOrigin = np.array(data["offOrigin"]) / 1000.0
        self.offDetector = np.array(data["offDetector"]) / 1000.0
        self.mode = data["mode"]
        self.tilt_angle = data.get("tilt_angle", 0)


def angle2pose(DSO: float, angle: float, tilt_deg: float) -> np.ndarray:
    """tigre.py:530-572: rot = Rz(angle) Rz(+90deg) Rx(-90deg) Rx_cw(tilt); trans = DSO*(cos, sin, tan tilt)."""
    def rx(p):
        return np.array([[1.0, 0, 0], [0, np.cos(p), -np.sin(p)], [0, np.sin(p), np.cos(p)]])

    def rz(p):
        return np.array([[np.cos(p), -np.sin(p), 0], [np.sin(p), np.cos(p), 0], [0, 0, 1.0]])

    tilt = np.radians(tilt_deg)
    rot = (rz(angle) @ rz(np.pi / 2)) @ rx(-np.pi / 2)
    rot = rot @ rx(-tilt)  # "clockwise" about x == rx(-tilt)
    T = np.eye(4)
    T[:3, :3] = rot
    T[:3, 3] = [DSO * np.cos(angle), DSO * np.sin(angle), DSO * np.tan(tilt)]
    return T


def detector_grid(geo: Geometry):
    """tigre.py:422-429: uu[row, col] = (col + .5 - W/2) dDet0 + offDet0 ; vv[row, col] = (row + .5 - H/2) dDet1 + offDet1 (fp32 ops)."""
    W, H = int(geo.nDetector[0]), int(geo.nDetector[1])
    col = torch.linspace(0, W - 1, W)
    row = torch.linspace(0, H - 1, H)
    uu = ((col + 0.5 - W / 2) * geo.dDetector[0] + geo.offDetector[0])[None, :].expand(H, W)
    vv = ((row + 0.5 - H / 2) * geo.dDetector[1] + geo.offDetector[1])[:, None].expand(H, W)
    return uu.to(torch.float32), vv.to(torch.float32)


def get_rays(angles, geo: Geometry) -> torch.Tensor:
    """tigre.py:402-456 (cone + parallel) / :463-528 (parallel).  -> [P, H, W, 6] fp32 (o, d)."""
    uu, vv = detector_grid(geo)
    out = []
    for a in angles:
        pose = torch.Tensor(angle2pose(geo.DSO, a, geo.tilt_angle))
        R, t = pose[:3, :3], pose[:3, 3]
        if geo.mode == "cone":
            dirs = torch.stack([uu / geo.DSD, vv / geo.DSD, torch.ones_like(uu)], -1)
            d = torch.matmul(R, dirs[..., None]).squeeze(-1)
            o = t.expand(d.shape)
        elif geo.mode == "parallel":
            dirs = torch.stack([torch.zeros_like(uu), torch.zeros_like(uu), torch.ones_like(uu)], -1)
            d = torch.matmul(R, dirs[..., None]).squeeze(-1)
            o = torch.matmul(R, torch.stack([uu, vv, torch.zeros_like(uu)], -1)[..., None]).squeeze(-1) + t.expand(d.shape)
        else:
            raise NotImplementedError("Unknown CT scanner type!")
        out.append(torch.cat([o, d], -1))
    return torch.stack(out, 0)


def get_near_far(geo: Geometry, tolerance=0.005):
    """tigre.py:575-586."""
    dmax = 0.0
    for sx in (-1, 1):
        for sy in (-1, 1):
            dmax = max(dmax, float(np.linalg.norm([geo.offOrigin[0] + sx * geo.sVoxel[0] / 2, geo.offOrigin[1] + sy * geo.sVoxel[1] / 2])))
    near = max(0.0, geo.DSO - dmax - tolerance)
    far = min(geo.DSO * 2, geo.DSO + dmax + tolerance)
    return near, far


def get_voxels(geo: Geometry) -> np.ndarray:
    """tigre.py:388-400 -> float64 [n1,n2,n3,3] (cast to fp32 by the caller, tigre.py:277)."""
    n1, n2, n3 = [int(v) for v in geo.nVoxel]
    s1, s2, s3 = geo.sVoxel / 2 - geo.dVoxel / 2
    xyz = np.meshgrid(np.linspace(-s1, s1, n1), np.linspace(-s2, s2, n2), np.linspace(-s3, s3, n3), indexing="ij")
    return np.asarray(xyz).transpose([1, 2, 3, 0])


def psnr_3d(a, b, pixel_max=1.0) -> float:
    """util.py:55-84 (float64, single volume)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return 100.0
    return float(20 * np.log10(pixel_max / np.sqrt(mse)))


# --------------------------------------------------------------------------- training step (CPU baseline)


def train_step(net, optimizer, rays, projs, n_samples, perturb, mask=None, chunk=None, t_rand=None, netchunk=409600):
    """One complete iteration as src/trainer.py:134-142 + train.py:48-135 define it
    (zero_grad, render, masked chunked MSE, backward, Adam)."""
    optimizer.zero_grad()
    if t_rand is None and perturb:
        t_rand = torch.rand(rays.shape[0], n_samples)
    ret = render(rays, net, n_samples, perturb, netchunk, t_rand)
    loss = chunked_masked_mse(ret["acc"], projs, mask, chunk)
    loss.backward()
    optimizer.step()
    return loss.detach()


def ssim_3d(a, b, data_range=2.0, win_size=7) -> float:
    """N-dimensional SSIM as skimage.metrics.structural_similarity evaluates it for a 3-D array without channel axis
    (what util.py:87-139 averages three times): scipy uniform_filter (mode 'reflect'), sample covariance, K1 .01, K2 .03,
    mean over the interior.  skimage itself is absent from this image; this is an independent scipy evaluation."""
    from scipy.ndimage import uniform_filter
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    NP = win_size ** a.ndim
    cov_norm = NP / (NP - 1.0)
    f = lambda t: uniform_filter(t, size=win_size, mode="reflect")
    ux, uy = f(a), f(b)
    uxx, uyy, uxy = f(a * a), f(b * b), f(a * b)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad, pad:-pad].mean())
