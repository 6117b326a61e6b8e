"""Analytic phantom: a sum of ellipsoids with closed-form line integrals.

The reference's projections come from the TIGRE forward projector (dataGenerator/generateData.py:153-211), which is not
available here; a phantom whose projections are exact line integrals gives consistent (volume, projections) pairs for
throughput runs, smoke tests and the PSNR parity test.  Lengths in metres, attenuation values in [0, 1] like the
normalised images of the reference's pickles (format_data.py:25-58).
"""
from __future__ import annotations

import numpy as np
import torch

from .geometry import ConeGeometry, get_voxels


def default_ellipsoids(half_extent: float):
    """A small Shepp-Logan-like arrangement inside a cube of the given half extent: (centre[3], semi-axes[3], value)."""
    h = float(half_extent)
    return [
        (np.array([0.0, 0.0, 0.0]) * h, np.array([0.80, 0.70, 0.75]) * h, 0.30),
        (np.array([0.25, 0.10, 0.0]) * h, np.array([0.25, 0.35, 0.40]) * h, 0.35),
        (np.array([-0.30, -0.05, 0.10]) * h, np.array([0.20, 0.30, 0.25]) * h, -0.20),
        (np.array([0.0, 0.40, -0.20]) * h, np.array([0.15, 0.12, 0.18]) * h, 0.30),
        (np.array([0.05, -0.45, 0.25]) * h, np.array([0.10, 0.10, 0.10]) * h, 0.25),
    ]


def phantom_volume(geo: ConeGeometry, ellipsoids) -> np.ndarray:
    """float32 [n1,n2,n3]: the phantom sampled at the voxel centres of tigre.py:388-400."""
    xyz = get_voxels(geo)
    vol = np.zeros(xyz.shape[:3], dtype=np.float64)
    for c, a, v in ellipsoids:
        vol += v * (np.sum(((xyz - c) / a) ** 2, axis=-1) <= 1.0)
    return vol.astype(np.float32)


def phantom_projections(rays: torch.Tensor, ellipsoids) -> torch.Tensor:
    """Exact line integrals of the phantom along rays [...,8] = (origin, direction, near, far) between near and far:
    sum_e value_e * |chord_e| * |d| (the direction is not normalised, tigre.py:434-437)."""
    r = rays.reshape(-1, 8).to(torch.float64)
    o, d, near, far = r[:, 0:3], r[:, 3:6], r[:, 6], r[:, 7]
    out = torch.zeros(r.shape[0], dtype=torch.float64, device=r.device)
    for c, a, v in ellipsoids:
        c_t = torch.as_tensor(c, dtype=torch.float64, device=r.device)
        a_t = torch.as_tensor(a, dtype=torch.float64, device=r.device)
        p, q = (o - c_t) / a_t, d / a_t
        A, B, C = (q * q).sum(-1), (p * q).sum(-1), (p * p).sum(-1) - 1.0
        disc = B * B - A * C
        root = torch.sqrt(torch.clamp(disc, min=0.0))
        t0, t1 = (-B - root) / A, (-B + root) / A
        chord = torch.clamp(torch.minimum(t1, far) - torch.maximum(t0, near), min=0.0)
        out += torch.where(disc > 0, v * chord, torch.zeros_like(chord))
    return (out * d.norm(dim=-1)).to(torch.float32).reshape(rays.shape[:-1])


# ----------------------------------------------------------------------------- reference pickle schema
def make_dataset_dict(geometry: dict, n_train: int, n_val: int, ellipsoids=None, amplitude: float = 1.0) -> dict:
    """A dataset in the schema the reference's TIGREDataset reads (format_data.py:25-58, src/dataset/tigre.py:230-302):
    geometry block + 'image' + 'full_proj' (complex, A*exp(i*phase): the laminography pipeline trains on the phase) +
    'train' / 'val' {angles, projections}.  Projections are exact line integrals of the analytic phantom.
    `geometry` carries DSD, DSO, nDetector, dDetector, nVoxel, dVoxel, offOrigin, offDetector, mode (+ tilt_angle)."""
    from .geometry import rays_with_near_far
    geo = ConeGeometry(geometry)
    ells = ellipsoids if ellipsoids is not None else default_ellipsoids(float(min(geo.sVoxel)) / 2)
    span = np.pi if geo.mode == "cone" else np.pi
    tr_angles = np.linspace(0, span, n_train + 1)[:-1]
    va_angles = np.linspace(0, span, n_val + 1)[:-1] + span / (2 * max(n_val, 1))

    def project(angles):
        rays = rays_with_near_far(angles, geo, "cpu")
        return phantom_projections(rays, ells).numpy().astype(np.float32)

    tr, va = project(tr_angles), project(va_angles)
    data = dict(geometry)
    data.update(numTrain=n_train, numVal=n_val, accuracy=geometry.get("accuracy", 0.5), filter=geometry.get("filter"),
                totalAngle=180, startAngle=0, randomAngle=False, convert=False, rescale_slope=1.0, rescale_intercept=0.0,
                normalize=True, noise=0, tilt_angle=geometry.get("tilt_angle", 0), image=phantom_volume(geo, ells),
                full_proj=(amplitude * np.exp(1j * tr)).astype(np.complex64),
                train={"angles": tr_angles, "projections": tr}, val={"angles": va_angles, "projections": va})
    return data


def save_pickle(data: dict, path: str) -> None:
    import pickle
    with open(path, "wb") as f:
        pickle.dump(data, f)


def load_pickle(path: str):
    """-> (data dict, ConeGeometry) of a dataset in the reference's schema."""
    import pickle
    with open(path, "rb") as f:
        data = pickle.load(f)
    return data, ConeGeometry(data)
