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
