"""Scanner geometry and ray generation (host-side initialisation code, not the per-step path).

Mirrors the reference's src/dataset/tigre.py: ConeGeometry (:183-217, mm -> m), angle2pose
(:530-572, with the laminography tilt), get_rays for cone and parallel beams (:402-456 /
:463-528), get_near_far (:575-586) and get_voxels (:388-400).  Poses are float64 numpy, ray
grids are fp32 torch ops on whatever device is asked for (each op rounds on its own, like the
reference), so the rays are bit-identical to the reference's.
"""
from __future__ import annotations

import numpy as np
import torch


class ConeGeometry:
    """Cone-beam / parallel-beam CT geometry; lengths converted from millimetres to metres."""

    def __init__(self, data: dict):
        mm = 1000.0
        self.DSD = data["DSD"] / mm
        self.DSO = data["DSO"] / mm
        self.nDetector = np.array(data["nDetector"])
        self.dDetector = np.array(data["dDetector"]) / mm
        self.sDetector = self.nDetector * self.dDetector
        self.nVoxel = np.array(data["nVoxel"])
        self.dVoxel = np.array(data["dVoxel"]) / mm
        self.sVoxel = self.nVoxel * self.dVoxel
        self.offOrigin = np.array(data["offOrigin"]) / mm
        self.offDetector = np.array(data["offDetector"]) / mm
        self.accuracy = data.get("accuracy", 0.5)
        self.mode = data["mode"]
        self.filter = data.get("filter")
        self.magnification = 1
        self.tilt_angle = data.get("tilt_angle", 0)  # degrees


def _rot_x(p):
    c, s = np.cos(p), np.sin(p)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def _rot_z(p):
    c, s = np.cos(p), np.sin(p)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def angle2pose(DSO: float, angle: float, tilt_angle: float = 0.0) -> np.ndarray:
    """Source pose for a scan angle (rad) and a laminography tilt (DEGREES): float64 [4,4]."""
    tilt = np.radians(tilt_angle)
    rot = np.dot(np.dot(_rot_z(angle), _rot_z(np.pi / 2)), _rot_x(-np.pi / 2))
    rot = rot @ _rot_x(-tilt)  # clockwise about x
    T = np.eye(4)
    T[:3, :3] = rot
    T[:3, 3] = [DSO * np.cos(angle), DSO * np.sin(angle), DSO * np.tan(tilt)]
    return T


def get_rays(angles, geo: ConeGeometry, device="cpu") -> torch.Tensor:
    """[P, H, W, 6] fp32 (origin, direction) for every detector pixel of every projection."""
    W, H = int(geo.nDetector[0]), int(geo.nDetector[1])
    col = torch.linspace(0, W - 1, W, device=device)
    row = torch.linspace(0, H - 1, H, device=device)
    uu = ((col + 0.5 - W / 2) * geo.dDetector[0] + geo.offDetector[0])[None, :].expand(H, W)
    vv = ((row + 0.5 - H / 2) * geo.dDetector[1] + geo.offDetector[1])[:, None].expand(H, W)
    one, zero = torch.ones_like(uu), torch.zeros_like(uu)
    out = []
    for a in angles:
        pose = torch.Tensor(angle2pose(geo.DSO, float(a), geo.tilt_angle)).to(device)
        R, t = pose[:3, :3], pose[:3, 3]
        if geo.mode == "cone":
            dirs = torch.stack([uu / geo.DSD, vv / geo.DSD, one], -1)
            d = torch.matmul(R, dirs[..., None]).squeeze(-1)
            o = t.expand(d.shape)
        elif geo.mode == "parallel":
            d = torch.matmul(R, torch.stack([zero, zero, one], -1)[..., None]).squeeze(-1)
            o = torch.matmul(R, torch.stack([uu, vv, zero], -1)[..., None]).squeeze(-1) + t.expand(d.shape)
        else:
            raise NotImplementedError("Unknown CT scanner type!")
        out.append(torch.cat([o, d], dim=-1))
    return torch.stack(out, 0)


def get_near_far(geo: ConeGeometry, tolerance=0.005):
    """One global (near, far) pair from the xy-extent of the volume."""
    corners = [np.linalg.norm([geo.offOrigin[0] + sx * geo.sVoxel[0] / 2, geo.offOrigin[1] + sy * geo.sVoxel[1] / 2])
               for sx in (-1, 1) for sy in (-1, 1)]
    dist_max = np.max(corners)
    near = np.max([0, geo.DSO - dist_max - tolerance])
    far = np.min([geo.DSO * 2, geo.DSO + dist_max + tolerance])
    return near, far


def voxel_half_extent(geo: ConeGeometry):
    """End points of the float64 linspace that defines voxel centres: sVoxel/2 - dVoxel/2."""
    return geo.sVoxel / 2 - geo.dVoxel / 2


def get_voxels(geo: ConeGeometry) -> np.ndarray:
    """float64 [n1,n2,n3,3] voxel centres.  (The engine's voxel_query never materialises this.)"""
    n1, n2, n3 = [int(v) for v in geo.nVoxel]
    s1, s2, s3 = voxel_half_extent(geo)
    xyz = np.meshgrid(np.linspace(-s1, s1, n1), np.linspace(-s2, s2, n2), np.linspace(-s3, s3, n3), indexing="ij")
    return np.asarray(xyz).transpose([1, 2, 3, 0])


def rays_with_near_far(angles, geo: ConeGeometry, device="cpu") -> torch.Tensor:
    """[P, H, W, 8]: rays + the global near / far columns (tigre.py:248-255)."""
    rays = get_rays(angles, geo, device)
    near, far = get_near_far(geo)
    return torch.cat([rays, torch.ones_like(rays[..., :1]) * near, torch.ones_like(rays[..., :1]) * far], dim=-1)


def chest50_like(n_voxel=128, n_detector=256, n_proj=50) -> dict:
    """Synthetic stand-in for data/chest_50.pickle's geometry block:
    cone beam, DSD 1500 mm, DSO 1000 mm, 1 mm pixels/voxels, angles linspace(0, pi, n+1)[:-1]."""
    return dict(DSD=1500.0, DSO=1000.0, nDetector=[n_detector, n_detector], dDetector=[1.0, 1.0], nVoxel=[n_voxel] * 3,
                dVoxel=[128.0 / n_voxel] * 3, offOrigin=[0, 0, 0], offDetector=[0, 0], accuracy=0.5, mode="cone", filter=None,
                angles=np.linspace(0, np.pi, n_proj + 1)[:-1])


# ----------------------------------------------------------------------------- in-kernel ray generation
def pose_table(angles, geo: ConeGeometry, device="cpu") -> torch.Tensor:
    """[P, 12] fp32: rotation (3x3 row-major) | translation of every projection -- the fp32 cast of angle2pose that
    get_rays applies (``torch.Tensor(pose)``, tigre.py:479).  Input of the kernels' pixel source (nafb_sampler.poses)."""
    out = np.stack([np.concatenate([angle2pose(geo.DSO, float(a), geo.tilt_angle)[:3, :3].reshape(-1),
                                    angle2pose(geo.DSO, float(a), geo.tilt_angle)[:3, 3]]) for a in angles])
    return torch.from_numpy(out.astype(np.float32)).to(device).contiguous()


def detector_fields(geo: ConeGeometry) -> dict:
    """The scalar fields of nafb_sampler that describe the detector; every value is the fp32 cast torch applies when the
    reference multiplies / adds these numpy float64 scalars to fp32 tensors (tigre.py:428-429, 434, 248-255)."""
    near, far = get_near_far(geo)
    f = lambda v: float(np.float32(v))
    return dict(det_w=int(geo.nDetector[0]), det_h=int(geo.nDetector[1]), det_du=f(geo.dDetector[0]), det_dv=f(geo.dDetector[1]),
                det_u0=f(geo.offDetector[0]), det_v0=f(geo.offDetector[1]), det_dsd=f(geo.DSD), det_near=f(near), det_far=f(far),
                det_parallel=1 if geo.mode == "parallel" else 0)
