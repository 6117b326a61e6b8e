from .geometry import ConeGeometry, angle2pose, chest50_like, get_near_far, get_rays, get_voxels, rays_with_near_far, voxel_half_extent  # noqa: F401
from .tigre import TIGREDataset  # noqa: F401,E402
