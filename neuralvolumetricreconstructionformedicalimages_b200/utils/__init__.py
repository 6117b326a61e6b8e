from .metrics import get_mse, get_psnr, get_psnr_3d, get_ssim_3d  # noqa: F401
