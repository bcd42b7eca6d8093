from ..metrics import PSNR, SSIM  # noqa: F401
