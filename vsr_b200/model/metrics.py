from ..metrics import PSNR, SSIM, CardiacPSNR, CardiacSSIM  # noqa: F401
