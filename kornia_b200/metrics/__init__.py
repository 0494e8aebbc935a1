from .ssim import SSIM, ssim

__all__ = ["SSIM", "ssim"]
