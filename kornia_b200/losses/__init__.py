from .ssim import SSIMLoss, ssim_loss

__all__ = ["SSIMLoss", "ssim_loss"]
