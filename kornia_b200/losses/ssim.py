"""Drop-in ``ssim_loss`` and ``SSIMLoss`` (reference: kornia/losses/ssim.py:26-82,85-130)."""
from __future__ import annotations

import torch
from torch import nn

from .. import metrics

__all__ = ["ssim_loss", "SSIMLoss"]


def ssim_loss(img1: torch.Tensor, img2: torch.Tensor, window_size: int, max_val: float = 1.0, eps: float = 1e-12,
              reduction: str = "mean", padding: str = "same") -> torch.Tensor:
    """Structural dissimilarity ``clamp((1 - ssim) / 2, 0, 1)``, reduced by 'mean' | 'sum' | 'none'."""
    loss = torch.clamp((1.0 - metrics.ssim(img1, img2, window_size, max_val, eps, padding)) / 2, min=0, max=1)
    if reduction == "mean":
        return torch.mean(loss)
    if reduction == "sum":
        return torch.sum(loss)
    if reduction == "none":
        return loss
    raise NotImplementedError("Invalid reduction option.")


class SSIMLoss(nn.Module):
    """Module form of :func:`ssim_loss` (same constructor as the reference's)."""

    def __init__(self, window_size: int, max_val: float = 1.0, eps: float = 1e-12, reduction: str = "mean",
                 padding: str = "same") -> None:
        super().__init__()
        self.window_size = window_size
        self.max_val = max_val
        self.eps = eps
        self.reduction = reduction
        self.padding = padding

    def forward(self, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
        return ssim_loss(img1, img2, self.window_size, self.max_val, self.eps, self.reduction, self.padding)
