"""Drop-in ``unsharp_mask`` and ``UnsharpMask`` (reference: kornia/filters/unsharp.py:27-54,57-95)."""
from __future__ import annotations

import torch
from torch import nn

from .gaussian import gaussian_blur2d

__all__ = ["unsharp_mask", "UnsharpMask"]


def unsharp_mask(input: torch.Tensor, kernel_size: tuple[int, int] | int, sigma: tuple[float, float] | torch.Tensor,
                 border_type: str = "reflect") -> torch.Tensor:
    """Sharpen: ``2 * input - gaussian_blur2d(input)``, evaluated as the reference's
    ``lerp(blur, input, 2)`` so the rounding matches."""
    blurred = gaussian_blur2d(input, kernel_size, sigma, border_type)
    return torch.lerp(blurred, input, weight=2.0)


class UnsharpMask(nn.Module):
    """Module form of :func:`unsharp_mask` (same constructor as the reference's)."""

    def __init__(self, kernel_size: tuple[int, int] | int, sigma: tuple[float, float] | torch.Tensor,
                 border_type: str = "reflect") -> None:
        super().__init__()
        self.kernel_size = kernel_size
        self.sigma = sigma
        self.border_type = border_type

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return unsharp_mask(input, self.kernel_size, self.sigma, self.border_type)
