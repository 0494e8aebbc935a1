"""Drop-in ``laplacian`` and ``Laplacian`` (reference: kornia/filters/laplacian.py:27-63,66-110)."""
from __future__ import annotations

import torch
from torch import nn

from .filter import filter2d
from .kernels import get_laplacian_kernel2d, normalize_kernel2d

__all__ = ["laplacian", "Laplacian"]


def laplacian(input: torch.Tensor, kernel_size: tuple[int, int] | int, border_type: str = "reflect",
              normalized: bool = True) -> torch.Tensor:
    """Filter every channel of ``input`` (B,C,H,W) with the ones-and-centre Laplacian kernel;
    ``normalized`` scales the kernel to unit L1 norm first."""
    kernel = get_laplacian_kernel2d(kernel_size, device=input.device, dtype=input.dtype)[None, ...]
    if normalized:
        kernel = normalize_kernel2d(kernel)
    return filter2d(input, kernel, border_type)


class Laplacian(nn.Module):
    """Module form of :func:`laplacian` (same constructor as the reference's)."""

    def __init__(self, kernel_size: tuple[int, int] | int, border_type: str = "reflect", normalized: bool = True) -> None:
        super().__init__()
        self.kernel_size = kernel_size
        self.border_type = border_type
        self.normalized = normalized

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(kernel_size={self.kernel_size}, normalized={self.normalized}, "
                f"border_type={self.border_type})")

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return laplacian(input, self.kernel_size, self.border_type, self.normalized)
