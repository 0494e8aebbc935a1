"""Drop-in ``box_blur`` and ``BoxBlur`` (reference: kornia/filters/blur.py:29-76,79-130)."""
from __future__ import annotations

import torch
from torch import nn

from ..core.check import check_is_tensor
from .filter import filter2d, filter2d_separable
from .kernels import _unpack_2d_ks, get_box_kernel1d, get_box_kernel2d

__all__ = ["box_blur", "BoxBlur"]


def box_blur(input: torch.Tensor, kernel_size: tuple[int, int] | int, border_type: str = "reflect",
             separable: bool = False) -> torch.Tensor:
    """Mean filter over a ``kernel_size`` window of every channel of ``input`` (B,C,H,W).
    ``separable=True`` runs the one-pass row+column kernel instead of the 2-D one."""
    check_is_tensor(input)
    if separable:
        ky, kx = _unpack_2d_ks(kernel_size)
        kernel_y = get_box_kernel1d(ky, device=input.device, dtype=input.dtype)
        kernel_x = get_box_kernel1d(kx, device=input.device, dtype=input.dtype)
        return filter2d_separable(input, kernel_x, kernel_y, border_type)
    return filter2d(input, get_box_kernel2d(kernel_size, device=input.device, dtype=input.dtype), border_type)


class BoxBlur(nn.Module):
    """Module form of :func:`box_blur` (same constructor as the reference's)."""

    def __init__(self, kernel_size: tuple[int, int] | int, border_type: str = "reflect", separable: bool = False) -> None:
        super().__init__()
        self.kernel_size = kernel_size
        self.border_type = border_type
        self.separable = separable

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(kernel_size={self.kernel_size}, border_type={self.border_type}, "
                f"separable={self.separable})")

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return box_blur(input, self.kernel_size, self.border_type, self.separable)
