"""Drop-in ``gaussian_blur2d`` and the ``GaussianBlur2d`` module (reference:
kornia/filters/gaussian.py:32-120,123-200)."""
from __future__ import annotations

import torch
from torch import nn

from ..core.check import check, check_is_tensor, check_shape
from .filter import filter2d, filter2d_separable
from .kernels import _check_kernel_size, _unpack_2d_ks, get_gaussian_kernel1d, get_gaussian_kernel2d

__all__ = ["gaussian_blur2d", "GaussianBlur2d"]


def gaussian_blur2d(
    input: torch.Tensor,
    kernel_size: tuple[int, int] | int,
    sigma: tuple[float, float] | torch.Tensor,
    border_type: str = "reflect",
    separable: bool = True,
) -> torch.Tensor:
    """Gaussian-blur every channel of ``input`` (B,C,H,W).

    ``kernel_size``: odd int or (ky, kx); ``sigma``: (sigma_y, sigma_x) floats or a (B,2) tensor
    (positive).  ``separable=True`` (default) runs the fused row+column kernel, ``False`` the
    single 2-D kernel.  Note the reference's axis convention: the horizontal taps come from
    ``sigma[:, 1]``, the vertical ones from ``sigma[:, 0]`` (gaussian.py:113-114).
    """
    check_is_tensor(input)
    check_shape(input, ["B", "C", "H", "W"])
    _check_kernel_size(kernel_size, min_value=0)

    if isinstance(sigma, tuple):
        sigma = torch.tensor([sigma], device=input.device, dtype=input.dtype)
    else:
        check_is_tensor(sigma)
        sigma = sigma.to(device=input.device, dtype=input.dtype)
    check_shape(sigma, ["B", "2"])
    if not torch.compiler.is_compiling():
        ok = bool((sigma > 0).all())  # device->host sync, as in the reference (gaussian.py:107)
        check(ok, "sigma must be positive" if ok else f"sigma must be positive, got {sigma}")

    if separable:
        ky, kx = _unpack_2d_ks(kernel_size)
        bs = sigma.shape[0]
        kernel_x = get_gaussian_kernel1d(kx, sigma[:, 1].view(bs, 1))
        kernel_y = get_gaussian_kernel1d(ky, sigma[:, 0].view(bs, 1))
        return filter2d_separable(input, kernel_x, kernel_y, border_type)
    return filter2d(input, get_gaussian_kernel2d(kernel_size, sigma), border_type)


class GaussianBlur2d(nn.Module):
    """Module form of :func:`gaussian_blur2d` (same constructor as the reference's)."""

    def __init__(self, kernel_size, sigma, border_type: str = "reflect", separable: bool = True) -> None:
        super().__init__()
        self.kernel_size = kernel_size
        self.sigma = sigma
        self.border_type = border_type
        self.separable = separable

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(kernel_size={self.kernel_size}, sigma={self.sigma}, "
                f"border_type={self.border_type}, separable={self.separable})")

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return gaussian_blur2d(input, self.kernel_size, self.sigma, self.border_type, self.separable)
