"""Kernel-tap generators on the gaussian_blur2d path (reference: kornia/filters/kernels.py:29-47,
68-74,77-120,552-585,661-715).  Tiny (B,k) tensors: kept in torch so the taps are bit-identical
to the reference's and gradients w.r.t. ``sigma`` come from autograd."""
from __future__ import annotations

from typing import Optional, Union

import torch

from ..core.check import check, check_is_tensor, check_shape


def _check_kernel_size(kernel_size, min_value: int = 0, allow_even: bool = False) -> None:
    sizes = (kernel_size,) if isinstance(kernel_size, int) else kernel_size
    kind = "even or odd" if allow_even else "odd"
    for s in sizes:
        check(isinstance(s, int) and ((s % 2 == 1) or allow_even) and s > min_value,
              f"Kernel size must be an {kind} integer bigger than {min_value}. Gotcha {s} on {sizes}")


def _unpack_2d_ks(kernel_size) -> tuple[int, int]:
    if isinstance(kernel_size, int):
        return kernel_size, kernel_size
    check(len(kernel_size) == 2, "2D Kernel size should have a length of 2.")
    return int(kernel_size[0]), int(kernel_size[1])


def normalize_kernel2d(kernel: torch.Tensor) -> torch.Tensor:
    """Divide every (..., kh, kw) kernel by the sum of the absolute values of its taps."""
    check_shape(kernel, ["*", "H", "W"])
    return kernel / kernel.abs().sum(dim=-1).sum(dim=-1)[..., None, None]


def gaussian(window_size: int, sigma, *, mean=None, device=None, dtype=None) -> torch.Tensor:
    """(B, window_size) samples of exp(-(x - mean)^2 / (2 sigma^2)), each row normalised to sum 1.
    ``sigma`` is a float or a (B,1) tensor; even windows are shifted by half a tap."""
    if isinstance(sigma, float):
        sigma = torch.tensor([[sigma]], device=device, dtype=dtype)
    check_is_tensor(sigma)
    check_shape(sigma, ["B", "1"])
    mean = float(window_size // 2) if mean is None else mean
    if isinstance(mean, float):
        mean = torch.tensor([[mean]], device=sigma.device, dtype=sigma.dtype)
    check_is_tensor(mean)
    check_shape(mean, ["B", "1"])
    x = (torch.arange(window_size, device=sigma.device, dtype=sigma.dtype) - mean).expand(sigma.shape[0], -1)
    if window_size % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma.pow(2.0)))
    return g / g.sum(-1, keepdim=True)


def get_gaussian_kernel1d(kernel_size: int, sigma: Union[float, torch.Tensor], force_even: bool = False, *,
                          device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    _check_kernel_size(kernel_size, allow_even=force_even)
    return gaussian(kernel_size, sigma, device=device, dtype=dtype)


def get_gaussian_kernel2d(kernel_size, sigma, force_even: bool = False, *, device=None, dtype=None) -> torch.Tensor:
    """(B, ky, kx) outer product of the two 1-D kernels; ``sigma`` is (sigma_y, sigma_x) or (B,2)."""
    if isinstance(sigma, tuple):
        sigma = torch.tensor([sigma], device=device, dtype=dtype)
    check_is_tensor(sigma)
    check_shape(sigma, ["B", "2"])
    ky, kx = _unpack_2d_ks(kernel_size)
    col = get_gaussian_kernel1d(ky, sigma[:, 0, None], force_even, device=device, dtype=dtype)[..., None]
    row = get_gaussian_kernel1d(kx, sigma[:, 1, None], force_even, device=device, dtype=dtype)[..., None]
    return col * row.view(-1, 1, kx)
