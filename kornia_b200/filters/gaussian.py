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

    sigma_key = None
    if isinstance(sigma, tuple):
        # constant sigmas: validate on the host (same verdict as the reference's device-side
        # `(sigma > 0).all()`, without its device->host sync) and reuse the taps across calls
        sigma_key = tuple(float(v) for v in sigma)
        if len(sigma_key) != 2:  # let the reference's shape check word the error
            check_shape(torch.tensor([sigma_key]), ["B", "2"])
        ok = all(v > 0 for v in sigma_key)
        sigma_t = None
    else:
        check_is_tensor(sigma)
        sigma_t = sigma.to(device=input.device, dtype=input.dtype)
        check_shape(sigma_t, ["B", "2"])
        ok = True
        if not torch.compiler.is_compiling():
            ok = bool((sigma_t > 0).all())  # device->host sync, as in the reference (gaussian.py:107)
    if sigma_key is not None and not ok:
        sigma_t = torch.tensor([sigma_key], device=input.device, dtype=input.dtype)
    check(ok, "sigma must be positive" if ok else f"sigma must be positive, got {sigma_t}")

    if sigma_key is not None:
        kernels = _constant_taps(kernel_size, sigma_key, separable, input.device, input.dtype)
    else:
        kernels = _taps(kernel_size, sigma_t, separable)
    if separable:
        return filter2d_separable(input, kernels[0], kernels[1], border_type)
    return filter2d(input, kernels[0], border_type)


def _taps(kernel_size, sigma: torch.Tensor, separable: bool):
    """(kernel_x, kernel_y) or (kernel2d,) built with the reference's torch ops (gaussian.py:111-117)."""
    if separable:
        ky, kx = _unpack_2d_ks(kernel_size)
        bs = sigma.shape[0]
        return (get_gaussian_kernel1d(kx, sigma[:, 1].view(bs, 1)), get_gaussian_kernel1d(ky, sigma[:, 0].view(bs, 1)))
    return (get_gaussian_kernel2d(kernel_size, sigma),)


_TAPS_CACHE: dict = {}


def _constant_taps(kernel_size, sigma_key, separable, device, dtype):
    """Taps of constant sigmas, cached per (size, sigma, device, dtype).  Built outside the caller's inference-mode / grad
    scope so that a tensor cached during an eval pass can still be saved for a later backward; rebuilt in-graph under
    ``torch.compile`` tracing (no global state there)."""
    if torch.compiler.is_compiling():
        return _taps(kernel_size, torch.tensor([sigma_key], device=device, dtype=dtype), separable)
    ks = kernel_size if isinstance(kernel_size, int) else tuple(kernel_size)
    key = (ks, sigma_key, bool(separable), str(device), dtype)
    hit = _TAPS_CACHE.get(key)
    if hit is None:
        if len(_TAPS_CACHE) >= 256:
            _TAPS_CACHE.clear()
        with torch.inference_mode(False), torch.no_grad():
            hit = _TAPS_CACHE[key] = _taps(kernel_size, torch.tensor([sigma_key], device=device, dtype=dtype), separable)
    return hit


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
