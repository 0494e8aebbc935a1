"""Drop-in ``unsharp_mask`` and ``UnsharpMask`` (reference: kornia/filters/unsharp.py:27-54,57-95)."""
from __future__ import annotations

import torch
from torch import nn

from .gaussian import gaussian_blur2d

__all__ = ["unsharp_mask", "UnsharpMask"]


@torch.compiler.disable  # opaque to torch.compile: a clean graph break around the CUDA op
def unsharp_mask(input: torch.Tensor, kernel_size: tuple[int, int] | int, sigma: tuple[float, float] | torch.Tensor,
                 border_type: str = "reflect") -> torch.Tensor:
    """Sharpen: ``2 * input - gaussian_blur2d(input)``, evaluated as the reference's ``lerp(blur, input, 2)`` so
    the rounding matches.  Without a gradient the blend runs in the epilogue of the blur kernel (one pass over the
    image instead of two kernels and five full-size tensor streams); with one, blur and ``torch.lerp`` compose."""
    if not (torch.is_grad_enabled() and isinstance(input, torch.Tensor) and input.requires_grad
            or isinstance(sigma, torch.Tensor) and torch.is_grad_enabled() and sigma.requires_grad):
        fused = _fused_unsharp(input, kernel_size, sigma, border_type)
        if fused is not None:
            return fused
    blurred = gaussian_blur2d(input, kernel_size, sigma, border_type)
    return torch.lerp(blurred, input, weight=2.0)


def _fused_unsharp(input, kernel_size, sigma, border_type):
    """The blur's validation and taps, then kb200_sepfilter_lerp_forward; None when the request is outside the fused
    kernel's envelope (the caller composes)."""
    from .. import _lib, _ops
    from . import gaussian as G

    if not (isinstance(input, torch.Tensor) and input.is_cuda and input.dtype == torch.float32 and input.dim() == 4 and input.numel() > 0):
        return None
    ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    if len(ks) != 2 or ks[0] != ks[1] or not isinstance(ks[0], int) or ks[0] % 2 == 0 or not 3 <= ks[0] <= 11:
        return None
    code = _lib.BORDERS.get(str(border_type))
    if code is None or code == _lib.CIRCULAR:
        return None
    if isinstance(sigma, tuple):
        if len(sigma) != 2 or not all(float(v) > 0 for v in sigma):
            return None  # let gaussian_blur2d word the error
        kx, ky = G._constant_taps(kernel_size, tuple(float(v) for v in sigma), True, input.device, input.dtype)
    elif isinstance(sigma, torch.Tensor) and sigma.dim() == 2 and sigma.shape[-1] == 2:
        st = sigma.to(device=input.device, dtype=input.dtype)
        if not bool((st > 0).all()) or input.shape[0] % st.shape[0] != 0:
            return None
        kx, ky = G._taps(kernel_size, st, True)
    else:
        return None
    B, C, H, W = input.shape
    if ks[0] // 2 >= min(H, W):
        return None
    x = input.contiguous()
    kx, ky = kx.contiguous(), ky.contiguous()
    out = torch.empty_like(x)
    try:
        with torch.cuda.device(x.device), _ops._Timed("sepfilter_lerp_forward", x):
            _lib.call("kb200_sepfilter_lerp_forward", x.data_ptr(), kx.data_ptr(), ky.data_ptr(), out.data_ptr(), B, C, H, W, kx.shape[0],
                      kx.shape[1], ky.shape[0], ky.shape[1], code, 1, 2.0, _lib.F32, _ops._stream(x))
    except _lib.Unsupported:
        return None
    _ops._bump()
    return out


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
