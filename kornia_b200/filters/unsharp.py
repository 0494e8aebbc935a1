"""Drop-in ``unsharp_mask`` and ``UnsharpMask`` (reference: kornia/filters/unsharp.py:27-54,57-95)."""
from __future__ import annotations

import torch
from torch import nn

from .gaussian import gaussian_blur2d

__all__ = ["unsharp_mask", "UnsharpMask"]


def unsharp_mask(input: torch.Tensor, kernel_size: tuple[int, int] | int, sigma: tuple[float, float] | torch.Tensor,
                 border_type: str = "reflect") -> torch.Tensor:
    """Sharpen: ``2 * input - gaussian_blur2d(input)``, evaluated as the reference's ``lerp(blur, input, 2)`` so
    the rounding matches.  Without a gradient the blend runs in the epilogue of the blur kernel (one pass over the
    image instead of two kernels and five full-size tensor streams); with one, blur and ``torch.lerp`` compose."""
    needs_grad = torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in (input, sigma))
    if not needs_grad:
        fused = _fused_unsharp(input, kernel_size, sigma, border_type)
        if fused is not None:
            return fused
    blurred = gaussian_blur2d(input, kernel_size, sigma, border_type)
    return torch.lerp(blurred, input, weight=2.0)


def _fused_request(input, kernel_size, sigma, border_type):
    """(kernel_x, kernel_y, border code) when the request lies inside the fused kernel's envelope -- fp32 (B,C,H,W),
    square odd kernel of 3..11 taps, non-circular border, positive sigmas -- else None (the caller composes, and
    gaussian_blur2d words any error).  Pure host logic."""
    from .. import _lib
    from .gaussian import _constant_taps, _taps

    if not (isinstance(input, torch.Tensor) and input.dtype == torch.float32 and input.dim() == 4 and input.numel() > 0):
        return None
    ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    if len(ks) != 2 or ks[0] != ks[1] or not isinstance(ks[0], int) or ks[0] % 2 == 0 or not 3 <= ks[0] <= 11:
        return None
    if ks[0] // 2 >= min(input.shape[-2:]):
        return None
    code = _lib.BORDERS.get(str(border_type))
    if code is None or code == _lib.CIRCULAR:
        return None
    if isinstance(sigma, tuple):
        if len(sigma) != 2 or not all(float(v) > 0 for v in sigma):
            return None
        kx, ky = _constant_taps(kernel_size, tuple(float(v) for v in sigma), True, input.device, input.dtype)
    elif isinstance(sigma, torch.Tensor) and sigma.dim() == 2 and sigma.shape[-1] == 2:
        st = sigma.to(device=input.device, dtype=input.dtype)
        if input.shape[0] % st.shape[0] != 0 or not bool((st > 0).all()):  # the sync the reference pays too (gaussian.py:107)
            return None
        kx, ky = _taps(kernel_size, st, True)
    else:
        return None
    return kx.contiguous(), ky.contiguous(), code


def _fused_unsharp(input, kernel_size, sigma, border_type):
    """kb200_sepfilter_lerp_forward on the blur's own taps; None when the request is outside the fused envelope."""
    from .. import _lib, _ops

    if not (isinstance(input, torch.Tensor) and input.is_cuda):
        return None
    req = _fused_request(input, kernel_size, sigma, border_type)
    if req is None:
        return None
    kx, ky, code = req
    try:
        return _ops.ops.sepfilter_lerp_fwd(input, kx, ky, code, 2.0)
    except _lib.Unsupported:
        return None


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
