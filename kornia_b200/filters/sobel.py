"""Drop-in ``spatial_gradient`` / ``sobel`` and their modules (reference: kornia/filters/sobel.py:
32-74,134-167,170-235,286-340).

The reference pads a replicate-border copy and runs ``F.conv2d`` with a (2|3,1,k,k) weight; ``sobel``
then slices the result and runs five more elementwise passes.  Here one CUDA kernel reads the image
once and writes either all derivative planes or directly the gradient magnitude.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _ops
from ..core.check import check_is_tensor, check_shape
from .kernels import get_spatial_gradient_kernel2d, normalize_kernel2d

__all__ = ["spatial_gradient", "sobel", "SpatialGradient", "Sobel"]

_TAPS: dict = {}


@torch.compiler.assume_constant_result
def _host_taps(mode: str, order: int, normalized: bool, dtype: torch.dtype):
    """Derivative taps as python floats, built (and normalised) with the reference's torch ops in
    ``dtype`` so that they carry exactly the values the reference convolves with.  Constants of the call: evaluated with
    real tensors even while a tracer (dynamo, torch.export) runs the caller under fake tensors."""
    key = (mode, order, bool(normalized), dtype)
    hit = _TAPS.get(key)
    if hit is None:
        from torch._subclasses.fake_tensor import unset_fake_temporarily

        with unset_fake_temporarily(), torch.inference_mode(False), torch.no_grad():
            kernel = get_spatial_gradient_kernel2d(mode, order, dtype=dtype)
            if normalized:
                kernel = normalize_kernel2d(kernel)
            hit = _TAPS[key] = (tuple(kernel.double().flatten().tolist()), kernel.shape[0], kernel.shape[-1])
    return hit


def spatial_gradient(input: torch.Tensor, mode: str = "sobel", order: int = 1, normalized: bool = True) -> torch.Tensor:
    """Image derivatives of every channel of ``input`` (B,C,H,W) over a replicate border:
    (B,C,2,H,W) = (d/dx, d/dy) for ``order=1``, (B,C,3,H,W) = (dxx, dxy, dyy) for ``order=2``;
    ``mode`` 'sobel' or 'diff'; ``normalized`` scales each stencil to unit L1 norm."""
    check_is_tensor(input)
    check_shape(input, ["B", "C", "H", "W"])
    taps, nout, k = _host_taps(mode, order, normalized, input.dtype)
    return _ops.spatial_gradient(input, taps, nout, k, False, 0.0)


def sobel(input: torch.Tensor, normalized: bool = True, eps: float = 1e-6) -> torch.Tensor:
    """Sobel edge magnitude ``sqrt(gx^2 + gy^2 + eps)`` of every channel of ``input`` (B,C,H,W)."""
    check_is_tensor(input)
    check_shape(input, ["B", "C", "H", "W"])
    taps, nout, k = _host_taps("sobel", 1, normalized, input.dtype)
    if torch.is_grad_enabled() and input.requires_grad:
        edges = _ops.spatial_gradient(input, taps, nout, k, False, 0.0)
        gx, gy = edges[:, :, 0], edges[:, :, 1]
        return torch.sqrt(gx * gx + gy * gy + eps)
    return _ops.spatial_gradient(input, taps, nout, k, True, eps)


class SpatialGradient(nn.Module):
    """Module form of :func:`spatial_gradient` (same constructor as the reference's)."""

    def __init__(self, mode: str = "sobel", order: int = 1, normalized: bool = True) -> None:
        super().__init__()
        self.normalized = normalized
        self.order = order
        self.mode = mode

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(order={self.order}, normalized={self.normalized}, mode={self.mode})"

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return spatial_gradient(input, self.mode, self.order, self.normalized)


class Sobel(nn.Module):
    """Module form of :func:`sobel` (same constructor as the reference's)."""

    def __init__(self, normalized: bool = True, eps: float = 1e-6) -> None:
        super().__init__()
        self.normalized = normalized
        self.eps = eps

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(normalized={self.normalized})"

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return sobel(input, self.normalized, self.eps)
