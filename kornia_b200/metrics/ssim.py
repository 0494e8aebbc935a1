"""Drop-in ``ssim`` and ``SSIM`` (reference: kornia/metrics/ssim.py:34-139,142-214).

The reference blurs img1, img2, img1^2, img2^2 and img1*img2 with five ``filter2d_separable`` calls and
combines them with fourteen elementwise kernels.  Inference-style calls (fp32, odd window <= 11, no
gradient) run as ONE CUDA kernel that reads the two images once and writes the index map; every other
call composes the same map from this library's ``filter2d_separable`` (differentiable)."""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib, _ops
from ..filters.filter import _compute_padding, filter2d_separable
from ..filters.kernels import get_gaussian_kernel1d

__all__ = ["ssim", "SSIM"]


def _composed(img1, img2, kernel, C1: float, C2: float, eps: float, crop):
    """ssim.py:103-139 on top of the one-pass separable filter; ``crop`` is None or the 'valid' margins."""

    def blur(t):
        out = filter2d_separable(t, kernel, kernel)
        return out if crop is None else torch.nn.functional.pad(out, crop)

    mu1, mu2 = blur(img1), blur(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    sigma1_sq = blur(img1 ** 2) - mu1_sq
    sigma2_sq = blur(img2 ** 2) - mu2_sq
    sigma12 = blur(img1 * img2) - mu1_mu2
    num = (2.0 * mu1_mu2 + C1) * (2.0 * sigma12 + C2)
    den = (mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2)
    return num / (den + eps)


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int, max_val: float = 1.0, eps: float = 1e-12,
         padding: str = "same") -> torch.Tensor:
    """Structural-similarity index map (B,C,H,W) of two image batches: Gaussian window of
    ``window_size`` taps (sigma 1.5), dynamic range ``max_val``; ``padding='valid'`` crops the
    half-window margin (the MATLAB convention)."""
    if not isinstance(img1, torch.Tensor):
        raise TypeError(f"Input img1 type is not a torch.Tensor. Got {type(img1)}")
    if not isinstance(img2, torch.Tensor):
        raise TypeError(f"Input img2 type is not a torch.Tensor. Got {type(img2)}")
    if not isinstance(max_val, float):
        raise TypeError(f"Input max_val type is not a float. Got {type(max_val)}")
    if img1.dim() != 4:
        raise ValueError(f"Invalid img1 shape, we expect BxCxHxW. Got: {img1.shape}")
    if img2.dim() != 4:
        raise ValueError(f"Invalid img2 shape, we expect BxCxHxW. Got: {img2.shape}")
    if img1.shape != img2.shape:
        raise ValueError(f"img1 and img2 shapes must be the same. Got: {img1.shape} and {img2.shape}")

    kernel = get_gaussian_kernel1d(window_size, 1.5, device=img1.device, dtype=img1.dtype)
    C1, C2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    crop = None
    if padding == "valid":
        m = _compute_padding([kernel.shape[-1], kernel.shape[-1]])
        crop = (-m[2], -m[3], -m[0], -m[1])

    if not torch.compiler.is_compiling():
        _ops._require_cuda(img1, "img1")
        _ops._require_cuda(img2, "img2")
    B, C, H, W = img1.shape
    K = kernel.shape[-1]
    needs_grad = torch.is_grad_enabled() and (img1.requires_grad or img2.requires_grad)
    fused = (img1.dtype == torch.float32 and img2.dtype == torch.float32 and not needs_grad and K % 2 == 1 and 3 <= K <= 11
             and K // 2 < min(H, W) and img1.numel() > 0)
    if fused:
        try:
            out = _ops.ops.ssim_fwd(img1, img2, kernel, float(C1), float(C2), float(eps))
            return out if crop is None else torch.nn.functional.pad(out, crop)
        except _lib.Unsupported:
            pass
    return _composed(img1, img2, kernel, C1, C2, eps, crop)


class SSIM(nn.Module):
    """Module form of :func:`ssim` (same constructor as the reference's)."""

    def __init__(self, window_size: int, max_val: float = 1.0, eps: float = 1e-12, padding: str = "same") -> None:
        super().__init__()
        self.window_size = window_size
        self.max_val = max_val
        self.eps = eps
        self.padding = padding

    def forward(self, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
        return ssim(img1, img2, self.window_size, self.max_val, self.eps, self.padding)
