"""Drop-in ``affine`` / ``rotate`` / ``translate`` / ``scale`` / ``shear`` (reference:
kornia/geometry/transform/affwarp.py:52-134,136-193,257-325,401-573).  Thin callers of
:func:`warp_affine`: they build a (B,2,3) pixel matrix and warp onto the input's own size, so they
inherit the fused CUDA warp."""
from __future__ import annotations

from typing import Optional

import torch

from .imgwarp import warp_affine
from .matrices import get_rotation_matrix2d

__all__ = ["affine", "rotate", "translate", "scale", "shear"]


def _tensor_center(tensor: torch.Tensor) -> torch.Tensor:
    """(x, y) of the middle of the last two axes: ((W-1)/2, (H-1)/2)."""
    if not 2 <= tensor.dim() <= 4:
        raise AssertionError(f"Must be a 3D tensor as HW, CHW and BCHW. Got {tensor.shape}.")
    height, width = tensor.shape[-2:]
    return torch.tensor([float(width - 1) / 2, float(height - 1) / 2], device=tensor.device, dtype=tensor.dtype)


def _eye3(like: torch.Tensor) -> torch.Tensor:
    return torch.eye(3, device=like.device, dtype=like.dtype)[None].repeat(like.shape[0], 1, 1)


@torch.compiler.disable  # opaque to torch.compile: a clean graph break around the CUDA ops
def affine(tensor: torch.Tensor, matrix: torch.Tensor, mode: str = "bilinear", padding_mode: str = "zeros",
           align_corners: bool = True) -> torch.Tensor:
    """Warp ``tensor`` ((C,H,W) or (B,C,H,W)) with the source->destination pixel matrices ``matrix``
    (B,2,3) onto its own size; a single image is broadcast over a batch of matrices."""
    unbatched = tensor.dim() == 3
    if unbatched:
        tensor = tensor.unsqueeze(0)
    if tensor.shape[0] == 1 and matrix.shape[0] != 1:
        tensor = tensor.expand(matrix.shape[0], -1, -1, -1)
    matrix = matrix.expand(tensor.shape[0], -1, -1)
    out = warp_affine(tensor, matrix, (tensor.shape[-2], tensor.shape[-1]), mode, padding_mode, align_corners)
    return out.squeeze(0) if unbatched else out


def _check_image(tensor, other, other_name: str) -> None:
    if not isinstance(tensor, torch.Tensor):
        raise TypeError(f"Input tensor type is not a torch.Tensor. Got {type(tensor)}")
    if not isinstance(other, torch.Tensor):
        raise TypeError(f"Input {other_name} type is not a torch.Tensor. Got {type(other)}")


@torch.compiler.disable  # opaque to torch.compile: a clean graph break around the CUDA ops
def rotate(tensor: torch.Tensor, angle: torch.Tensor, center: Optional[torch.Tensor] = None, mode: str = "bilinear",
           padding_mode: str = "zeros", align_corners: bool = True) -> torch.Tensor:
    """Rotate counter-clockwise (as displayed) by ``angle`` (B,) degrees about ``center`` (B,2; x,y),
    by default the image centre."""
    _check_image(tensor, angle, "angle")
    if center is not None and not isinstance(center, torch.Tensor):
        raise TypeError(f"Input center type is not a torch.Tensor. Got {type(center)}")
    if tensor.dim() not in (3, 4):
        raise ValueError(f"Invalid tensor shape, we expect CxHxW or BxCxHxW. Got: {tensor.shape}")
    if center is None:
        center = _tensor_center(tensor)
    angle = angle.expand(tensor.shape[0])
    center = center.expand(tensor.shape[0], -1)
    matrix = get_rotation_matrix2d(center, angle, torch.ones_like(center))
    return affine(tensor, matrix[..., :2, :3], mode, padding_mode, align_corners)


@torch.compiler.disable  # opaque to torch.compile: a clean graph break around the CUDA ops
def translate(tensor: torch.Tensor, translation: torch.Tensor, mode: str = "bilinear", padding_mode: str = "zeros",
              align_corners: bool = True) -> torch.Tensor:
    """Shift by ``translation`` (B,2) = (dx, dy) pixels."""
    _check_image(tensor, translation, "translation")
    if tensor.dim() not in (3, 4):
        raise ValueError(f"Invalid tensor shape, we expect CxHxW or BxCxHxW. Got: {tensor.shape}")
    matrix = _eye3(translation)
    dx, dy = torch.chunk(translation, chunks=2, dim=-1)
    matrix[..., 0, 2:3] += dx
    matrix[..., 1, 2:3] += dy
    return affine(tensor, matrix[..., :2, :3], mode, padding_mode, align_corners)


@torch.compiler.disable  # opaque to torch.compile: a clean graph break around the CUDA ops
def scale(tensor: torch.Tensor, scale_factor: torch.Tensor, center: Optional[torch.Tensor] = None, mode: str = "bilinear",
          padding_mode: str = "zeros", align_corners: bool = True) -> torch.Tensor:
    """Zoom by ``scale_factor`` ((B,) isotropic or (B,2) = (sx, sy)) about ``center`` (default: image centre)."""
    _check_image(tensor, scale_factor, "scale_factor")
    if scale_factor.dim() == 1:
        scale_factor = scale_factor.repeat(1, 2)
    if center is None:
        center = _tensor_center(tensor)
    center = center.expand(tensor.shape[0], -1)
    scale_factor = scale_factor.expand(tensor.shape[0], 2)
    no_turn = torch.zeros(scale_factor.shape[:1], device=scale_factor.device, dtype=scale_factor.dtype)
    matrix = get_rotation_matrix2d(center, no_turn, scale_factor)
    return affine(tensor, matrix[..., :2, :3], mode, padding_mode, align_corners)


@torch.compiler.disable  # opaque to torch.compile: a clean graph break around the CUDA ops
def shear(tensor: torch.Tensor, shear: torch.Tensor, mode: str = "bilinear", padding_mode: str = "zeros",
          align_corners: bool = False) -> torch.Tensor:
    """Skew by ``shear`` (B,2) = (shx, shy)."""
    _check_image(tensor, shear, "shear")
    if tensor.dim() not in (3, 4):
        raise ValueError(f"Invalid tensor shape, we expect CxHxW or BxCxHxW. Got: {tensor.shape}")
    matrix = _eye3(shear)
    shx, shy = torch.chunk(shear, chunks=2, dim=-1)
    matrix[..., 0, 1:2] += shx
    matrix[..., 1, 0:1] += shy
    return affine(tensor, matrix[..., :2, :3], mode, padding_mode, align_corners)
