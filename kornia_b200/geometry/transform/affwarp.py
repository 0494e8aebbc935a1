"""Drop-in ``affine`` / ``rotate`` / ``translate`` / ``scale`` / ``shear`` (reference:
kornia/geometry/transform/affwarp.py:52-134,136-193,257-325,401-573).  Thin callers of
:func:`warp_affine`: they build a (B,2,3) pixel matrix and warp onto the input's own size, so they
inherit the fused CUDA warp.

``resize`` / ``rescale`` / ``resize_to_be_divisible`` (affwarp.py:576-763) are different animals: the reference
resamples with ``torch.nn.functional.interpolate`` (not with the warp), and only their ``antialias=True`` pre-filter
is on this library's path -- a Gaussian blur sized from the shrink factor, which runs in the one-pass separable
kernel here.  The resampling call itself stays the reference's own ATen call."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from ...filters.gaussian import gaussian_blur2d
from .imgwarp import warp_affine
from .matrices import get_rotation_matrix2d

__all__ = ["affine", "rotate", "translate", "scale", "shear", "resize", "rescale", "resize_to_be_divisible", "Resize", "Rescale"]


def _tensor_center(tensor: torch.Tensor) -> torch.Tensor:
    """(x, y) of the middle of the last two axes: ((W-1)/2, (H-1)/2)."""
    if not 2 <= tensor.dim() <= 4:
        raise AssertionError(f"Must be a 3D tensor as HW, CHW and BCHW. Got {tensor.shape}.")
    height, width = tensor.shape[-2:]
    return torch.tensor([float(width - 1) / 2, float(height - 1) / 2], device=tensor.device, dtype=tensor.dtype)


def _eye3(like: torch.Tensor) -> torch.Tensor:
    return torch.eye(3, device=like.device, dtype=like.dtype)[None].repeat(like.shape[0], 1, 1)


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


def _size_from_side(side_size: int, aspect_ratio: float, side: str) -> Tuple[int, int]:
    """(h, w) with the named side equal to ``side_size`` and the other one following ``aspect_ratio`` = w / h
    (affwarp.py:576-585)."""
    if side not in ("short", "long", "vert", "horz"):
        raise ValueError(f"side can be one of 'short', 'long', 'vert', and 'horz'. Got '{side}'")
    fix_height = side == "vert" or (side != "horz" and ((side == "short") != (aspect_ratio < 1.0)))
    if fix_height:
        return side_size, int(side_size * aspect_ratio)
    return int(side_size / aspect_ratio), side_size


def _odd_at_least_3(sigma: float) -> int:
    k = int(max(2.0 * 2 * sigma, 3))
    return k if k % 2 else k + 1


def resize(input: torch.Tensor, size: Union[int, Tuple[int, int]], interpolation: str = "bilinear",
           align_corners: Optional[bool] = None, side: str = "short", antialias: bool = False) -> torch.Tensor:
    """Resample the last two axes of ``input`` ((H,W), (C,H,W), (B,C,H,W) or (*,C,H,W)) onto ``size`` = (h, w), or
    onto the size that makes ``side`` equal to an int ``size`` at the same aspect ratio.  With ``antialias`` a
    shrinking resize first blurs with sigma = (factor - 1) / 2 per axis (affwarp.py:659-664)."""
    if not isinstance(input, torch.Tensor):
        raise TypeError(f"Input tensor type is not a torch.Tensor. Got {type(input)}")
    if input.dim() < 2:
        raise ValueError(f"Input tensor must have at least two dimensions. Got {input.dim()}")
    shape = input.shape
    h, w = shape[-2:]
    if isinstance(size, int):
        size = _size_from_side(size, w / h, side)
    if len(shape) == 2:
        x = input[None, None]
    elif len(shape) == 3:
        x = input[None]
    elif len(shape) > 4:
        x = input.reshape(-1, *shape[-3:])
    else:
        x = input
    shrink = (h / size[0], w / size[1])
    if antialias and max(shrink) > 1:
        sigmas = (max((shrink[0] - 1.0) / 2.0, 0.001), max((shrink[1] - 1.0) / 2.0, 0.001))
        x = gaussian_blur2d(x, (_odd_at_least_3(sigmas[0]), _odd_at_least_3(sigmas[1])), sigmas)
    out = torch.nn.functional.interpolate(x, size=size, mode=interpolation, align_corners=align_corners)
    if len(shape) == 2:
        return out[0, 0]
    if len(shape) == 3:
        return out[0]
    if len(shape) > 4:
        return out.reshape(*shape[:-2], size[0], size[1])
    return out


def resize_to_be_divisible(input: torch.Tensor, divisible_factor: int, interpolation: str = "bilinear",
                           align_corners: Optional[bool] = None, side: str = "short", antialias: bool = False) -> torch.Tensor:
    """:func:`resize` onto the nearest multiples of ``divisible_factor`` (affwarp.py:679-715; 3-D / 4-D inputs)."""
    height, width = input.shape[-2:]
    height = round(height / divisible_factor) * divisible_factor
    width = round(width / divisible_factor) * divisible_factor
    return resize(input, (height, width), interpolation, align_corners, side, antialias)


def rescale(input: torch.Tensor, factor: Union[float, Tuple[float, float]], interpolation: str = "bilinear",
            align_corners: Optional[bool] = None, antialias: bool = False) -> torch.Tensor:
    """:func:`resize` onto (int(H * fv), int(W * fh)); ``factor`` is a float or a (vertical, horizontal) pair
    (affwarp.py:718-763)."""
    fv, fh = (factor, factor) if isinstance(factor, float) else factor
    height, width = input.shape[-2:]
    return resize(input, (int(height * fv), int(width * fh)), interpolation=interpolation, align_corners=align_corners,
                  antialias=antialias)


class Resize(torch.nn.Module):
    """Module form of :func:`resize` (affwarp.py:766-837)."""

    def __init__(self, size: Union[int, Tuple[int, int]], interpolation: str = "bilinear", align_corners: Optional[bool] = None,
                 side: str = "short", antialias: bool = False) -> None:
        super().__init__()
        self.size, self.interpolation, self.align_corners, self.side, self.antialias = size, interpolation, align_corners, side, antialias

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return resize(input, self.size, self.interpolation, align_corners=self.align_corners, side=self.side, antialias=self.antialias)


class Rescale(torch.nn.Module):
    """Module form of :func:`rescale` (affwarp.py:957-1012)."""

    def __init__(self, factor: Union[float, Tuple[float, float]], interpolation: str = "bilinear", align_corners: Optional[bool] = None,
                 antialias: bool = False) -> None:
        super().__init__()
        self.factor, self.interpolation, self.align_corners, self.antialias = factor, interpolation, align_corners, antialias

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return rescale(input, self.factor, self.interpolation, align_corners=self.align_corners, antialias=self.antialias)
