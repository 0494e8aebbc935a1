"""Drop-in ``warp_perspective`` / ``warp_affine`` / ``remap``.

Same names, signatures, defaults, validation order and exception types as
kornia/geometry/transform/imgwarp.py:69,177,625.  What changes is the execution: the reference
builds a base grid, runs ~15 broadcast elementwise kernels, stacks a (B,h,w,2) grid and calls
``F.grid_sample``; here the tiny (B,3,3) prelude stays in torch (same ops, so the matrices are
bit-identical and differentiable) and ONE fused CUDA kernel does map + divide + gather.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _prelude as P
from ... import _lib
from ... import _ops
from ...core.check import check_shape

__all__ = ["warp_perspective", "warp_affine", "remap"]


def _mode_codes(mode: str, padding_mode: str, allow_fill: bool):
    if mode not in _lib.INTERP:
        # same text F.grid_sample produces for an unknown mode
        raise ValueError(f"nn.functional.grid_sample(): expected mode to be 'bilinear', 'nearest' or 'bicubic', but got: '{mode}'")
    if padding_mode not in _lib.PADDING or (padding_mode == "fill" and not allow_fill):
        raise ValueError("nn.functional.grid_sample(): expected padding_mode to be 'zeros', 'border', or 'reflection', "
                         f"but got: '{padding_mode}'")
    return _lib.INTERP[mode], _lib.PADDING[padding_mode]


def warp_perspective(
    src: torch.Tensor,
    M: torch.Tensor,
    dsize: tuple[int, int],
    mode: str = "bilinear",
    padding_mode: str = "zeros",
    align_corners: bool = True,
    fill_value: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Resample ``src`` (B,C,H,W) through the source->destination pixel homographies ``M`` (B,3,3)
    into an image of size ``dsize = (h, w)``.

    ``mode``: 'bilinear' | 'nearest' | 'bicubic'; ``padding_mode``: 'zeros' | 'border' |
    'reflection' | 'fill' (``fill_value`` of shape (3,), RGB only).  ``align_corners`` only affects
    the sampler: the homography is always normalised corner-aligned, as in the reference
    (imgwarp.py:147,157).
    """
    if not isinstance(src, torch.Tensor):
        raise TypeError(f"Input src type is not a torch.Tensor. Got {type(src)}")
    if not isinstance(M, torch.Tensor):
        raise TypeError(f"Input M type is not a torch.Tensor. Got {type(M)}")
    if src.dim() != 4:
        raise ValueError(f"Input src must be a BxCxHxW torch.Tensor. Got {src.shape}")
    if not (M.dim() == 3 and tuple(M.shape[-2:]) == (3, 3)):
        raise ValueError(f"Input M must be a Bx3x3 torch.Tensor. Got {M.shape}")
    if fill_value is None:
        fill_value = torch.zeros(3)
    if padding_mode == "fill" and fill_value.shape != torch.Size([3]):
        raise ValueError(f"Padding_tensor only supported for 3 channels. Got {fill_value.shape}")
    interp, pad = _mode_codes(mode, padding_mode, allow_fill=True)

    B, C, H, W = src.shape
    h_out, w_out = int(dsize[0]), int(dsize[1])
    if M.shape[0] != B:
        # grid_sample's own complaint in the reference (the grid inherits M's batch)
        raise RuntimeError(f"grid_sampler(): expected grid and input to have same batch size, but got input with sizes "
                           f"{list(src.shape)} and grid with sizes {[M.shape[0], h_out, w_out, 2]}")
    # (B,3,3) prelude, identical op sequence to imgwarp.py:147-153
    m = P.sampling_matrix(M, (H, W), (h_out, w_out), affine=False)
    bx, by = P.meshgrid_axes(h_out, w_out, src.device, src.dtype)
    fill = None
    if pad == _lib.FILL:
        if C != 3:
            # the reference's (1,3,1,1) fill cannot broadcast against C != 3 channels
            raise RuntimeError(f"The size of tensor a ({C}) must match the size of tensor b (3) at non-singleton dimension 1")
        fill = fill_value
    return _ops.warp(src, m, bx, by, fill, h_out, w_out, True, interp, pad, bool(align_corners))


def warp_affine(
    src: torch.Tensor,
    M: torch.Tensor,
    dsize: tuple[int, int],
    mode: str = "bilinear",
    padding_mode: str = "zeros",
    align_corners: bool = True,
    fill_value: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Resample ``src`` (B,C,H,W) through the source->destination pixel affine maps ``M``
    ((B,2,3), or (1,2,3) shared by the whole batch) into ``dsize = (h, w)``.

    The base grid honours ``align_corners`` (imgwarp.py:271-276) while the matrix normalisation
    stays corner-aligned (:250) -- reference behaviour, reproduced.  ``fill_value``: (C,), (1,)
    or 0-d.
    """
    if not isinstance(src, torch.Tensor):
        raise TypeError(f"Input src type is not a torch.Tensor. Got {type(src)}")
    if not isinstance(M, torch.Tensor):
        raise TypeError(f"Input M type is not a torch.Tensor. Got {type(M)}")
    if src.dim() != 4:
        raise ValueError(f"Input src must be a BxCxHxW torch.Tensor. Got {src.shape}")
    if not (M.dim() == 3 or tuple(M.shape[-2:]) == (2, 3)):
        raise ValueError(f"Input M must be a Bx2x3 torch.Tensor. Got {M.shape}")
    interp, pad = _mode_codes(mode, padding_mode, allow_fill=True)

    B, C, H, W = src.shape
    h_out, w_out = int(dsize[0]), int(dsize[1])
    if not (M.dim() == 3 and tuple(M.shape[-2:]) == (2, 3)):  # conversions.py:375-376
        raise ValueError(f"Input matrix must be a Bx2x3 tensor. Got {M.shape}")
    m = P.sampling_matrix(M, (H, W), (h_out, w_out), affine=True)
    B_M = M.shape[0]
    if B_M != B and not (B_M == 1 and B > 1):
        raise RuntimeError(f"grid_sampler(): expected grid and input to have same batch size, but got input with sizes "
                           f"{list(src.shape)} and grid with sizes {[B_M, h_out, w_out, 2]}")
    bx, by = P.affine_axes(h_out, w_out, bool(align_corners), src.device, src.dtype)
    fill = None
    if pad == _lib.FILL:
        fill = torch.zeros(C, device=src.device, dtype=src.dtype) if fill_value is None else fill_value
        fill = fill.to(device=src.device, dtype=src.dtype)
        if fill.ndim == 0 or (fill.ndim == 1 and fill.numel() == 1):
            fill = fill.reshape(1).expand(C)
        elif fill.ndim != 1 or fill.numel() != C:
            raise RuntimeError(f"The size of tensor a ({C}) must match the size of tensor b ({fill.shape[-1] if fill.ndim else 1}) "
                               "at non-singleton dimension 1")
    return _ops.warp(src, m, bx, by, fill, h_out, w_out, False, interp, pad, bool(align_corners))


def remap(
    image: torch.Tensor,
    map_x: torch.Tensor,
    map_y: torch.Tensor,
    mode: str = "bilinear",
    padding_mode: str = "zeros",
    align_corners: Optional[bool] = None,
    normalized_coordinates: bool = False,
) -> torch.Tensor:
    """``out[b,:,y,x] = image[b,:, map_y[b,y,x], map_x[b,y,x]]`` with interpolation.

    Maps are (B,h,w) or (1,h,w) pixel coordinates (or already in [-1,1] when
    ``normalized_coordinates``).  As in the reference, pixel maps are normalised corner-aligned
    (conversions.py:1487-1498) while ``align_corners=None`` resolves to ``False`` for the sampler
    (imgwarp.py:698-699).
    """
    check_shape(image, ["B", "C", "H", "W"])
    check_shape(map_x, ["B", "H", "W"])
    check_shape(map_y, ["B", "H", "W"])
    interp, pad = _mode_codes(mode, padding_mode, allow_fill=False)
    B = image.shape[0]
    if map_x.shape != map_y.shape:
        raise RuntimeError(f"stack expects each tensor to be equal size, but got {list(map_x.shape)} at entry 0 and "
                           f"{list(map_y.shape)} at entry 1")
    if map_x.shape[0] not in (1, B):
        raise RuntimeError(f"The expanded size of the tensor ({B}) must match the existing size ({map_x.shape[0]}) at "
                           "non-singleton dimension 0")
    if align_corners is None:
        align_corners = False
    return _ops.remap(image, map_x, map_y, bool(normalized_coordinates), interp, pad, bool(align_corners))
