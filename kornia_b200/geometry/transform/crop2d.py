"""Drop-in ``crop_and_resize`` / ``center_crop`` / ``crop_by_boxes`` / ``crop_by_transform_mat``
(reference: kornia/geometry/transform/crop2d.py:41-122,125-206,209-296,299-402).  Host logic over
:func:`warp_affine` / :func:`warp_perspective`."""
from __future__ import annotations

from typing import Tuple

import torch

from .imgwarp import warp_affine, warp_perspective
from .matrices import get_perspective_transform

__all__ = ["crop_and_resize", "center_crop", "crop_by_boxes", "crop_by_transform_mat"]


def _need_bchw(t: torch.Tensor) -> None:
    if t.dim() != 4:
        raise AssertionError(f"Only torch.Tensor with shape (B, C, H, W) supported. Got {t.shape}.")


def _dst_corners(dst_h: int, dst_w: int, n: int, like: torch.Tensor) -> torch.Tensor:
    """(n,4,2) corners of a dst_h x dst_w patch: top-left, top-right, bottom-right, bottom-left (x,y)."""
    pts = torch.tensor([[[0, 0], [dst_w - 1, 0], [dst_w - 1, dst_h - 1], [0, dst_h - 1]]], device=like.device, dtype=like.dtype)
    return pts.expand(n, -1, -1)


def crop_and_resize(input_tensor: torch.Tensor, boxes: torch.Tensor, size: Tuple[int, int], mode: str = "bilinear",
                    padding_mode: str = "zeros", align_corners: bool = True) -> torch.Tensor:
    """Cut the quadrilaterals ``boxes`` (B,4,2; clockwise from top-left, x,y) out of ``input_tensor``
    (B,C,H,W) and resample each to ``size`` = (h, w)."""
    if not isinstance(input_tensor, torch.Tensor):
        raise TypeError(f"Input torch.tensor type is not a torch.Tensor. Got {type(input_tensor)}")
    if not isinstance(boxes, torch.Tensor):
        raise TypeError(f"Input boxes type is not a torch.Tensor. Got {type(boxes)}")
    if not isinstance(size, (tuple, list)) or len(size) != 2:
        raise ValueError(f"Input size must be a tuple/list of length 2. Got {size}")
    _need_bchw(input_tensor)
    src = boxes.to(input_tensor)
    return crop_by_boxes(input_tensor, src, _dst_corners(size[0], size[1], src.shape[0], input_tensor), mode, padding_mode,
                         align_corners)


def center_crop(input_tensor: torch.Tensor, size: Tuple[int, int], mode: str = "bilinear", padding_mode: str = "zeros",
                align_corners: bool = True) -> torch.Tensor:
    """Crop the central ``size`` = (h, w) window of every image of ``input_tensor`` (B,C,H,W)."""
    if not isinstance(input_tensor, torch.Tensor):
        raise TypeError(f"Input torch.tensor type is not a torch.Tensor. Got {type(input_tensor)}")
    if not isinstance(size, (tuple, list)) or len(size) != 2:
        raise ValueError(f"Input size must be a tuple/list of length 2. Got {size}")
    _need_bchw(input_tensor)
    dst_h, dst_w = size
    src_h, src_w = input_tensor.shape[-2:]
    x0, y0 = src_w / 2 - dst_w / 2, src_h / 2 - dst_h / 2
    x1, y1 = x0 + dst_w - 1, y0 + dst_h - 1
    src = torch.tensor([[[x0, y0], [x1, y0], [x1, y1], [x0, y1]]], device=input_tensor.device, dtype=input_tensor.dtype)
    return crop_by_boxes(input_tensor, src, _dst_corners(dst_h, dst_w, 1, input_tensor), mode, padding_mode, align_corners)


def crop_by_boxes(input_tensor: torch.Tensor, src_box: torch.Tensor, dst_box: torch.Tensor, mode: str = "bilinear",
                  padding_mode: str = "zeros", align_corners: bool = True, validate_boxes: bool = True) -> torch.Tensor:
    """Warp the ``src_box`` quadrilaterals onto the ``dst_box`` rectangles (both (B,4,2)); the
    output size is the (common) size of the destination boxes."""
    _need_bchw(input_tensor)
    dst_trans_src = get_perspective_transform(src_box.to(input_tensor), dst_box.to(input_tensor))
    widths = dst_box[:, 1, 0] - dst_box[:, 0, 0] + 1
    heights = dst_box[:, 2, 1] - dst_box[:, 0, 1] + 1
    if not ((heights == heights[0]).all() and (widths == widths[0]).all()):
        raise AssertionError(
            f"Cropping height, width and depth must be exact same in a batch. Got height {heights} and width {widths}.")
    return crop_by_transform_mat(input_tensor, dst_trans_src, (int(heights[0].item()), int(widths[0].item())), mode=mode,
                                 padding_mode=padding_mode, align_corners=align_corners)


def crop_by_transform_mat(input_tensor: torch.Tensor, transform: torch.Tensor, out_size: Tuple[int, int],
                          mode: str = "bilinear", padding_mode: str = "zeros", align_corners: bool = True) -> torch.Tensor:
    """Resample ``input_tensor`` (B,C,H,W) through ``transform`` ((B|1,2,3) affine or (B|1,3,3)
    projective, source->destination pixels) into ``out_size`` = (h, w) patches."""
    dst_trans_src = torch.as_tensor(transform.expand(input_tensor.shape[0], -1, -1), device=input_tensor.device,
                                    dtype=input_tensor.dtype)
    if transform.shape[-2:] == (2, 3):
        return warp_affine(input_tensor, dst_trans_src, out_size, mode=mode, padding_mode=padding_mode,
                           align_corners=align_corners)
    h_out, w_out = out_size
    if not align_corners and (h_out == 1 or w_out == 1):
        # the half-pixel reparametrisation below is singular for one-pixel outputs: affine sampling instead
        return warp_affine(input_tensor, dst_trans_src[:, :2, :], out_size, mode=mode, padding_mode=padding_mode,
                           align_corners=align_corners)
    if not align_corners:
        # warp_perspective always spaces its grid corner-aligned; map the half-pixel convention onto it
        fix = torch.tensor([[w_out / (w_out - 1.0), 0.0, -0.5], [0.0, h_out / (h_out - 1.0), -0.5], [0.0, 0.0, 1.0]],
                           device=dst_trans_src.device, dtype=dst_trans_src.dtype)
        dst_trans_src = fix.unsqueeze(0) @ dst_trans_src
    return warp_perspective(input_tensor, dst_trans_src, out_size, mode=mode, padding_mode=padding_mode,
                            align_corners=align_corners)
