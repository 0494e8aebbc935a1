"""Drop-in ``undistort_image`` (reference: kornia/geometry/calibration/undistort.py:138-198):
every output pixel is pushed through the lens model (``distort_points``) to find where the distorted image holds
it, and the image is resampled there by ``remap`` -- the tiled TMA kernel of csrc/remap_tiled.cuh (bilinear,
zeros, align_corners=True).  By default the maps are (B,H,W) torch tensors as in the reference.  The one-kernel form
(kb200_undistort_forward: the lens model evaluated per pixel in registers, no maps, ~45 fewer elementwise passes) is
written and passes on the host emulator but has not run on hardware yet: opt-in with KB200_FUSED_UNDISTORT=1 (DESIGN.md
section 9)."""
from __future__ import annotations

import os

import torch

from ... import _lib, _ops
from ..transform.imgwarp import remap
from .distort import distort_points

__all__ = ["undistort_image"]


def _fused_request(image: torch.Tensor, K: torch.Tensor, dist: torch.Tensor, B: int):
    """(B,16) lens numbers fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6, s1..s4 when the one-kernel path
    (kb200_undistort_forward: the lens model evaluated per pixel inside the sampling kernel, no maps) may serve the call,
    else None.  Off unless KB200_FUSED_UNDISTORT=1: written after the round's GPU budget was spent, not yet run on
    hardware.  Not taken when a gradient is needed or the tilt coefficients are set."""
    if os.environ.get("KB200_FUSED_UNDISTORT") != "1":
        return None
    if not (image.is_cuda and image.dtype == torch.float32):
        return None
    if torch.is_grad_enabled() and (image.requires_grad or K.requires_grad or dist.requires_grad):
        return None
    lens = pack_lens(K, dist, image)
    return lens if lens is not None and lens.shape[0] == B else None


def pack_lens(K: torch.Tensor, dist: torch.Tensor, like: torch.Tensor):
    """(B,16) rows fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 -- the layout kb200_undistort_forward
    reads (csrc/remap_tiled.cuh:lens_distort) -- or None when the tilt coefficients are set."""
    d = dist.reshape(-1, dist.shape[-1]).to(like)
    if d.shape[-1] == 14 and bool((d[:, 12:] != 0).any()):
        return None
    if d.shape[-1] < 12:
        d = torch.nn.functional.pad(d, [0, 12 - d.shape[-1]])
    Kf = K.reshape(-1, 3, 3).to(like)
    n = max(Kf.shape[0], d.shape[0])
    intrinsics = torch.stack([Kf[:, 0, 0], Kf[:, 1, 1], Kf[:, 0, 2], Kf[:, 1, 2]], -1)
    return torch.cat([intrinsics.expand(n, 4), d[:, :12].expand(n, 12)], -1).contiguous()


@torch.compiler.disable  # opaque to torch.compile: a clean graph break around the CUDA ops
def undistort_image(image: torch.Tensor, K: torch.Tensor, dist: torch.Tensor) -> torch.Tensor:
    """Remove the lens distortion ``dist`` (*,4|5|8|12|14) of camera ``K`` (*,3,3) from ``image`` (*,C,H,W)."""
    if image.dim() < 3:
        raise ValueError(f"Image shape is invalid. Got: {image.shape}.")
    if K.shape[-2:] != (3, 3):
        raise ValueError(f"K matrix shape is invalid. Got {K.shape}.")
    if dist.shape[-1] not in (4, 5, 8, 12, 14):
        raise ValueError(f"Invalid number of distortion coefficients. Got {dist.shape[-1]}.")
    if not image.is_floating_point():
        raise ValueError(f"Invalid input image data type. Input should be float. Got {image.dtype}.")
    lead = image.shape[:-3]
    if lead != K.shape[:-2] or lead != dist.shape[:-1]:
        # (1,C,H,W) with an unbatched K (3,3) and dist (n,) is accepted (undistort.py:174-181)
        if not (lead == (1,) and K.shape[:-2] == () and dist.shape[:-1] == ()):
            raise ValueError("Input shape is invalid. Input batch dimensions should match. "
                             f"Got {image.shape[:-3]}, {K.shape[:-2]}, {dist.shape[:-1]}.")
    channels, rows, cols = image.shape[-3:]
    B = image.numel() // (channels * rows * cols)
    lens = _fused_request(image, K, dist, B)
    if lens is not None:
        try:
            return _ops.undistort_fused(image.reshape(B, channels, rows, cols), lens).view_as(image)
        except _lib.Unsupported:
            pass
    # pixel grid (x, y) of the output, (rows*cols, 2), in the image's dtype (grid.py:65-79 with normalized=False)
    xs = torch.linspace(0, cols - 1, cols, device=image.device, dtype=image.dtype)
    ys = torch.linspace(0, rows - 1, rows, device=image.device, dtype=image.dtype)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    pts = torch.stack([gx, gy], -1).reshape(-1, 2)
    seen_at = distort_points(pts, K, dist)
    map_x = seen_at[..., 0].reshape(B, rows, cols)
    map_y = seen_at[..., 1].reshape(B, rows, cols)
    out = remap(image.reshape(B, channels, rows, cols), map_x, map_y, align_corners=True)
    return out.view_as(image)
