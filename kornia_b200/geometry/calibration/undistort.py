"""Drop-in ``undistort_image`` (reference: kornia/geometry/calibration/undistort.py:138-198):
every output pixel is pushed through the lens model (``distort_points``) to find where the distorted image holds
it, and the image is resampled there by ``remap`` -- the tiled TMA kernel of csrc/remap_tiled.cuh (bilinear,
zeros, align_corners=True).  The maps are (B,H,W) torch tensors as in the reference; fusing their evaluation into
the sampling kernel (8 B/pixel less traffic and ~40 fewer launches) is the listed next step."""
from __future__ import annotations

import torch

from ..transform.imgwarp import remap
from .distort import distort_points

__all__ = ["undistort_image"]


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
