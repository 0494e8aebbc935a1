"""``distort_points`` / ``tilt_projection`` (reference: kornia/geometry/calibration/distort.py:25-75,78-189): the
OpenCV lens model -- rational radial (k1..k6), tangential (p1,p2), thin-prism (s1..s4) and sensor tilt (tau_x,tau_y)
-- applied to pixel coordinates.  Pure torch elementwise ops on (B,N) coordinates, kept as the reference's op
sequence (one rounding per op, differentiable w.r.t. points, K and dist); what it feeds -- the per-pixel maps of
``undistort_image`` -- is consumed by the CUDA ``remap`` kernel."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

__all__ = ["distort_points", "tilt_projection"]


def tilt_projection(taux: torch.Tensor, tauy: torch.Tensor, return_inverse: bool = False) -> torch.Tensor:
    """(*,3,3) projection of a sensor tilted by ``taux`` / ``tauy`` radians about x / y, or its inverse."""
    if taux.shape != tauy.shape:
        raise ValueError(f"Shape of taux {taux.shape} and tauy {tauy.shape} do not match.")
    scalar = taux.dim() == 0
    taux, tauy = taux.reshape(-1), tauy.reshape(-1)
    cx, sx, cy, sy = torch.cos(taux), torch.sin(taux), torch.cos(tauy), torch.sin(tauy)
    o, l = torch.zeros_like(cx), torch.ones_like(cx)
    rot_x = torch.stack([l, o, o, o, cx, sx, o, -sx, cx], -1).reshape(-1, 3, 3)
    rot_y = torch.stack([cy, o, -sy, o, l, o, sy, o, cy], -1).reshape(-1, 3, 3)
    rot = rot_y @ rot_x
    if return_inverse:
        inv22 = 1 / rot[..., 2, 2]
        unproject = torch.stack([inv22, o, rot[..., 0, 2] * inv22, o, inv22, rot[..., 1, 2] * inv22, o, o, l], -1).reshape(-1, 3, 3)
        res = rot.transpose(-1, -2) @ unproject
    else:
        project = torch.stack([rot[..., 2, 2], o, -rot[..., 0, 2], o, rot[..., 2, 2], -rot[..., 1, 2], o, o, l], -1).reshape(-1, 3, 3)
        res = project @ rot.transpose(-1, -2)
    return torch.squeeze(res) if scalar else res


def _pad_to_14(dist: torch.Tensor) -> torch.Tensor:
    if dist.shape[-1] not in (4, 5, 8, 12, 14):
        raise ValueError(f"Invalid number of distortion coefficients. Got {dist.shape[-1]}")
    return F.pad(dist, [0, 14 - dist.shape[-1]]) if dist.shape[-1] < 14 else dist


def distort_points(points: torch.Tensor, K: torch.Tensor, dist: torch.Tensor, new_K: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pixel coordinates ``points`` (*,N,2) of an ideal camera ``new_K`` -> where the distorted camera ``K`` with
    coefficients ``dist`` (*,4|5|8|12|14) = (k1,k2,p1,p2[,k3[,k4,k5,k6[,s1,s2,s3,s4[,tx,ty]]]]) sees them."""
    if points.dim() < 2 and points.shape[-1] != 2:
        raise ValueError(f"points shape is invalid. Got {points.shape}.")
    if K.shape[-2:] != (3, 3):
        raise ValueError(f"K matrix shape is invalid. Got {K.shape}.")
    if new_K is None:
        new_K = K
    elif new_K.shape[-2:] != (3, 3):
        raise ValueError(f"new_K matrix shape is invalid. Got {new_K.shape}.")
    d = _pad_to_14(dist)

    def coef(i):
        return d[..., i:i + 1]

    # pixels -> normalised camera coordinates of the ideal camera
    x = (points[..., 0] - new_K[..., 0:1, 2]) / new_K[..., 0:1, 0]
    y = (points[..., 1] - new_K[..., 1:2, 2]) / new_K[..., 1:2, 1]
    r2 = x * x + y * y
    r4 = r2 * r2
    r6 = r4 * r2
    radial = (1 + coef(0) * r2 + coef(1) * r4 + coef(4) * r6) / (1 + coef(5) * r2 + coef(6) * r4 + coef(7) * r6)
    xd = x * radial + 2 * coef(2) * x * y + coef(3) * (r2 + 2 * x * x) + coef(8) * r2 + coef(9) * r4
    yd = y * radial + coef(2) * (r2 + 2 * y * y) + 2 * coef(3) * x * y + coef(10) * r2 + coef(11) * r4
    if torch.any(d[..., 12] != 0) or torch.any(d[..., 13] != 0):
        tilt = tilt_projection(d[..., 12], d[..., 13])
        tilted = torch.stack([xd, yd, torch.ones_like(xd)], -1) @ tilt.transpose(-2, -1)
        xd = tilted[..., 0] / tilted[..., 2]
        yd = tilted[..., 1] / tilted[..., 2]
    # normalised -> pixels of the distorted camera
    return torch.stack([K[..., 0:1, 0] * xd + K[..., 0:1, 2], K[..., 1:2, 1] * yd + K[..., 1:2, 2]], -1)
