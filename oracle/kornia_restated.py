"""ORACLE (test infrastructure, NOT product code).

Restatement, in plain torch ops, of the Kornia 0.9.0rc1 composition for the warp /
filter hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; ``kornia_b200`` never does.

Why torch ops: the reference has no native code.  Its arithmetic is a sequence of ATen
calls (``F.grid_sample``, ``F.pad`` + ``F.conv2d``, elementwise ops); restating the
*sequence* with the same ATen calls reproduces the reference bit for bit on the same
device (pinned by ``tests/test_oracle_vs_golden.py`` against vectors produced by the
imported reference, see ``tests/golden/make_golden.py``).  The ATen sampler/convolution
themselves (third-party, torch>=2.0, lock-pinned 2.9.1, ``uv.lock:2525``) are restated
independently in ``oracle/aten_restated.py`` (numpy) so the chain is checked end to end.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# B x 3 x 3 prelude
# --------------------------------------------------------------------------------------
def pixel_to_norm_matrix(height: int, width: int, like: torch.Tensor) -> torch.Tensor:
    """kornia/geometry/conversions.py:1729-1765 (normal_transform_pixel): python-float
    scale 2/(size-1) (1e-14 denominator when size == 1), offsets -1; returned as (1,3,3)."""
    wd = 1e-14 if width == 1 else width - 1.0
    hd = 1e-14 if height == 1 else height - 1.0
    mat = torch.tensor([[2.0 / wd, 0.0, -1.0], [0.0, 2.0 / hd, -1.0], [0.0, 0.0, 1.0]])
    return mat.unsqueeze(0).to(like)


def inv3x3(a: torch.Tensor) -> torch.Tensor:
    """kornia/core/utils.py:159-166: adjugate rows are cross products of the columns,
    divided by det = col_a . (col_b x col_c)."""
    ca, cb, cc = a[..., :, 0], a[..., :, 1], a[..., :, 2]
    r0 = torch.linalg.cross(cb, cc, dim=-1)
    r1 = torch.linalg.cross(cc, ca, dim=-1)
    r2 = torch.linalg.cross(ca, cb, dim=-1)
    det = (ca * r0).sum(-1)
    return torch.stack([r0, r1, r2], dim=-2) / det[..., None, None]


def normalized_homography(M: torch.Tensor, src_hw, dst_hw) -> torch.Tensor:
    """kornia/geometry/conversions.py:1717-1725: N_dst @ (M @ inv(N_src))."""
    n_src = pixel_to_norm_matrix(src_hw[0], src_hw[1], M)
    n_dst = pixel_to_norm_matrix(dst_hw[0], dst_hw[1], M)
    return n_dst @ (M @ inv3x3(n_src))


def affine_to_3x3(A: torch.Tensor) -> torch.Tensor:
    """kornia/geometry/conversions.py:342-345,372-378."""
    if not isinstance(A, torch.Tensor):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(A)}")
    if not (A.dim() == 3 and tuple(A.shape[-2:]) == (2, 3)):
        raise ValueError(f"Input matrix must be a Bx2x3 tensor. Got {A.shape}")
    Hm = F.pad(A, [0, 0, 0, 1], "constant", value=0.0)
    Hm[..., -1, -1] += 1.0
    return Hm


# --------------------------------------------------------------------------------------
# base grids
# --------------------------------------------------------------------------------------
def meshgrid_axes(h: int, w: int, device) -> tuple[torch.Tensor, torch.Tensor]:
    """kornia/geometry/grid.py:65-78: fp32 linspace, (x/(w-1) - 0.5)*2 (always corner aligned)."""
    xs = torch.linspace(0, w - 1, w, device=device)
    ys = torch.linspace(0, h - 1, h, device=device)
    xs = (xs / (w - 1) - 0.5) * 2
    ys = (ys / (h - 1) - 0.5) * 2
    return xs, ys


def affine_axes(h: int, w: int, align_corners: bool, device, dtype):
    """kornia/geometry/transform/imgwarp.py:271-276."""
    if align_corners:
        xs = torch.linspace(-1.0, 1.0, w, device=device, dtype=dtype)
        ys = torch.linspace(-1.0, 1.0, h, device=device, dtype=dtype)
    else:
        xs = torch.linspace(-1.0 + 1.0 / w, 1.0 - 1.0 / w, w, device=device, dtype=dtype)
        ys = torch.linspace(-1.0 + 1.0 / h, 1.0 - 1.0 / h, h, device=device, dtype=dtype)
    return xs, ys


# --------------------------------------------------------------------------------------
# warps
# --------------------------------------------------------------------------------------
def perspective_grid(m: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor) -> torch.Tensor:
    """kornia/geometry/transform/imgwarp.py:165-170 (eager branch)."""
    gx0 = xs[None, None, :]
    gy0 = ys[None, :, None]
    den = m[:, 2, 0, None, None] * gx0 + m[:, 2, 1, None, None] * gy0 + m[:, 2, 2, None, None]
    gx = (m[:, 0, 0, None, None] * gx0 + m[:, 0, 1, None, None] * gy0 + m[:, 0, 2, None, None]) / den
    gy = (m[:, 1, 0, None, None] * gx0 + m[:, 1, 1, None, None] * gy0 + m[:, 1, 2, None, None]) / den
    return torch.stack([gx, gy], dim=-1)


def affine_grid(m: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor) -> torch.Tensor:
    """kornia/geometry/transform/imgwarp.py:277-281."""
    by, bx = torch.meshgrid(ys, xs, indexing="ij")
    gx = m[:, 0, 0, None, None] * bx + m[:, 0, 1, None, None] * by + m[:, 0, 2, None, None]
    gy = m[:, 1, 0, None, None] * bx + m[:, 1, 1, None, None] * by + m[:, 1, 2, None, None]
    return torch.stack([gx, gy], dim=-1)


def fill_and_sample(src, grid, mode, align_corners, fill_value):
    """kornia/geometry/transform/imgwarp.py:308-320: sample + (1 - sample(ones)) * fill."""
    ones = torch.ones_like(src)
    fill_value = fill_value.to(ones)
    if fill_value.ndim == 0:
        fill_value = fill_value.view(1, 1, 1, 1)
    elif fill_value.ndim == 1:
        fill_value = fill_value.view(1, -1, 1, 1)
    inv_cover = 1 - F.grid_sample(ones, grid, align_corners=align_corners, mode=mode, padding_mode="zeros")
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode="zeros") + inv_cover * fill_value


def warp_perspective(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True, fill_value=None):
    """kornia/geometry/transform/imgwarp.py:124-174."""
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
    H, W = src.shape[-2:]
    h, w = dsize
    m = inv3x3(normalized_homography(M, (H, W), (h, w)))
    xs, ys = meshgrid_axes(h, w, src.device)
    grid = perspective_grid(m, xs.to(src.dtype), ys.to(src.dtype))
    if padding_mode == "fill":
        return fill_and_sample(src, grid, mode, align_corners, fill_value)
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode=padding_mode)


def warp_affine(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True, fill_value=None):
    """kornia/geometry/transform/imgwarp.py:234-290."""
    if not isinstance(src, torch.Tensor):
        raise TypeError(f"Input src type is not a torch.Tensor. Got {type(src)}")
    if not isinstance(M, torch.Tensor):
        raise TypeError(f"Input M type is not a torch.Tensor. Got {type(M)}")
    if src.dim() != 4:
        raise ValueError(f"Input src must be a BxCxHxW torch.Tensor. Got {src.shape}")
    if not (M.dim() == 3 or tuple(M.shape[-2:]) == (2, 3)):
        raise ValueError(f"Input M must be a Bx2x3 torch.Tensor. Got {M.shape}")
    B, C, H, W = src.shape
    m = inv3x3(normalized_homography(affine_to_3x3(M), (H, W), dsize))
    xs, ys = affine_axes(dsize[0], dsize[1], align_corners, src.device, src.dtype)
    grid = affine_grid(m, xs, ys)
    if M.shape[0] == 1 and B > 1:
        grid = grid.expand(B, -1, -1, -1)
    if padding_mode == "fill":
        if fill_value is None:
            fill_value = torch.zeros(C, device=src.device, dtype=src.dtype)
        return fill_and_sample(src, grid, mode, align_corners, fill_value)
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode=padding_mode)


def normalize_pixel_coords(xy: torch.Tensor, height: int, width: int, eps: float = 1e-8) -> torch.Tensor:
    """kornia/geometry/conversions.py:1484-1498: (2 / clamp(size-1, eps)) * p - 1, x first."""
    hw = torch.stack([
        torch.tensor(width, device=xy.device, dtype=xy.dtype),
        torch.tensor(height, device=xy.device, dtype=xy.dtype),
    ])
    factor = torch.tensor(2.0, device=xy.device, dtype=xy.dtype) / (hw - 1).clamp(eps)
    return factor * xy - 1


def remap(image, map_x, map_y, mode="bilinear", padding_mode="zeros", align_corners: Optional[bool] = None,
          normalized_coordinates=False):
    """kornia/geometry/transform/imgwarp.py:681-702."""
    B, _, H, W = image.shape
    xy = torch.stack([map_x, map_y], -1)
    if not normalized_coordinates:
        xy = normalize_pixel_coords(xy, H, W)
    xy = xy.expand(B, -1, -1, -1)
    if align_corners is None:
        align_corners = False
    return F.grid_sample(image, xy, mode=mode, padding_mode=padding_mode, align_corners=align_corners)


# --------------------------------------------------------------------------------------
# filters
# --------------------------------------------------------------------------------------
def same_padding(kh: int, kw: int) -> list[int]:
    """kornia/filters/filter.py:31-51: front=(k-1)//2, rear=(k-1)-front; last dim first."""
    out = []
    for k in (kw, kh):
        front = (k - 1) // 2
        out += [front, (k - 1) - front]
    return out


def filter2d(input, kernel, border_type="reflect", normalized=False, padding="same", behaviour="corr"):
    """kornia/filters/filter.py:121-152."""
    b, c, h, w = input.shape
    k = kernel.flip((-2, -1)) if str(behaviour).lower() == "conv" else kernel
    k = k[:, None, ...].to(device=input.device, dtype=input.dtype)
    if normalized:
        k = k / k.abs().sum(dim=-1).sum(dim=-1)[..., None, None]  # kernels.py:72-74
    k = k.expand(-1, c, -1, -1)
    kh, kw = k.shape[-2:]
    if padding == "same":
        input = F.pad(input, same_padding(kh, kw), mode=border_type)
    k = k.reshape(-1, 1, kh, kw)
    input = input.view(-1, k.size(0), input.size(-2), input.size(-1))
    out = F.conv2d(input, k, groups=k.size(0), padding=0, stride=1)
    if padding == "same":
        return out.view(b, c, h, w)
    return out.view(b, c, h - kh + 1, w - kw + 1)


def filter2d_separable(input, kernel_x, kernel_y, border_type="reflect", normalized=False, padding="same"):
    """kornia/filters/filter.py:205-207: x pass (1 x kw) then y pass (kh x 1)."""
    out_x = filter2d(input, kernel_x[..., None, :], border_type, normalized, padding)
    return filter2d(out_x, kernel_y[..., None], border_type, normalized, padding)


def gaussian_taps(window_size: int, sigma: torch.Tensor) -> torch.Tensor:
    """kornia/filters/kernels.py:104-120 with the default mean window_size // 2; sigma is (B,1)."""
    bs = sigma.shape[0]
    mean = torch.tensor([[float(window_size // 2)]], device=sigma.device, dtype=sigma.dtype)
    x = (torch.arange(window_size, device=sigma.device, dtype=sigma.dtype) - mean).expand(bs, -1)
    if window_size % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma.pow(2.0)))
    return g / g.sum(-1, keepdim=True)


def gaussian_kernel2d(kernel_size, sigma: torch.Tensor) -> torch.Tensor:
    """kornia/filters/kernels.py:705-715: outer product ky[...,None] * kx.view(-1,1,kx)."""
    ky, kx = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    ty = gaussian_taps(int(ky), sigma[:, 0, None])[..., None]
    tx = gaussian_taps(int(kx), sigma[:, 1, None])[..., None]
    return ty * tx.view(-1, 1, int(kx))


def gaussian_blur2d(input, kernel_size, sigma, border_type="reflect", separable=True):
    """kornia/filters/gaussian.py:95-118 (validation omitted: the oracle is fed valid inputs)."""
    if isinstance(sigma, tuple):
        sigma = torch.tensor([sigma], device=input.device, dtype=input.dtype)
    else:
        sigma = sigma.to(device=input.device, dtype=input.dtype)
    ky, kx = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    if separable:
        bs = sigma.shape[0]
        kernel_x = gaussian_taps(int(kx), sigma[:, 1].view(bs, 1))
        kernel_y = gaussian_taps(int(ky), sigma[:, 0].view(bs, 1))
        return filter2d_separable(input, kernel_x, kernel_y, border_type)
    return filter2d(input, gaussian_kernel2d((ky, kx), sigma), border_type)
