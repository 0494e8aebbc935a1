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


# --------------------------------------------------------------------------------------
# filter family on the same loaders
# --------------------------------------------------------------------------------------
def box_blur(input, kernel_size, border_type="reflect", separable=False):
    """kornia/filters/blur.py:62-76 with the box taps of kernels.py:314-315,331-333 (1/k, 1/(ky*kx))."""
    ky, kx = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    if separable:
        tx = torch.full((1, int(kx)), 1.0 / int(kx), device=input.device, dtype=input.dtype)
        ty = torch.full((1, int(ky)), 1.0 / int(ky), device=input.device, dtype=input.dtype)
        return filter2d_separable(input, tx, ty, border_type)
    taps = torch.full((1, int(ky), int(kx)), 1.0 / (int(kx) * int(ky)), device=input.device, dtype=input.dtype)
    return filter2d(input, taps, border_type)


def laplacian(input, kernel_size, border_type="reflect", normalized=True):
    """kornia/filters/laplacian.py:59-63 with kernels.py:831-838 (ones, centre = 1 - sum)."""
    ky, kx = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    taps = torch.ones((int(ky), int(kx)), device=input.device, dtype=input.dtype)
    taps[int(ky) // 2, int(kx) // 2] = 1 - taps.sum()
    return filter2d(input, taps[None], border_type, normalized=normalized)


def unsharp_mask(input, kernel_size, sigma, border_type="reflect"):
    """kornia/filters/unsharp.py:53-54."""
    return torch.lerp(gaussian_blur2d(input, kernel_size, sigma, border_type), input, weight=2.0)


_D1 = {"sobel": [[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]],
       "diff": [[0.0, 0.0, 0.0], [-1.0, 0.0, 1.0], [0.0, 0.0, 0.0]]}
_D2 = {"sobel": ([[-1.0, 0.0, 2.0, 0.0, -1.0], [-4.0, 0.0, 8.0, 0.0, -4.0], [-6.0, 0.0, 12.0, 0.0, -6.0],
                  [-4.0, 0.0, 8.0, 0.0, -4.0], [-1.0, 0.0, 2.0, 0.0, -1.0]],
                 [[-1.0, -2.0, 0.0, 2.0, 1.0], [-2.0, -4.0, 0.0, 4.0, 2.0], [0.0, 0.0, 0.0, 0.0, 0.0],
                  [2.0, 4.0, 0.0, -4.0, -2.0], [1.0, 2.0, 0.0, -2.0, -1.0]]),
       "diff": ([[0.0, 0.0, 0.0], [1.0, -2.0, 1.0], [0.0, 0.0, 0.0]], [[-1.0, 0.0, 1.0], [0.0, 0.0, 0.0], [1.0, 0.0, -1.0]])}


def derivative_taps(mode: str, order: int, device, dtype) -> torch.Tensor:
    """kornia/filters/kernels.py:357-398,470-528: (2,3,3) [d/dx, d/dy] or (3,k,k) [dxx, dxy, dyy]."""
    if order == 1:
        kx = torch.tensor(_D1[mode], device=device, dtype=dtype)
        return torch.stack([kx, kx.t()])
    xx, xy = (torch.tensor(t, device=device, dtype=dtype) for t in _D2[mode])
    return torch.stack([xx, xy, xx.t()])


def spatial_gradient(input, mode="sobel", order=1, normalized=True):
    """kornia/filters/sobel.py:59-74: replicate pad of k//2, conv2d with the (nout,1,k,k) weight."""
    taps = derivative_taps(mode, order, input.device, input.dtype)
    if normalized:
        taps = taps / taps.abs().sum(dim=-1).sum(dim=-1)[..., None, None]
    b, c, h, w = input.shape
    half_h, half_w = taps.size(1) // 2, taps.size(2) // 2
    padded = F.pad(input.reshape(b * c, 1, h, w), [half_h, half_h, half_w, half_w], "replicate")
    out = F.conv2d(padded, taps[:, None], padding=0, stride=1)
    return out.reshape(b, c, taps.shape[0], h, w)


def sobel(input, normalized=True, eps=1e-6):
    """kornia/filters/sobel.py:158-167."""
    edges = spatial_gradient(input, normalized=normalized)
    gx, gy = edges[:, :, 0], edges[:, :, 1]
    return torch.sqrt(gx * gx + gy * gy + eps)


# --------------------------------------------------------------------------------------
# callers of the warps: matrix builders + affine / crop wrappers
# --------------------------------------------------------------------------------------
def rotation_matrix2d(center, angle_deg, scale):
    """kornia/geometry/transform/imgwarp.py:607-622 (T(c) @ R @ S @ T(-c)) with
    conversions.py:148 (deg2rad through the fp32 pi) and :1685-1688 ([[c, s], [-s, c]])."""
    n = center.shape[0]
    eye = torch.eye(3, device=center.device, dtype=center.dtype)[None].repeat(n, 1, 1)
    t_fwd, t_back, s_m, r_m = eye.clone(), eye.clone(), eye.clone(), eye.clone()
    t_fwd[:, :2, 2] = center
    t_back[:, :2, 2] = -center
    s_m[:, 0, 0] *= scale[:, 0]
    s_m[:, 1, 1] *= scale[:, 1]
    rad = angle_deg * torch.tensor(3.14159265358979323846).to(angle_deg.device).type(angle_deg.dtype) / 180.0
    c, s = torch.cos(rad), torch.sin(rad)
    r_m[:, :2, :2] = torch.stack([c, s, -s, c], dim=-1).view(n, 2, 2)
    return (t_fwd @ r_m @ s_m @ t_back)[:, :2, :]


def affine(tensor, matrix, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/affwarp.py:172-193."""
    single = tensor.dim() == 3
    if single:
        tensor = tensor[None]
    if tensor.shape[0] == 1 and matrix.shape[0] != 1:
        tensor = tensor.expand(matrix.shape[0], -1, -1, -1)
    matrix = matrix.expand(tensor.shape[0], -1, -1)
    out = warp_affine(tensor, matrix, tuple(tensor.shape[-2:]), mode, padding_mode, align_corners)
    return out[0] if single else out


def _center_xy(tensor):
    h, w = tensor.shape[-2:]
    return torch.tensor([float(w - 1) / 2, float(h - 1) / 2], device=tensor.device, dtype=tensor.dtype)


def rotate(tensor, angle, center=None, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/affwarp.py:312-325."""
    center = _center_xy(tensor) if center is None else center
    angle = angle.expand(tensor.shape[0])
    center = center.expand(tensor.shape[0], -1)
    return affine(tensor, rotation_matrix2d(center, angle, torch.ones_like(center)), mode, padding_mode, align_corners)


def translate(tensor, translation, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/affwarp.py:105-112,446-452."""
    m = torch.eye(3, device=translation.device, dtype=translation.dtype)[None].repeat(translation.shape[0], 1, 1)
    m[..., 0, 2:3] += translation[..., 0:1]
    m[..., 1, 2:3] += translation[..., 1:2]
    return affine(tensor, m[..., :2, :3], mode, padding_mode, align_corners)


def scale(tensor, scale_factor, center=None, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/affwarp.py:115-119,504-519."""
    if scale_factor.dim() == 1:
        scale_factor = scale_factor.repeat(1, 2)
    center = _center_xy(tensor) if center is None else center
    center = center.expand(tensor.shape[0], -1)
    scale_factor = scale_factor.expand(tensor.shape[0], 2)
    zero = torch.zeros(scale_factor.shape[:1], device=scale_factor.device, dtype=scale_factor.dtype)
    return affine(tensor, rotation_matrix2d(center, zero, scale_factor), mode, padding_mode, align_corners)


def shear(tensor, shear, mode="bilinear", padding_mode="zeros", align_corners=False):
    """kornia/geometry/transform/affwarp.py:122-133,566-573."""
    m = torch.eye(3, device=shear.device, dtype=shear.dtype)[None].repeat(shear.shape[0], 1, 1)
    m[..., 0, 1:2] += shear[..., 0:1]
    m[..., 1, 0:1] += shear[..., 1:2]
    return affine(tensor, m[..., :2, :3], mode, padding_mode, align_corners)


def square_to_quad(pts):
    """kornia/geometry/transform/imgwarp.py:411-441: unit square -> quadrilateral, Heckbert's closed form."""
    (x0, y0), (x1, y1), (x2, y2), (x3, y3) = [(pts[..., i, 0], pts[..., i, 1]) for i in range(4)]
    dx1, dx2, sx = x1 - x2, x3 - x2, x0 - x1 + x2 - x3
    dy1, dy2, sy = y1 - y2, y3 - y2, y0 - y1 + y2 - y3
    den = dx1 * dy2 - dy1 * dx2
    a31 = (sx * dy2 - sy * dx2) / den
    a32 = (dx1 * sy - dy1 * sx) / den
    r0 = torch.stack([x1 - x0 + a31 * x1, x3 - x0 + a32 * x3, x0], -1)
    r1 = torch.stack([y1 - y0 + a31 * y1, y3 - y0 + a32 * y3, y0], -1)
    r2 = torch.stack([a31, a32, torch.ones_like(x0)], -1)
    return torch.stack([r0, r1, r2], -2)


def perspective_from_points(points_src, points_dst):
    """kornia/geometry/transform/imgwarp.py:456-462: Q(dst) @ inv3x3(Q(src)), scaled to H[2,2] = 1."""
    h = square_to_quad(points_dst) @ inv3x3(square_to_quad(points_src))
    return h / h[..., 2:3, 2:3]


def crop_by_transform_mat(input_tensor, transform, out_size, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/crop2d.py:340-402."""
    t = transform.expand(input_tensor.shape[0], -1, -1).to(input_tensor)
    if transform.shape[-2:] == (2, 3):
        return warp_affine(input_tensor, t, out_size, mode, padding_mode, align_corners)
    h_out, w_out = out_size
    if not align_corners and (h_out == 1 or w_out == 1):
        return warp_affine(input_tensor, t[:, :2, :], out_size, mode, padding_mode, align_corners)
    if not align_corners:
        fix = torch.tensor([[w_out / (w_out - 1.0), 0.0, -0.5], [0.0, h_out / (h_out - 1.0), -0.5], [0.0, 0.0, 1.0]]).to(t)
        t = fix[None] @ t
    return warp_perspective(input_tensor, t, out_size, mode, padding_mode, align_corners)


def crop_by_boxes(input_tensor, src_box, dst_box, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/crop2d.py:277-296 (output size from the destination box, bbox.py infer_bbox_shape)."""
    t = perspective_from_points(src_box.to(input_tensor), dst_box.to(input_tensor))
    w_out = int((dst_box[0, 1, 0] - dst_box[0, 0, 0] + 1).item())
    h_out = int((dst_box[0, 2, 1] - dst_box[0, 0, 1] + 1).item())
    return crop_by_transform_mat(input_tensor, t, (h_out, w_out), mode, padding_mode, align_corners)


def _patch_corners(h, w, n, like):
    return torch.tensor([[[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]]], device=like.device, dtype=like.dtype).expand(n, -1, -1)


def crop_and_resize(input_tensor, boxes, size, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/crop2d.py:107-122."""
    src = boxes.to(input_tensor)
    return crop_by_boxes(input_tensor, src, _patch_corners(size[0], size[1], src.shape[0], input_tensor), mode, padding_mode,
                         align_corners)


def center_crop(input_tensor, size, mode="bilinear", padding_mode="zeros", align_corners=True):
    """kornia/geometry/transform/crop2d.py:173-206."""
    dst_h, dst_w = size
    src_h, src_w = input_tensor.shape[-2:]
    sx, sy = src_w / 2 - dst_w / 2, src_h / 2 - dst_h / 2
    ex, ey = sx + dst_w - 1, sy + dst_h - 1
    src = torch.tensor([[[sx, sy], [ex, sy], [ex, ey], [sx, ey]]], device=input_tensor.device, dtype=input_tensor.dtype)
    return crop_by_boxes(input_tensor, src, _patch_corners(dst_h, dst_w, 1, input_tensor), mode, padding_mode, align_corners)


def get_rotation_matrix2d(center, angle, scale):
    return rotation_matrix2d(center, angle, scale)


def get_perspective_transform(points_src, points_dst):
    return perspective_from_points(points_src, points_dst)


# --------------------------------------------------------------------------------------
# SSIM (five separable blurs + the index)
# --------------------------------------------------------------------------------------
def ssim(img1, img2, window_size, max_val=1.0, eps=1e-12, padding="same"):
    """kornia/metrics/ssim.py:92-139: Gaussian window (sigma 1.5), blurs of x, y, x^2, y^2, x*y over a
    reflect border, 'valid' = crop of the half-window margin."""
    sigma = torch.tensor([[1.5]], device=img1.device, dtype=img1.dtype)
    taps = gaussian_taps(window_size, sigma)
    C1, C2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    k = taps.shape[-1]
    front, rear = (k - 1) // 2, (k - 1) - (k - 1) // 2

    def blur(t):
        out = filter2d_separable(t, taps, taps)
        return out[..., front:out.shape[-2] - rear, front:out.shape[-1] - rear] if padding == "valid" else out

    mu1, mu2 = blur(img1), blur(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    sigma1_sq = blur(img1 ** 2) - mu1_sq
    sigma2_sq = blur(img2 ** 2) - mu2_sq
    sigma12 = blur(img1 * img2) - mu1_mu2
    num = (2.0 * mu1_mu2 + C1) * (2.0 * sigma12 + C2)
    den = (mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2)
    return num / (den + eps)


def ssim_loss(img1, img2, window_size, max_val=1.0, eps=1e-12, reduction="mean", padding="same"):
    """kornia/losses/ssim.py:67-82."""
    loss = torch.clamp((1.0 - ssim(img1, img2, window_size, max_val, eps, padding)) / 2, min=0, max=1)
    return loss.mean() if reduction == "mean" else loss.sum() if reduction == "sum" else loss


# --------------------------------------------------------------------------------------
# Pyramids, resize family, lens model
# --------------------------------------------------------------------------------------
def pyramid_taps() -> torch.Tensor:
    """kornia/geometry/transform/pyramid.py:32-47: the 5x5 literal, which is outer([1,4,6,4,1]) / 256."""
    b = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0])
    return torch.outer(b, b).unsqueeze(0) / 256.0


def pyrdown(input, border_type="reflect", align_corners=False, factor=2.0):
    """pyramid.py:444-457: filter2d with the pyramid taps, then F.interpolate(bilinear) onto
    (int(H / factor), int(W // factor))."""
    hh, ww = input.shape[2], input.shape[3]
    low = filter2d(input, pyramid_taps(), border_type)
    return F.interpolate(low, size=(int(float(hh) / factor), int(float(ww) // factor)), mode="bilinear", align_corners=align_corners)


def pyrup(input, border_type="reflect", align_corners=False):
    """pyramid.py:489-502: F.interpolate(bilinear) onto (2H, 2W), then filter2d with the pyramid taps."""
    big = F.interpolate(input, size=(2 * input.shape[2], 2 * input.shape[3]), mode="bilinear", align_corners=align_corners)
    return filter2d(big, pyramid_taps(), border_type)


def build_pyramid(input, max_level, border_type="reflect", align_corners=False):
    """pyramid.py:551-560."""
    out = [input]
    while len(out) < max_level:
        out.append(pyrdown(out[-1], border_type, align_corners))
    return out


def build_laplacian_pyramid(input, max_level, border_type="reflect", align_corners=False):
    """pyramid.py:632-665: reflect-pad right/bottom to the next powers of two unless H or W is one already; level i is
    gaussian[i] - pyrup(gaussian[i+1]); the coarsest gaussian level closes the list."""
    h, w = input.shape[2], input.shape[3]
    pow2 = lambda v: v != 0 and (v & (v - 1)) == 0  # noqa: E731
    if not pow2(h) and not pow2(w):
        input = F.pad(input, (0, (1 << (w - 1).bit_length()) - w, 0, (1 << (h - 1).bit_length()) - h), "reflect")
    g = build_pyramid(input, max_level, border_type, align_corners)
    return [a - pyrup(b, border_type, align_corners) for a, b in zip(g[:-1], g[1:])] + [g[-1]]


def resize(input, size, interpolation="bilinear", align_corners=None, side="short", antialias=False):
    """kornia/geometry/transform/affwarp.py:576-585,631-676: int size -> (h, w) by aspect ratio and side; leading
    dims folded into a batch; optional Gaussian pre-blur with sigma = max((factor - 1) / 2, 0.001) and an odd window
    of about 4 sigma (>= 3) when shrinking; F.interpolate."""
    lead, (h, w) = input.shape[:-2], input.shape[-2:]
    if isinstance(size, int):
        ar = w / h
        if side == "vert" or (side in ("short", "long") and ((side == "short") != (ar < 1.0))):
            size = (size, int(size * ar))
        else:
            size = (int(size / ar), size)
    x = input.reshape((-1,) + tuple(input.shape[-3:])) if input.dim() >= 4 else input.reshape((1,) * (4 - input.dim()) + tuple(input.shape))
    fy, fx = h / size[0], w / size[1]
    if antialias and max(fy, fx) > 1:
        sig = (max((fy - 1.0) / 2.0, 0.001), max((fx - 1.0) / 2.0, 0.001))
        win = [int(max(4.0 * s, 3)) for s in sig]
        win = tuple(k + 1 - (k % 2) for k in win)
        x = gaussian_blur2d(x, win, sig)
    y = F.interpolate(x, size=tuple(size), mode=interpolation, align_corners=align_corners)
    return y.reshape(tuple(lead) + (size[0], size[1]))


def rescale(input, factor, interpolation="bilinear", align_corners=None, antialias=False):
    """affwarp.py:756-763."""
    fv, fh = (factor, factor) if isinstance(factor, float) else factor
    return resize(input, (int(input.shape[-2] * fv), int(input.shape[-1] * fh)), interpolation, align_corners, "short", antialias)


def resize_to_be_divisible(input, divisible_factor, interpolation="bilinear", align_corners=None, side="short", antialias=False):
    """affwarp.py:707-715."""
    hh = round(input.shape[-2] / divisible_factor) * divisible_factor
    ww = round(input.shape[-1] / divisible_factor) * divisible_factor
    return resize(input, (hh, ww), interpolation, align_corners, side, antialias)


def tilt_projection(taux, tauy, return_inverse=False):
    """kornia/geometry/calibration/distort.py:25-75.  The 'inverse' form is the reference's R^T @ Pz^-1 (distort.py:58-66),
    used with row-vector points in undistort_points -- not the matrix inverse of the forward form."""
    if not return_inverse:
        return tilt_matrix(taux, tauy)
    tx, ty = taux.reshape(-1), tauy.reshape(-1)
    z, u = torch.zeros_like(tx), torch.ones_like(tx)
    Rx = torch.stack([u, z, z, z, tx.cos(), tx.sin(), z, -tx.sin(), tx.cos()], -1).reshape(-1, 3, 3)
    Ry = torch.stack([ty.cos(), z, -ty.sin(), z, u, z, ty.sin(), z, ty.cos()], -1).reshape(-1, 3, 3)
    R = Ry @ Rx
    i22 = 1 / R[..., 2, 2]
    Pinv = torch.stack([i22, z, R[..., 0, 2] * i22, z, i22, R[..., 1, 2] * i22, z, z, u], -1).reshape(-1, 3, 3)
    return R.transpose(-1, -2) @ Pinv


def tilt_matrix(taux, tauy):
    """kornia/geometry/calibration/distort.py:38-75 (forward form): Pz @ R^T with R = Ry @ Rx."""
    tx, ty = taux.reshape(-1), tauy.reshape(-1)
    z, u = torch.zeros_like(tx), torch.ones_like(tx)
    Rx = torch.stack([u, z, z, z, tx.cos(), tx.sin(), z, -tx.sin(), tx.cos()], -1).reshape(-1, 3, 3)
    Ry = torch.stack([ty.cos(), z, -ty.sin(), z, u, z, ty.sin(), z, ty.cos()], -1).reshape(-1, 3, 3)
    R = Ry @ Rx
    Pz = torch.stack([R[..., 2, 2], z, -R[..., 0, 2], z, R[..., 2, 2], -R[..., 1, 2], z, z, u], -1).reshape(-1, 3, 3)
    return Pz @ R.transpose(-1, -2)


def distort_points(points, K, dist, new_K=None):
    """distort.py:117-189: (points - c') / f' -> rational radial factor, tangential and thin-prism terms, optional
    tilt -> f * (.) + c."""
    Kn = K if new_K is None else new_K
    if dist.shape[-1] < 14:
        dist = F.pad(dist, [0, 14 - dist.shape[-1]])
    k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4 = (dist[..., i:i + 1] for i in range(12))
    x = (points[..., 0] - Kn[..., 0:1, 2]) / Kn[..., 0:1, 0]
    y = (points[..., 1] - Kn[..., 1:2, 2]) / Kn[..., 1:2, 1]
    r2 = x * x + y * y
    r4 = r2 * r2
    r6 = r4 * r2
    ratio = (1 + k1 * r2 + k2 * r4 + k3 * r6) / (1 + k4 * r2 + k5 * r4 + k6 * r6)
    xd = x * ratio + 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + s1 * r2 + s2 * r4
    yd = y * ratio + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + s3 * r2 + s4 * r4
    if bool((dist[..., 12] != 0).any()) or bool((dist[..., 13] != 0).any()):
        hom = torch.stack([xd, yd, torch.ones_like(xd)], -1) @ tilt_matrix(dist[..., 12], dist[..., 13]).transpose(-2, -1)
        xd, yd = hom[..., 0] / hom[..., 2], hom[..., 1] / hom[..., 2]
    return torch.stack([K[..., 0:1, 0] * xd + K[..., 0:1, 2], K[..., 1:2, 1] * yd + K[..., 1:2, 2]], -1)


def undistort_image(image, K, dist):
    """kornia/geometry/calibration/undistort.py:183-198: pixel grid (grid.py:65-79, unnormalised) -> distort_points
    -> remap(align_corners=True)."""
    C, H, W = image.shape[-3:]
    n = image.numel() // (C * H * W)
    xs = torch.linspace(0, W - 1, W, device=image.device, dtype=image.dtype)
    ys = torch.linspace(0, H - 1, H, device=image.device, dtype=image.dtype)
    grid = torch.stack([xs[None, :].expand(H, W), ys[:, None].expand(H, W)], -1).reshape(-1, 2)
    moved = distort_points(grid, K, dist)
    out = remap(image.reshape(n, C, H, W), moved[..., 0].reshape(n, H, W), moved[..., 1].reshape(n, H, W), align_corners=True)
    return out.view_as(image)


# --------------------------------------------------------------------------------------
# uint8 ingest
# --------------------------------------------------------------------------------------
def image_to_float(image, normalize=True):
    """kornia/image/utils.py:60-73 (image_to_tensor: (H,W,C) -> (1,C,H,W) with keepdim=False, (B,H,W,C) -> (B,C,H,W) by
    permute) followed by kornia/io/io.py:108-111 (_to_float32: image.float() / 255.0)."""
    x = image.unsqueeze(0) if image.dim() == 3 else image
    x = x.permute(0, 3, 1, 2)
    return x.float() / 255.0 if normalize else x.float()


def warp_perspective_from_uint8(image, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True, fill_value=None,
                                normalize=True):
    """The three steps a caller of the reference writes for decoder output: image_to_tensor, _to_float32, warp_perspective
    (imgwarp.py:69)."""
    return warp_perspective(image_to_float(image, normalize), M, dsize, mode, padding_mode, align_corners, fill_value)


def warp_affine_from_uint8(image, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True, fill_value=None,
                           normalize=True):
    """image_to_tensor, _to_float32, warp_affine (imgwarp.py:177)."""
    return warp_affine(image_to_float(image, normalize), M, dsize, mode, padding_mode, align_corners, fill_value)


def undistort_image_from_uint8(image, K, dist, normalize=True):
    """image_to_tensor, _to_float32, undistort_image (calibration/undistort.py:138-198)."""
    x = image_to_float(image, normalize)
    B = x.shape[0]
    Kb = K if K.dim() == 3 else K.expand(B, 3, 3)
    db = dist if dist.dim() == 2 else dist.expand(B, dist.shape[-1])
    return undistort_image(x, Kb, db)
