"""(B,3,3) prelude of the warps, kept as torch ops on purpose.

These few launches touch 9 floats per sample; keeping the exact op sequence of the reference makes
the matrices handed to the CUDA kernel bit-identical to the reference's on the same device and
leaves d(loss)/dM to torch autograd.  References:
kornia/geometry/conversions.py:342-345,1717-1725,1753-1765, kornia/core/utils.py:159-166,
kornia/geometry/grid.py:65-78, kornia/geometry/transform/imgwarp.py:271-276.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def pixel_to_norm(height: int, width: int, like: torch.Tensor) -> torch.Tensor:
    """[[2/(w-1),0,-1],[0,2/(h-1),-1],[0,0,1]] as (1,3,3); 1e-14 replaces a zero denominator."""
    sx = 2.0 / (1e-14 if width == 1 else width - 1.0)
    sy = 2.0 / (1e-14 if height == 1 else height - 1.0)
    return torch.tensor([[sx, 0.0, -1.0], [0.0, sy, -1.0], [0.0, 0.0, 1.0]]).unsqueeze(0).to(like)


def inverse3x3(a: torch.Tensor) -> torch.Tensor:
    """Adjugate / determinant: rows (b x c, c x a, a x b) over a . (b x c) for columns a, b, c."""
    a0, a1, a2 = a[..., :, 0], a[..., :, 1], a[..., :, 2]
    r0 = torch.linalg.cross(a1, a2, dim=-1)
    r1 = torch.linalg.cross(a2, a0, dim=-1)
    r2 = torch.linalg.cross(a0, a1, dim=-1)
    det = (a0 * r0).sum(-1)
    return torch.stack([r0, r1, r2], dim=-2) / det[..., None, None]


def normalize_homography(M: torch.Tensor, src_hw, dst_hw) -> torch.Tensor:
    """N_dst @ (M @ N_src^-1): pixel homography -> [-1,1] x [-1,1] homography."""
    n_src = pixel_to_norm(src_hw[0], src_hw[1], M)
    n_dst = pixel_to_norm(dst_hw[0], dst_hw[1], M)
    return n_dst @ (M @ inverse3x3(n_src))


def affine_to_homography(A: torch.Tensor) -> torch.Tensor:
    if not isinstance(A, torch.Tensor):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(A)}")
    if not (A.dim() == 3 and tuple(A.shape[-2:]) == (2, 3)):
        raise ValueError(f"Input matrix must be a Bx2x3 tensor. Got {A.shape}")
    Hm = F.pad(A, [0, 0, 0, 1], "constant", value=0.0)
    Hm[..., -1, -1] += 1.0
    return Hm


# The base-grid axes depend only on (h, w, device, dtype): built once with the reference's own torch ops
# (so they are bit-identical to what create_meshgrid / linspace produce on that device) and reused.
_AXES_CACHE: dict = {}
_AXES_CACHE_MAX = 64


def _cached(key, build):
    hit = _AXES_CACHE.get(key)
    if hit is None:
        if len(_AXES_CACHE) >= _AXES_CACHE_MAX:
            _AXES_CACHE.clear()
        hit = _AXES_CACHE[key] = build()
    return hit


def meshgrid_axes(h: int, w: int, device, dtype):
    """The two axes of create_meshgrid(normalized=True): built in fp32, then cast (imgwarp.py:157)."""

    def build():
        xs = torch.linspace(0, w - 1, w, device=device)
        ys = torch.linspace(0, h - 1, h, device=device)
        xs = (xs / (w - 1) - 0.5) * 2
        ys = (ys / (h - 1) - 0.5) * 2
        return xs.to(dtype), ys.to(dtype)

    return _cached(("mesh", h, w, str(device), dtype), build)


def affine_axes(h: int, w: int, align_corners: bool, device, dtype):
    def build():
        if align_corners:
            return (torch.linspace(-1.0, 1.0, w, device=device, dtype=dtype),
                    torch.linspace(-1.0, 1.0, h, device=device, dtype=dtype))
        return (torch.linspace(-1.0 + 1.0 / w, 1.0 - 1.0 / w, w, device=device, dtype=dtype),
                torch.linspace(-1.0 + 1.0 / h, 1.0 - 1.0 / h, h, device=device, dtype=dtype))

    return _cached(("affine", h, w, bool(align_corners), str(device), dtype), build)


# ---------------------------------------------------------------------------------------------
# one-launch prelude (kb200_warp_prelude) for the common no-grad-on-M case
# ---------------------------------------------------------------------------------------------
FUSED_VARIANT = 4       # the contraction orders that reproduce torch's CUDA kernels bit for bit (GPU-tested)
FUSED_MIN_BATCH = 2     # below this torch's bmm takes a different (gemv-like) path; keep the torch ops there


class _FusedPrelude(torch.autograd.Function):
    """kb200_warp_prelude forward (bit-identical to the torch op sequence) + its one-launch backward."""

    @staticmethod
    def forward(ctx, M, src_hw, dst_hw, affine):
        from .. import _lib, _ops

        Mc = M.contiguous()
        out = torch.empty((Mc.shape[0], 3, 3), device=M.device, dtype=M.dtype)
        dt = 0 if M.dtype == torch.float32 else 1
        with torch.cuda.device(M.device):
            _lib.call("kb200_warp_prelude", Mc.data_ptr(), out.data_ptr(), Mc.shape[0], 2 if affine else 3, int(src_hw[0]), int(src_hw[1]),
                      int(dst_hw[0]), int(dst_hw[1]), dt, FUSED_VARIANT, torch.cuda.current_stream(M.device).cuda_stream)
        _ops._bump()
        ctx.save_for_backward(out)
        ctx.cfg = (tuple(int(v) for v in src_hw), tuple(int(v) for v in dst_hw), bool(affine), dt)
        return out

    @staticmethod
    def backward(ctx, gm):
        from .. import _lib, _ops

        (m,) = ctx.saved_tensors
        src_hw, dst_hw, affine, dt = ctx.cfg
        rows = 2 if affine else 3
        gm = gm.contiguous()
        gM = torch.empty((m.shape[0], rows, 3), device=m.device, dtype=m.dtype)
        with torch.cuda.device(m.device):
            _lib.call("kb200_warp_prelude_backward", m.data_ptr(), gm.data_ptr(), gM.data_ptr(), m.shape[0], rows, src_hw[0], src_hw[1],
                      dst_hw[0], dst_hw[1], dt, torch.cuda.current_stream(m.device).cuda_stream)
        _ops._bump()
        return gM, None, None, None


def sampling_matrix(M: torch.Tensor, src_hw, dst_hw, affine: bool) -> torch.Tensor:
    """inverse(normalize_homography(M3)) -- the (B,3,3) dst-normalised -> src-normalised map the kernels
    consume.  One CUDA launch (and one more for its backward) when M is a CUDA fp32/fp64 tensor with batch >= 2;
    the reference's torch op sequence otherwise, under double backward, or when KORNIA_B200_TORCH_PRELUDE=1."""
    import os

    fused_ok = (M.is_cuda and M.dtype in (torch.float32, torch.float64) and M.shape[0] >= FUSED_MIN_BATCH
                and os.environ.get("KORNIA_B200_TORCH_PRELUDE", "0") != "1")
    if not fused_ok:
        M3 = affine_to_homography(M) if affine else M
        return inverse3x3(normalize_homography(M3, src_hw, dst_hw))
    return _FusedPrelude.apply(M, src_hw, dst_hw, affine)
