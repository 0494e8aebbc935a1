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
    """Process-wide cache of small constant tensors.  Entries are built outside any inference-mode / grad scope of the
    caller: a tensor created under ``torch.inference_mode()`` could not be saved for a later backward (the reference
    has no such state to trip over).  Under ``torch.compile`` tracing the tensors are rebuilt in-graph (no global state)."""
    if torch.compiler.is_compiling():
        return build()
    hit = _AXES_CACHE.get(key)
    if hit is None:
        if len(_AXES_CACHE) >= _AXES_CACHE_MAX:
            _AXES_CACHE.clear()
        with torch.inference_mode(False), torch.no_grad():
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
# one-launch prelude (torch.ops.kornia_b200.warp_prelude, autograd formula = warp_prelude_bwd)
# ---------------------------------------------------------------------------------------------
FUSED_MIN_BATCH = 2     # below this torch's bmm takes a different (gemv-like) path; keep the torch ops there


def torch_prelude_forced() -> bool:
    from .. import config

    return config.enabled("torch_prelude")


def sampling_matrix(M: torch.Tensor, src_hw, dst_hw, affine: bool) -> torch.Tensor:
    """inverse(normalize_homography(M3)) -- the (B,3,3) dst-normalised -> src-normalised map the kernels
    consume.  One CUDA launch (and one more for its backward) when M is a CUDA fp32/fp64 tensor with batch >= 2;
    the reference's torch op sequence otherwise or when ``config.set("torch_prelude", 1)`` (KB200_TORCH_PRELUDE=1).  The fused
    backward has no autograd formula of its own: double backward through M needs that switch."""
    from .. import _ops

    fused_ok = (M.is_cuda and M.dtype in (torch.float32, torch.float64) and M.shape[0] >= FUSED_MIN_BATCH and not torch_prelude_forced())
    if not fused_ok:
        M3 = affine_to_homography(M) if affine else M
        return inverse3x3(normalize_homography(M3, src_hw, dst_hw))
    return _ops.prelude(M, src_hw[0], src_hw[1], dst_hw[0], dst_hw[1], affine)
