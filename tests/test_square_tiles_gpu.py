"""The two-shape forward pass (warp_tma.cu): near-identity samples on 64x32 tiles, rotated / sheared / minified
samples on 32x32 tiles with a 56x56 box.  Every sample must be written exactly once and the result must equal, bit
for bit, the single-kernel tiled path and the generic kernel."""
import math

import pytest
import torch

import kornia_b200 as K
from kornia_b200 import _lib
from kornia_b200.geometry import _prelude as P

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _forward_into_nan(src, M, dsize, projective, pad, align, fill=None):
    """kb200_warp_forward into an output prefilled with NaN: anything the kernels skip stays NaN."""
    B, C, H, W = src.shape
    h, w = dsize
    m = P.sampling_matrix(M, (H, W), (h, w), affine=not projective).contiguous()
    bx, by = P.meshgrid_axes(h, w, src.device, src.dtype) if projective else P.affine_axes(h, w, align, src.device, src.dtype)
    out = torch.full((B, C, h, w), float("nan"), device=src.device)
    _lib.call("kb200_warp_forward", src.data_ptr(), m.data_ptr(), bx.contiguous().data_ptr(), by.contiguous().data_ptr(),
              None if fill is None else fill.data_ptr(), out.data_ptr(), B, C, H, W, h, w, m.shape[0], int(projective), _lib.BILINEAR,
              _lib.PADDING[pad], int(align), _lib.F32, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out


def _with_env(name, fn):
    """Run ``fn`` with one kernel family switched off (kornia_b200.config)."""
    option = {"KB200_DISABLE_SQUARE_TILES": "square_tiles", "KB200_DISABLE_TMA": "tma"}[name]
    with K.config.override(**{option: 0}):
        return fn()


def _mixed_affines(B, H, W):
    """identity-ish, small and large rotations, a quarter turn, minification, shear, flips -- one per sample, cycling."""
    cx, cy = (W - 1) / 2, (H - 1) / 2
    mats = []
    recipes = [(0.0, 1.0, 0.0), (3.0, 1.0, 0.0), (30.0, 1.0, 0.0), (45.0, 1.0, 0.0), (90.0, 1.0, 0.0), (-60.0, 0.9, 0.0), (10.0, 0.5, 0.0),
               (0.0, 1.0, 0.6), (180.0, 1.0, 0.0), (-135.0, 1.2, 0.1), (1.0, 1.0, 0.0)]
    for i in range(B):
        ang, sc, sh = recipes[i % len(recipes)]
        a = math.radians(ang)
        c, s = math.cos(a) * sc, math.sin(a) * sc
        A = torch.tensor([[c, s + sh, 0.0], [-s, c, 0.0]])
        A[0, 2] = cx - (A[0, 0] * cx + A[0, 1] * cy)
        A[1, 2] = cy - (A[1, 0] * cx + A[1, 1] * cy)
        mats.append(A)
    return torch.stack(mats).to(DEV)


@pytest.mark.parametrize("C", [3, 1])
@pytest.mark.parametrize("pad", ["zeros", "border", "reflection", "fill"])
@pytest.mark.parametrize("align", [True, False])
def test_affine_mixed_batch_written_once_and_bit_identical(C, pad, align):
    B, H, W = 11, 270, 480
    src = torch.rand(B, C, H, W, device=DEV)
    M = _mixed_affines(B, H, W)
    fill = torch.tensor([0.1, 0.5, 0.9], device=DEV)[:C].contiguous() if pad == "fill" else None
    two = _forward_into_nan(src, M, (301, 500), False, pad, align, fill)
    assert _lib.last_warp_variant() == "tma_tile"
    assert not torch.isnan(two).any(), "a sample was skipped by both kernels"
    one = _with_env("KB200_DISABLE_SQUARE_TILES", lambda: _forward_into_nan(src, M, (301, 500), False, pad, align, fill))
    gen = _with_env("KB200_DISABLE_TMA", lambda: _forward_into_nan(src, M, (301, 500), False, pad, align, fill))
    assert torch.equal(two, one) and torch.equal(two, gen)


@pytest.mark.parametrize("pad", ["zeros", "reflection"])
def test_projective_mixed_batch_written_once_and_bit_identical(pad):
    B, H, W = 12, 270, 480
    src = torch.rand(B, 3, H, W, device=DEV)
    A = _mixed_affines(B, H, W)
    M = torch.eye(3, device=DEV).repeat(B, 1, 1)
    M[:, :2] = A
    g = torch.Generator().manual_seed(3)
    M[:, 2, :2] = (torch.rand(B, 2, generator=g) - 0.5).to(DEV) * 4e-4   # mild keystone on top of the rotations
    two = _forward_into_nan(src, M, (H, W), True, pad, True)
    assert not torch.isnan(two).any()
    one = _with_env("KB200_DISABLE_SQUARE_TILES", lambda: _forward_into_nan(src, M, (H, W), True, pad, True))
    gen = _with_env("KB200_DISABLE_TMA", lambda: _forward_into_nan(src, M, (H, W), True, pad, True))
    assert torch.equal(two, one) and torch.equal(two, gen)


def test_shared_matrix_and_public_rotate():
    """One (1,2,3) matrix for the whole batch (warp_affine's broadcast) and the public rotate() at 1080p."""
    src = torch.rand(5, 3, 540, 960, device=DEV)
    M = _mixed_affines(4, 540, 960)[3:4]  # 45 degrees, shared
    two = _forward_into_nan(src, M, (540, 960), False, "zeros", True)
    gen = _with_env("KB200_DISABLE_TMA", lambda: _forward_into_nan(src, M, (540, 960), False, "zeros", True))
    assert not torch.isnan(two).any() and torch.equal(two, gen)
    big = torch.rand(2, 3, 1080, 1920, device=DEV)
    ang = torch.tensor([33.0, -2.0], device=DEV)
    got = K.geometry.transform.rotate(big, ang)
    want = _with_env("KB200_DISABLE_TMA", lambda: K.geometry.transform.rotate(big, ang))
    assert torch.equal(got, want)


def test_small_problems_stay_on_one_launch():
    src = torch.rand(2, 3, 64, 64, device=DEV)
    M = _mixed_affines(2, 64, 64)
    out = _forward_into_nan(src, M, (64, 64), False, "zeros", True)
    gen = _with_env("KB200_DISABLE_TMA", lambda: _forward_into_nan(src, M, (64, 64), False, "zeros", True))
    assert torch.equal(out, gen)
