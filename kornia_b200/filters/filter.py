"""Drop-in ``filter2d`` / ``filter2d_separable`` (reference: kornia/filters/filter.py:54-207).

The reference materialises a padded copy (``F.pad``) and runs a grouped ``F.conv2d`` per pass;
here the border mode is index arithmetic inside the CUDA kernel and the separable filter is a
single pass over HBM.
"""
from __future__ import annotations

import torch

from .. import _lib
from .. import _ops
from ..core.check import check, check_is_tensor, check_shape
from .kernels import normalize_kernel2d

_VALID_BORDERS = {"constant", "reflect", "replicate", "circular"}
_VALID_PADDING = {"valid", "same"}
_VALID_BEHAVIOUR = {"conv", "corr"}

__all__ = ["filter2d", "filter2d_separable"]


def _compute_padding(kernel_size: list[int]) -> list[int]:
    """F.pad-style list (last dimension first): front = (k-1)//2, rear = (k-1) - front."""
    if len(kernel_size) < 2:
        raise AssertionError(kernel_size)
    out: list[int] = []
    for k in reversed(kernel_size):
        front = (k - 1) // 2
        out += [front, (k - 1) - front]
    return out


def _border_code(border_type: str, x: torch.Tensor, kh: int, kw: int, same: bool) -> int:
    code = _lib.BORDERS.get(str(border_type))
    if code is None:  # only reachable with checks disabled; F.pad would raise here
        raise NotImplementedError(f"Unrecognised padding mode {border_type}")
    if same:
        H, W = x.shape[-2:]
        ph, pw = (kh - 1) - (kh - 1) // 2, (kw - 1) - (kw - 1) // 2
        if code == _lib.REFLECT and (ph >= H or pw >= W):
            raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: padding "
                               f"({(kw - 1) // 2}, {pw}, {(kh - 1) // 2}, {ph}) at dimension of input {list(x.shape)}")
        if code == _lib.CIRCULAR and (ph > H or pw > W):
            raise RuntimeError("Padding value causes wrapping around more than once.")
    return code


def filter2d(
    input: torch.Tensor,
    kernel: torch.Tensor,
    border_type: str = "reflect",
    normalized: bool = False,
    padding: str = "same",
    behaviour: str = "corr",
) -> torch.Tensor:
    """Filter every channel of ``input`` (B,C,H,W) with ``kernel`` ((1,kH,kW) shared or (B,kH,kW)
    per sample): cross-correlation by default, true convolution with ``behaviour='conv'``.

    ``border_type``: 'constant' | 'reflect' | 'replicate' | 'circular' (used with
    ``padding='same'``; 'valid' shrinks the output instead).  ``normalized`` divides the kernel
    by the sum of its absolute taps.  A kernel batch Bk that is neither 1 nor B cycles over the
    samples (sample b uses kernel b mod Bk), as the reference's view/grouped-conv does.
    """
    check_is_tensor(input)
    check_shape(input, ["B", "C", "H", "W"])
    check_is_tensor(kernel)
    check_shape(kernel, ["B", "H", "W"])
    check(str(border_type).lower() in _VALID_BORDERS, f"Invalid border, {border_type}. Expected one of {_VALID_BORDERS}")
    check(str(padding).lower() in _VALID_PADDING, f"Invalid padding mode, {padding}. Expected one of {_VALID_PADDING}")
    check(str(behaviour).lower() in _VALID_BEHAVIOUR, f"Invalid padding mode, {behaviour}. Expected one of {_VALID_BEHAVIOUR}")

    taps = kernel.flip((-2, -1)) if str(behaviour).lower() == "conv" else kernel
    taps = taps.to(device=input.device, dtype=input.dtype)
    if normalized:
        taps = normalize_kernel2d(taps)
    same = padding == "same"
    kh, kw = taps.shape[-2:]
    return _ops.filter2d(input, taps, _border_code(border_type, input, kh, kw, same), same)


def filter2d_separable(
    input: torch.Tensor,
    kernel_x: torch.Tensor,
    kernel_y: torch.Tensor,
    border_type: str = "reflect",
    normalized: bool = False,
    padding: str = "same",
) -> torch.Tensor:
    """Filter with a horizontal kernel ``kernel_x`` ((1,kW) or (B,kW)) and then a vertical one
    ``kernel_y`` ((1,kH) or (B,kH)).  Equivalent to two :func:`filter2d` calls (filter.py:205-207)
    but executed by one kernel that reads and writes the image once."""
    # the two nested filter2d calls of the reference validate in this order
    check_is_tensor(input)
    check_shape(input, ["B", "C", "H", "W"])
    check_is_tensor(kernel_x)
    check_shape(kernel_x[..., None, :], ["B", "H", "W"])
    check(str(border_type).lower() in _VALID_BORDERS, f"Invalid border, {border_type}. Expected one of {_VALID_BORDERS}")
    check(str(padding).lower() in _VALID_PADDING, f"Invalid padding mode, {padding}. Expected one of {_VALID_PADDING}")
    check_is_tensor(kernel_y)
    check_shape(kernel_y[..., None], ["B", "H", "W"])

    kx = kernel_x.to(device=input.device, dtype=input.dtype)
    ky = kernel_y.to(device=input.device, dtype=input.dtype)
    if normalized:
        kx = normalize_kernel2d(kx[:, None, :])[:, 0, :]
        ky = normalize_kernel2d(ky[:, :, None])[:, :, 0]
    same = padding == "same"
    code = _border_code(border_type, input, ky.shape[-1], kx.shape[-1], same)
    return _ops.sepfilter(input, kx, ky, code, same)
