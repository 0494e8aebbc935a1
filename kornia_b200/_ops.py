"""torch.autograd bridges onto the C ABI.  Device memory, streams and autograd bookkeeping are
torch's; every FLOP on an image is executed by kornia_b200/csrc kernels."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float64: _lib.F64}

# number of kernels of this library launched since import (bench.py reports it as gpu_launches)
launch_count = 0


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"kornia_b200: {what} lives on {t.device}; this engine is CUDA-only (sm_100a) and has no CPU path. "
            "Move the tensors to a B200 (`.cuda()`)."
        )


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"kornia_b200: unsupported dtype {t.dtype}; float32 and float64 are implemented") from None


def _same(ref: torch.Tensor, other: torch.Tensor, what: str) -> None:
    if other.device != ref.device:
        raise RuntimeError(f"Expected all tensors to be on the same device, but {what} is on {other.device} and the image on {ref.device}")
    if other.dtype != ref.dtype:
        raise RuntimeError(f"expected {what} to have the same dtype as the image, but got {other.dtype} and {ref.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _bump(n: int = 1) -> None:
    global launch_count
    launch_count += n


# Optional per-kernel CUDA-event timing of the forward kernels (bench.py's roofline leg): when a list
# is installed here, every forward C call is bracketed by events recorded on the launching stream.
kernel_events = None


class _Timed:
    def __init__(self, tag: str, ref: torch.Tensor):
        self.tag, self.ref = tag, ref

    def __enter__(self):
        if kernel_events is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.stop = torch.cuda.Event(enable_timing=True)
            self.start.record(torch.cuda.current_stream(self.ref.device))
        return self

    def __exit__(self, *exc):
        if kernel_events is not None and exc[0] is None:
            self.stop.record(torch.cuda.current_stream(self.ref.device))
            kernel_events.append((self.tag, self.start, self.stop))
        return False


class WarpFunction(torch.autograd.Function):
    """out = sample(src, map(m, bx, by)); differentiable w.r.t. ``src`` and ``m``."""

    @staticmethod
    def forward(ctx, src, m, bx, by, fill, h, w, projective, interp, pad, align):
        _require_cuda(src, "src")
        _same(src, m, "the transformation matrix")
        dt = _dtype_code(src)
        src_c = src.contiguous()
        m_c = m.contiguous()
        bx = bx.to(device=src.device, dtype=src.dtype).contiguous()
        by = by.to(device=src.device, dtype=src.dtype).contiguous()
        fill_c = None if fill is None else fill.to(device=src.device, dtype=src.dtype).contiguous()
        B, C, H, W = src_c.shape
        out = torch.empty((B, C, h, w), device=src.device, dtype=src.dtype)
        if out.numel() > 0 and src_c.numel() > 0:
            with torch.cuda.device(src.device), _Timed("warp_forward", src):
                _lib.call("kb200_warp_forward", _ptr(src_c), _ptr(m_c), _ptr(bx), _ptr(by), _ptr(fill_c), _ptr(out),
                          B, C, H, W, h, w, m_c.shape[0], int(projective), interp, pad, int(align), dt, _stream(src))
            _bump(_lib.last_warp_launches())
        else:
            out.zero_()
        ctx.save_for_backward(src_c, m_c, bx, by, fill_c if fill_c is not None else torch.empty(0, device=src.device))
        ctx.cfg = (h, w, int(projective), interp, pad, int(align), dt, fill_c is not None)
        return out

    @staticmethod
    def backward(ctx, gout):
        src, m, bx, by, fill = ctx.saved_tensors
        h, w, projective, interp, pad, align, dt, has_fill = ctx.cfg
        need_src, need_m = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_src or need_m):
            return (None,) * 11
        B, C, H, W = src.shape
        gout = gout.contiguous()
        gsrc = torch.zeros_like(src) if need_src else None
        if gout.numel() == 0 or src.numel() == 0:
            return (gsrc, torch.zeros_like(m) if need_m else None) + (None,) * 9
        gm = ws = None
        if need_m:
            gm = torch.empty_like(m)
            nbytes = _lib.load().kb200_warp_backward_workspace_bytes(B, h, w, dt)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=src.device)
        with torch.cuda.device(src.device), _Timed("warp_backward", src):
            _lib.call("kb200_warp_backward", _ptr(gout), _ptr(src), _ptr(m), _ptr(bx), _ptr(by),
                      _ptr(fill) if has_fill else None, _ptr(gsrc), _ptr(gm), _ptr(ws),
                      B, C, H, W, h, w, m.shape[0], projective, interp, pad, align, dt, _stream(src))
        _bump(2 if need_m else 1)
        return (gsrc, gm) + (None,) * 9


def warp_u8hwc(image: torch.Tensor, m: torch.Tensor, bx: torch.Tensor, by: torch.Tensor, fill: Optional[torch.Tensor], h: int, w: int,
               projective: bool, interp: int, pad: int, align: bool, normalize: int) -> torch.Tensor:
    """Warp of an interleaved uint8 batch (B,H,W,C) into planar fp32 (B,C,h,w) in one kernel (kb200_warp_u8hwc_forward):
    the bytes are converted tap by tap inside the sampler (``normalize``: 0 raw, 1 times 1/255 as torch's CUDA backend
    does, 2 divided by 255).  ``m`` is the (B|1,3,3) fp32 sampling matrix of the warp prelude.  Forward only."""
    _require_cuda(image, "image")
    if image.dtype != torch.uint8:
        raise RuntimeError(f"kornia_b200: expected a uint8 image, got {image.dtype}")
    if m.device != image.device:
        raise RuntimeError(f"Expected all tensors to be on the same device, but the transformation matrix is on {m.device} and the image on {image.device}")
    img = image.contiguous()
    f32 = dict(device=image.device, dtype=torch.float32)
    m_c, bx, by = m.to(**f32).contiguous(), bx.to(**f32).contiguous(), by.to(**f32).contiguous()
    fill_c = None if fill is None else fill.to(**f32).contiguous()
    B, H, W, C = img.shape
    out = torch.empty((B, C, h, w), **f32)
    if out.numel() > 0 and img.numel() > 0:
        with torch.cuda.device(image.device), _Timed("warp_u8hwc_forward", image):
            _lib.call("kb200_warp_u8hwc_forward", _ptr(img), _ptr(m_c), _ptr(bx), _ptr(by), _ptr(fill_c), _ptr(out), B, C, H, W, h, w,
                      m_c.shape[0], int(projective), interp, pad, int(align), int(normalize), _stream(image))
        _bump()
    else:
        out.zero_()
    return out


def undistort_u8hwc(image: torch.Tensor, lens: torch.Tensor, normalize: int) -> torch.Tensor:
    """undistort_image of an interleaved uint8 batch (B,H,W,C) -> planar fp32 (B,C,H,W) in one kernel
    (kb200_undistort_u8hwc_forward): lens (B,16) as for undistort_fused.  Raises ``_lib.Unsupported`` outside the kernel's
    envelope (the caller converts the image and takes the fp32 path).  Forward only."""
    _require_cuda(image, "image")
    if image.dtype != torch.uint8:
        raise RuntimeError(f"kornia_b200: expected a uint8 image, got {image.dtype}")
    img = image.contiguous()
    ln = lens.to(device=image.device, dtype=torch.float32).contiguous()
    B, H, W, C = img.shape
    out = torch.empty((B, C, H, W), device=image.device, dtype=torch.float32)
    if out.numel() == 0:
        raise _lib.Unsupported("empty image")
    with torch.cuda.device(image.device), _Timed("undistort_u8hwc_forward", image):
        _lib.call("kb200_undistort_u8hwc_forward", _ptr(img), _ptr(ln), _ptr(out), B, C, H, W, int(normalize), _stream(image))
    _bump()
    return out


class RemapFunction(torch.autograd.Function):
    """out = sample(image, (map_x, map_y)); differentiable w.r.t. the image and both maps."""

    @staticmethod
    def forward(ctx, image, map_x, map_y, normalized, interp, pad, align):
        _require_cuda(image, "image")
        _same(image, map_x, "map_x")
        _same(image, map_y, "map_y")
        dt = _dtype_code(image)
        img = image.contiguous()
        mx, my = map_x.contiguous(), map_y.contiguous()
        B, C, H, W = img.shape
        Bmap, h, w = mx.shape
        out = torch.empty((B, C, h, w), device=img.device, dtype=img.dtype)
        with torch.cuda.device(img.device):
            _lib.call("kb200_remap_forward", _ptr(img), _ptr(mx), _ptr(my), _ptr(out), B, C, H, W, h, w, Bmap,
                      int(normalized), interp, pad, int(align), dt, _stream(img))
        _bump()
        ctx.save_for_backward(img, mx, my)
        ctx.cfg = (int(normalized), interp, pad, int(align), dt)
        return out

    @staticmethod
    def backward(ctx, gout):
        img, mx, my = ctx.saved_tensors
        normalized, interp, pad, align, dt = ctx.cfg
        need_img = ctx.needs_input_grad[0]
        need_map = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        if not (need_img or need_map):
            return (None,) * 7
        B, C, H, W = img.shape
        Bmap, h, w = mx.shape
        gout = gout.contiguous()
        gimg = torch.zeros_like(img) if need_img else None
        gmx = gmy = None
        if need_map:
            gmx = torch.empty((B, h, w), device=img.device, dtype=img.dtype)
            gmy = torch.empty_like(gmx)
        with torch.cuda.device(img.device):
            _lib.call("kb200_remap_backward", _ptr(gout), _ptr(img), _ptr(mx), _ptr(my), _ptr(gimg), _ptr(gmx), _ptr(gmy),
                      B, C, H, W, h, w, Bmap, normalized, interp, pad, align, dt, _stream(img))
        _bump()
        if need_map and Bmap == 1 and B > 1:  # maps were broadcast over the batch (imgwarp.py:695)
            gmx, gmy = gmx.sum(0, keepdim=True), gmy.sum(0, keepdim=True)
        return gimg, gmx if ctx.needs_input_grad[1] else None, gmy if ctx.needs_input_grad[2] else None, None, None, None, None


class Filter2dFunction(torch.autograd.Function):
    """Depthwise correlation with fused border handling; ``kernel`` is (Bk,kh,kw), already
    flipped / normalised.  Differentiable w.r.t. input and kernel."""

    @staticmethod
    def forward(ctx, x, kernel, border, same):
        _require_cuda(x, "input")
        dt = _dtype_code(x)
        xc = x.contiguous()
        kc = kernel.contiguous()
        B, C, H, W = xc.shape
        Bk, kh, kw = kc.shape
        if B % Bk != 0:
            # the reference's view(-1, Bk*C, H, W) (filter.py:142) fails the same way
            raise RuntimeError(f"shape '[-1, {Bk * C}, {H}, {W}]' is invalid for input of size {xc.numel()}")
        Ho, Wo = (H, W) if same else (H - kh + 1, W - kw + 1)
        out = torch.empty((B, C, max(Ho, 0), max(Wo, 0)), device=x.device, dtype=x.dtype)
        if out.numel() > 0:
            with torch.cuda.device(x.device):
                _lib.call("kb200_filter2d_forward", _ptr(xc), _ptr(kc), _ptr(out), B, C, H, W, Bk, kh, kw, border, int(same), dt,
                          _stream(x))
            _bump()
        ctx.save_for_backward(xc, kc)
        ctx.cfg = (border, int(same), dt)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, k = ctx.saved_tensors
        border, same, dt = ctx.cfg
        B, C, H, W = x.shape
        Bk, kh, kw = k.shape
        gout = gout.contiguous()
        gx = gk = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                _lib.call("kb200_filter2d_backward_input", _ptr(gout), _ptr(k), _ptr(gx), B, C, H, W, Bk, kh, kw, border, same,
                          dt, _stream(x))
                _bump()
            if ctx.needs_input_grad[1]:
                gk = torch.empty_like(k)
                nbytes = _lib.load().kb200_filter2d_backward_kernel_workspace_bytes(B, C, H, W, Bk, kh, kw, dt)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                _lib.call("kb200_filter2d_backward_kernel", _ptr(gout), _ptr(x), _ptr(gk), _ptr(ws), B, C, H, W, Bk, kh, kw,
                          border, same, dt, _stream(x))
                _bump(2)
        return gx, gk, None, None


def undistort_fused(image: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """Lens model + bilinear resampling in one kernel (forward only): image (B,C,H,W), lens (B,16).  Raises
    ``_lib.Unsupported`` outside the kernel's envelope (the caller then builds the maps and calls remap)."""
    _require_cuda(image, "image")
    dt = _dtype_code(image)
    img, ln = image.contiguous(), lens.to(device=image.device, dtype=image.dtype).contiguous()
    B, C, H, W = img.shape
    out = torch.empty_like(img)
    if out.numel() == 0:
        raise _lib.Unsupported("empty image")
    with torch.cuda.device(image.device), _Timed("undistort_forward", image):
        _lib.call("kb200_undistort_forward", _ptr(img), _ptr(ln), _ptr(out), B, C, H, W, dt, _stream(image))
    _bump()
    return out


def pyrdown_fused(x: torch.Tensor, kernel: torch.Tensor, border: int) -> torch.Tensor:
    """5x5 correlation + exact 2x bilinear decimation in one kernel (forward only): (B,C,H,W) -> (B,C,H/2,W/2).
    Raises ``_lib.Unsupported`` outside the kernel's envelope (the caller then composes filter2d + interpolate)."""
    _require_cuda(x, "input")
    dt = _dtype_code(x)
    xc, kc = x.contiguous(), kernel.contiguous()
    B, C, H, W = xc.shape
    out = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=x.dtype)
    if out.numel() == 0:
        raise _lib.Unsupported("empty output")
    with torch.cuda.device(x.device), _Timed("pyrdown_forward", x):
        _lib.call("kb200_pyrdown_forward", _ptr(xc), _ptr(kc), _ptr(out), B, C, H, W, kc.shape[0], border, dt, _stream(x))
    _bump()
    return out


class SepFilterFunction(torch.autograd.Function):
    """Row pass then column pass in one kernel (one read + one write of the image).  The backward
    composes the 1-D adjoints: g_mid = Fy^T(gout), gx = Fx^T(g_mid), with mid = Fx(x) recomputed."""

    @staticmethod
    def forward(ctx, x, kx, ky, border, same):
        _require_cuda(x, "input")
        dt = _dtype_code(x)
        xc, kxc, kyc = x.contiguous(), kx.contiguous(), ky.contiguous()
        B, C, H, W = xc.shape
        (Bkx, kw), (Bky, kh) = kxc.shape, kyc.shape
        if B % Bkx != 0 or B % Bky != 0:
            raise RuntimeError(f"shape '[-1, {max(Bkx, Bky) * C}, {H}, {W}]' is invalid for input of size {xc.numel()}")
        Ho, Wo = (H, W) if same else (H - kh + 1, W - kw + 1)
        out = torch.empty((B, C, Ho, Wo), device=x.device, dtype=x.dtype)
        if out.numel() > 0:
            try:
                with torch.cuda.device(x.device), _Timed("sepfilter_forward", x):
                    _lib.call("kb200_sepfilter_forward", _ptr(xc), _ptr(kxc), _ptr(kyc), _ptr(out), B, C, H, W, Bkx, kw, Bky, kh, border,
                              int(same), dt, _stream(x))
                _bump()
            except _lib.Unsupported:
                # kernels too large for the one-pass shared-memory tile: two 1-D passes of the 2-D kernel
                mid = Filter2dFunction.apply(xc, kxc[:, None, :], border, same)
                out = Filter2dFunction.apply(mid, kyc[:, :, None], border, same)
        ctx.save_for_backward(xc, kxc, kyc)
        ctx.cfg = (border, int(same))
        return out

    @staticmethod
    def backward(ctx, gout):
        x, kx, ky = ctx.saved_tensors
        border, same = ctx.cfg
        need = ctx.needs_input_grad
        if need[0] and not (need[1] or need[2]) and same and os.environ.get("KB200_FAST_FILTER_BWD") == "1":
            # opt-in: input-only gradient without rebuilding the forward through the generic kernels
            return _sep_input_gradient(gout.contiguous(), kx, ky, border), None, None, None, None
        with torch.enable_grad():
            xl = x.detach().requires_grad_(need[0])
            kxl = kx.detach().requires_grad_(need[1])
            kyl = ky.detach().requires_grad_(need[2])
            mid = Filter2dFunction.apply(xl, kxl[:, None, :], border, same)
            out = Filter2dFunction.apply(mid, kyl[:, :, None], border, same)
            wrt = [t for t, n in zip((xl, kxl, kyl), need[:3]) if n]
            grads = list(torch.autograd.grad(out, wrt, gout.contiguous())) if wrt else []
        res = [grads.pop(0) if n else None for n in need[:3]]
        return res[0], res[1], res[2], None, None


def _filter2d_backward_input(gout: torch.Tensor, kernel: torch.Tensor, border: int) -> torch.Tensor:
    """The C call Filter2dFunction.backward makes for d/dinput ('same'), without the autograd graph around it."""
    B, C, H, W = gout.shape
    Bk, kh, kw = kernel.shape
    gx = torch.empty_like(gout)
    if gx.numel() > 0:
        with torch.cuda.device(gout.device):
            _lib.call("kb200_filter2d_backward_input", _ptr(gout), _ptr(kernel), _ptr(gx), B, C, H, W, Bk, kh, kw, border, 1, _dtype_code(gout),
                      _stream(gout))
        _bump()
    return gx


def _sep_input_gradient(gout: torch.Tensor, kx: torch.Tensor, ky: torch.Tensor, border: int) -> torch.Tensor:
    """d/dinput of the 'same' separable filter.  Exact form: the two adjoint passes of the composition (the very calls
    autograd makes, minus the two forward passes it rebuilds first).  When the one-pass tiled kernel covers the shape, the
    image-sized work runs through it instead (adjoint = correlation with the flipped taps under a 'constant' border) and
    only the border bands take the exact form (filters/_adjoint.py)."""
    from .filters._adjoint import separable_adjoint

    def exact(g, kx_, ky_):
        g_mid = _filter2d_backward_input(g, ky_[:, :, None].contiguous(), border)
        return _filter2d_backward_input(g_mid, kx_[:, None, :].contiguous(), border)

    def forward_constant(g, kx_, ky_):
        out = torch.empty_like(g)
        B, C, H, W = g.shape
        with torch.cuda.device(g.device):
            _lib.call("kb200_sepfilter_forward", _ptr(g), _ptr(kx_.contiguous()), _ptr(ky_.contiguous()), _ptr(out), B, C, H, W, kx_.shape[0],
                      kx_.shape[1], ky_.shape[0], ky_.shape[1], _lib.CONSTANT, 1, _dtype_code(g), _stream(g))
        _bump()
        return out

    kw, kh = kx.shape[-1], ky.shape[-1]
    fast = gout.dtype == torch.float32 and kw == kh and kw % 2 == 1 and 3 <= kw <= 17 and border != _lib.CIRCULAR and gout.shape[-1] % 4 == 0
    return separable_adjoint(gout, kx, ky, border, forward_constant, exact) if fast else exact(gout, kx, ky)


class SpatialGradientFunction(torch.autograd.Function):
    """``nout`` k x k derivative stencils over a replicate border in one kernel: x (B,C,H,W) ->
    (B,C,nout,H,W), or the Sobel magnitude (B,C,H,W) when ``magnitude`` (forward only).  ``taps`` is
    a host tuple of nout*k*k floats whose values are exact in x.dtype."""

    @staticmethod
    def forward(ctx, x, taps, nout, k, magnitude, eps):
        _require_cuda(x, "input")
        dt = _dtype_code(x)
        xc = x.contiguous()
        B, C, H, W = xc.shape
        shape = (B, C, H, W) if magnitude else (B, C, nout, H, W)
        out = torch.empty(shape, device=x.device, dtype=x.dtype)
        host_taps = (ctypes.c_double * len(taps))(*taps)
        if out.numel() > 0:
            with torch.cuda.device(x.device), _Timed("spatial_gradient_forward", x):
                _lib.call("kb200_spatial_gradient_forward", _ptr(xc), host_taps, _ptr(out), B * C, H, W, nout, k, int(magnitude),
                          float(eps), dt, _stream(x))
            _bump()
        ctx.cfg = (host_taps, nout, k, dt, bool(magnitude))
        return out

    @staticmethod
    def backward(ctx, gout):
        host_taps, nout, k, dt, magnitude = ctx.cfg
        if magnitude:
            raise RuntimeError("kornia_b200: the fused Sobel magnitude is forward-only; sobel() composes the "
                               "differentiable path when the input requires grad")
        gout = gout.contiguous()
        B, C, _, H, W = gout.shape
        gx = torch.empty((B, C, H, W), device=gout.device, dtype=gout.dtype)
        if gx.numel() > 0:
            with torch.cuda.device(gout.device):
                _lib.call("kb200_spatial_gradient_backward", _ptr(gout), host_taps, _ptr(gx), B * C, H, W, nout, k, dt, _stream(gout))
            _bump()
        return gx, None, None, None, None, None
