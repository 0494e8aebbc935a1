"""Timing of the opt-in kernels of DESIGN.md section 9 against the defaults they would replace, on one GPU
(B x 3 x 1080 x 1920 fp32, CUDA events, 3 warm-ups, inputs larger than L2).  Each pair is first compared bit for
bit; a variant that differs is reported as MISMATCH and its time is not a result.

    python tools/bench_unverified.py            # B = 64;  BENCH_B=256 for the BASELINE batch of cfg3
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import kornia_b200 as K  # noqa: E402

dev = "cuda"
B = int(os.environ.get("BENCH_B", "64"))
H, W = 1080, 1920
peak = 6568.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
x = torch.rand(B, 3, H, W, device=dev)
y = (x + 0.05 * torch.randn_like(x)).clamp_(0, 1)
elems = B * 3 * H * W


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def pair(name, switch, fn, bytes_per_elem):
    os.environ.pop(switch, None)
    ref = fn()
    t0 = timed(fn)
    os.environ[switch] = "1"
    try:
        got = fn()
        same = torch.equal(got, ref)
        t1 = timed(fn)
    except Exception as exc:  # keep going: one broken variant must not hide the others
        print(f"{name:34s} default {t0:7.3f} ms | {switch}=1 FAILED: {exc}")
        os.environ.pop(switch, None)
        return
    os.environ.pop(switch, None)
    gb = elems * bytes_per_elem / 1e9
    print(f"{name:34s} default {t0:7.3f} ms ({gb / t0 / peak * 1e3 * 100:5.1f} % HBM) | {switch}=1 {t1:7.3f} ms "
          f"({gb / t1 / peak * 1e3 * 100:5.1f} % HBM) | x{t0 / t1:4.2f} | {'bit-identical' if same else 'MISMATCH'}")
    del ref, got


def sweep(name, switch, fn):
    """The opt-in persistent kernels at 1 / 2 / 3 resident CTAs per SM (KB200_GRID_PER_SM; values above a kernel's launch bounds
    are ignored by the library)."""
    os.environ[switch] = "1"
    out = []
    for g in ("1", "2", "3"):
        os.environ["KB200_GRID_PER_SM"] = g
        try:
            out.append(f"{g}: {timed(fn):7.3f} ms")
        except Exception as exc:
            out.append(f"{g}: FAILED {exc}")
    os.environ.pop("KB200_GRID_PER_SM", None)
    os.environ.pop(switch, None)
    print(f"{name:34s} {switch}=1, CTAs per SM the grid is sized for -> " + " | ".join(out))


with torch.no_grad():
    print(f"B={B} x 3 x {H} x {W} fp32, measured HBM peak {peak:.0f} GB/s; bytes = algorithmic bytes of the fused op")
    pair("gaussian_blur2d k=11 (cfg3)", "KB200_SEP_VWALK", lambda: K.gaussian_blur2d(x, (11, 11), (2.0, 2.0)), 8)
    pair("gaussian_blur2d k=5", "KB200_SEP_VWALK", lambda: K.gaussian_blur2d(x, (5, 5), (1.0, 1.0)), 8)
    pair("gaussian_blur2d k=17", "KB200_SEP_VWALK", lambda: K.gaussian_blur2d(x, (17, 17), (3.0, 3.0)), 8)
    pair("unsharp_mask 5x5", "KB200_SEP_VWALK", lambda: K.filters.unsharp_mask(x, (5, 5), (1.5, 1.5)), 8)
    pair("ssim window 11", "KB200_SSIM_VWALK", lambda: K.metrics.ssim(x, y, 11), 12)
    pair("ssim window 5", "KB200_SSIM_VWALK", lambda: K.metrics.ssim(x, y, 5), 12)
    pair("spatial_gradient sobel order 1", "KB200_TILED_GRADIENT", lambda: K.filters.spatial_gradient(x, "sobel", 1), 12)
    pair("spatial_gradient sobel order 2", "KB200_TILED_GRADIENT", lambda: K.filters.spatial_gradient(x, "sobel", 2), 16)
    pair("spatial_gradient diff order 2", "KB200_TILED_GRADIENT", lambda: K.filters.spatial_gradient(x, "diff", 2), 16)
    pair("sobel magnitude", "KB200_TILED_GRADIENT", lambda: K.filters.sobel(x), 8)
    pair("pyrdown", "KB200_FUSED_PYRDOWN", lambda: K.geometry.transform.pyrdown(x), 5)
    cam = torch.tensor([[1500.0, 0.0, 960.0], [0.0, 1500.0, 540.0], [0.0, 0.0, 1.0]], device=dev).expand(B, 3, 3).contiguous()
    dist = torch.tensor([[-0.2, 0.05, 0.001, -0.002, 0.01]], device=dev).expand(B, 5).contiguous()
    yy, xx = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    r2 = ((xx - W / 2) / W) ** 2 + ((yy - H / 2) / H) ** 2
    mx, my = (xx + 40 * r2 * (xx - W / 2) / W)[None].contiguous(), (yy + 40 * r2 * (yy - H / 2) / H)[None].contiguous()
    pair("remap (shared radial map)", "KB200_REMAP_V2", lambda: K.remap(x, mx, my, align_corners=True), 8 + 8 / (3 * B))
    mxb, myb = mx.expand(B, H, W).contiguous(), my.expand(B, H, W).contiguous()
    pair("remap (per-sample maps)", "KB200_REMAP_V2", lambda: K.remap(x, mxb, myb, align_corners=True), 8 + 8 / 3)
    pair("undistort_image (maps + remap v2)", "KB200_REMAP_V2", lambda: K.geometry.calibration.undistort_image(x, cam, dist), 8)
    pair("undistort_image (5 coefficients)", "KB200_FUSED_UNDISTORT", lambda: K.geometry.calibration.undistort_image(x, cam, dist), 8)

def ingest():
    """uint8 ingest warp against the three steps it replaces, on the headline homographies."""
    import bench

    KT = K.geometry.transform
    frames = torch.randint(0, 256, (B, H, W, 3), device=dev, dtype=torch.uint8)
    M = bench.make_homographies(B, 1000).to(dev)
    steps = lambda: KT.warp_perspective((frames.permute(0, 3, 1, 2).float() / 255.0), M, (H, W))  # noqa: E731
    fused = lambda: KT.warp_perspective_from_uint8(frames, M, (H, W))  # noqa: E731
    try:
        ref = steps()
        t0 = timed(steps)
        got = fused()
        same = torch.equal(got, ref)
        t1 = timed(fused)
        os.environ["KB200_U8_SIMPLE"] = "1"  # the per-tap kernel (warp_fwd_u8hwc) instead of the tiled one
        same_simple = torch.equal(fused(), ref)
        t2 = timed(fused)
    except Exception as exc:
        print(f"{'warp_perspective_from_uint8':34s} FAILED: {exc}")
        return
    finally:
        os.environ.pop("KB200_U8_SIMPLE", None)
    gb = B * H * W * 15 / 1e9  # 3 bytes read + 12 written per pixel
    print(f"{'warp_perspective_from_uint8':34s} permute+float+/255+warp {t0:7.3f} ms | tiled kernel {t1:7.3f} ms ({gb / t1 / peak * 1e3 * 100:5.1f} % HBM "
          f"of 15 B/pixel) x{t0 / t1:4.2f} {'bit-identical' if same else 'MISMATCH'} | per-tap kernel {t2:7.3f} ms {'bit-identical' if same_simple else 'MISMATCH'}")


def undistort_from_bytes():
    KC = K.geometry.calibration
    frames = torch.randint(0, 256, (B, H, W, 3), device=dev, dtype=torch.uint8)
    steps = lambda: KC.undistort_image((frames.permute(0, 3, 1, 2).float() / 255.0), cam, dist)  # noqa: E731
    fused = lambda: KC.undistort_image_from_uint8(frames, cam, dist)  # noqa: E731
    try:
        ref = steps()
        t0 = timed(steps, 3)
        got = fused()
        err = float((got - ref).norm() / ref.norm())
        t1 = timed(fused)
    except Exception as exc:
        print(f"{'undistort_image_from_uint8':34s} FAILED: {exc}")
        return
    gb = B * H * W * 15 / 1e9
    print(f"{'undistort_image_from_uint8':34s} permute+float+/255+maps+remap {t0:7.3f} ms | one kernel {t1:7.3f} ms ({gb / t1 / peak * 1e3 * 100:5.1f} % HBM of 15 B/pixel) "
          f"x{t0 / t1:4.2f} | rel-L2 vs the composition {err:.1e}")


with torch.no_grad():
    ingest()
    undistort_from_bytes()


def blur_backward():
    xx = x.detach().requires_grad_(True)
    (g,) = torch.autograd.grad(K.gaussian_blur2d(xx, (11, 11), (2.0, 2.0)), [xx], y)
    return g


os.environ.pop("KB200_FAST_FILTER_BWD", None)
ref = blur_backward()
t0 = timed(blur_backward, 3)
os.environ["KB200_FAST_FILTER_BWD"] = "1"
got = blur_backward()
t1 = timed(blur_backward, 3)
os.environ.pop("KB200_FAST_FILTER_BWD", None)
err = float((got - ref).norm() / ref.norm())
print(f"{'gaussian_blur2d k=11 fwd + d/dinput':34s} default {t0:7.3f} ms | KB200_FAST_FILTER_BWD=1 {t1:7.3f} ms | x{t0 / t1:4.2f} | rel-L2 vs default {err:.1e}")
del ref, got
with torch.no_grad():
    sweep("gaussian_blur2d k=11 (cfg3)", "KB200_SEP_VWALK", lambda: K.gaussian_blur2d(x, (11, 11), (2.0, 2.0)))
    sweep("ssim window 11", "KB200_SSIM_VWALK", lambda: K.metrics.ssim(x, y, 11))
    sweep("remap (per-sample maps)", "KB200_REMAP_V2", lambda: K.remap(x, mxb, myb, align_corners=True))
