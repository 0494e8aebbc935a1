"""GPU experiment 2: prelude det-order variants; TMA tile configs with and without the copy-only probe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from kornia_b200 import _lib, _ops
from kornia_b200.geometry import _prelude as P

dev = "cuda"
lib = _lib.load()

def fused(M, rows, H, W, h, w, variant):
    out = torch.empty(M.shape[0], 3, 3, device=dev, dtype=M.dtype)
    rc = lib.kb200_warp_prelude(M.data_ptr(), out.data_ptr(), M.shape[0], rows, H, W, h, w, 0 if M.dtype == torch.float32 else 1, variant, None)
    assert rc == 0, lib.kb200_last_error()
    return out

g = torch.Generator().manual_seed(0)
for dt in (torch.float32, torch.float64):
    for (H, W, h, w) in ((1080, 1920, 1080, 1920), (720, 1280, 360, 640), (37, 53, 29, 41)):
        for n in (512, 7, 1):
            Mr = torch.eye(3)[None].repeat(n, 1, 1) + 0.2 * torch.randn(n, 3, 3, generator=g)
            Mr[:, 2, :2] *= 0.001
            M = Mr.to(dt).to(dev)
            want = P.inverse3x3(P.normalize_homography(M, (H, W), (h, w)))
            res = {v: int((fused(M, 3, H, W, h, w, v) != want).sum()) for v in (0, 4, 8, 1, 5)}
            print(f"prelude {str(dt)[6:]:8s} {H}x{W}->{h}x{w} n={n} mismatches by variant: {res}", flush=True)

B = 256
src = torch.rand(B, 3, 1080, 1920, device=dev)
M = bench.make_homographies(B, 1000).to(dev)
m = P.inverse3x3(P.normalize_homography(M, (1080, 1920), (1080, 1920)))
bx, by = P.meshgrid_axes(1080, 1920, dev, torch.float32)

def run(cfg, copy):
    os.environ["KB200_TMA_CFG"] = cfg
    os.environ["KB200_TMA_COPYONLY"] = "1" if copy else "0"
    for _ in range(3):
        out = _ops.WarpFunction.apply(src, m, bx, by, None, 1080, 1920, True, 0, 0, True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = _ops.WarpFunction.apply(src, m, bx, by, None, 1080, 1920, True, 0, 0, True); e.record()
        torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[7], out

ref = None
for cfg in ("64x32x72x40x2x2", "64x32x72x40x2x2x256", "64x32x72x40x2x2x0", "64x32x72x40x3x2", "64x32x72x40x2x1", "128x16x136x24x2x2", "128x16x136x24x3x2",
            "128x32x136x40x2x1", "64x16x72x24x2x2", "64x16x72x24x4x2", "64x16x72x24x4x3", "32x32x40x40x3x2", "32x32x40x40x3x3"):
    t, out = run(cfg, False)
    if ref is None: ref = out.clone()
    eq = bool(torch.equal(out, ref))
    tc, _ = run(cfg, True)
    print(f"cfg {cfg:22s}: warp {t:.3f} ms ({24*B*1080*1920/t/1e6:5.0f} GB/s) equal={eq} | copy-only probe {tc:.3f} ms ({24*B*1080*1920/tc/1e6:5.0f} GB/s)", flush=True)
    del out
