"""GPU experiment 3: L2 promotion x tile configs (with tail-balanced scheduling)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from kornia_b200 import _lib, _ops
from kornia_b200.geometry import _prelude as P
dev = "cuda"
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
for cfg in ("64x32x72x40x2x2x128", "64x32x72x40x2x2x256", "64x32x72x40x3x2x256", "128x16x136x24x2x2x256", "128x16x136x24x3x2x256",
            "128x32x136x40x2x1x256", "64x16x72x24x4x2x256", "64x16x72x24x2x2x256", "64x32x72x40x2x2x256"):
    t, out = run(cfg, False)
    if ref is None: ref = out.clone()
    eq = bool(torch.equal(out, ref))
    tc, _ = run(cfg, True)
    print(f"cfg {cfg:24s}: warp {t:.3f} ms ({24*B*1080*1920/t/1e6:5.0f} GB/s) equal={eq} | copy-only probe {tc:.3f} ms ({24*B*1080*1920/tc/1e6:5.0f} GB/s)", flush=True)
    del out
