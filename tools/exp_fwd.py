"""GPU experiment: (1) which fused-prelude variant is bit-identical to the torch prelude, (2) kernel time per TMA config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kornia_b200 as K
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
    for (H, W, h, w) in ((1080, 1920, 1080, 1920), (720, 1280, 360, 640), (37, 53, 29, 41), (1, 5, 3, 1)):
        Mb = bench.make_homographies(512, 3).to(dt)
        Mr = torch.eye(3)[None].repeat(512, 1, 1) + 0.2 * torch.randn(512, 3, 3, generator=g)
        Mr[:, 2, :2] *= 0.001
        for name, M in (("bench", Mb), ("rand", Mr.to(dt))):
            M = M.to(dev)
            want_p = P.inverse3x3(P.normalize_homography(M, (H, W), (h, w)))
            A = M[:, :2].contiguous()
            want_a = P.inverse3x3(P.normalize_homography(P.affine_to_homography(A), (H, W), (h, w)))
            res = []
            for v in range(4):
                gp = fused(M, 3, H, W, h, w, v)
                ga = fused(A, 2, H, W, h, w, v)
                bad_p = int((gp != want_p).sum() - ((gp != gp) & (want_p != want_p)).sum())
                bad_a = int((ga != want_a).sum() - ((ga != ga) & (want_a != want_a)).sum())
                res.append((bad_p, bad_a))
            print(f"prelude {str(dt)[6:]:8s} {H}x{W}->{h}x{w} {name:5s} mismatching elements (persp, affine) per variant 0..3: {res}", flush=True)

# ---- kernel timing per config
B = 256
src = torch.rand(B, 3, 1080, 1920, device=dev)
M = bench.make_homographies(B, 1000).to(dev)
m = P.inverse3x3(P.normalize_homography(M, (1080, 1920), (1080, 1920)))
bx, by = P.meshgrid_axes(1080, 1920, dev, torch.float32)
ref = None
for cfg in ("2x72", "3x72", "2x96", "3x96", "2x72"):
    os.environ["KB200_TMA_CFG"] = cfg
    for _ in range(3):
        out = _ops.WarpFunction.apply(src, m, bx, by, None, 1080, 1920, True, 0, 0, True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = _ops.WarpFunction.apply(src, m, bx, by, None, 1080, 1920, True, 0, 0, True); e.record()
        torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort()
    if ref is None: ref = out.clone()
    print(f"cfg {cfg}: median {ts[10]:.3f} ms  min {ts[0]:.3f} ms  -> {24*B*1080*1920/ts[10]/1e6:.0f} GB/s  equal_to_first={bool(torch.equal(out, ref))}", flush=True)
    del out
