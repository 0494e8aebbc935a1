"""The BASELINE configs against the reference's own torch composition ON THE SAME GPU ("time torch eager on the
same GPU -- the real bar to beat").  The reference cannot travel to the GPU box; oracle/kornia_restated.py issues exactly the
ATen calls the reference makes, so it stands in for it.  CUDA events, 3 warm-ups; B is reduced for the torch arm where its
temporaries (base grid + 4.25 GB sampling grid at B=256) would not leave room.  Not run in round 1 (no GPU time left):

    gpurun -- 'python tools/bench_vs_torch.py > gpurun_out/vs_torch.txt'
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import kornia_b200 as K  # noqa: E402
from oracle import kornia_restated as R  # noqa: E402

dev = "cuda"


def t(fn, n=10):
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


def report(name, ours_ms, torch_ms, B):
    print(f"{name:52s} B={B:4d}  ours {ours_ms:8.3f} ms   torch eager {torch_ms:8.3f} ms   x{torch_ms / ours_ms:.2f}", flush=True)


B = int(os.environ.get("VS_TORCH_B", "64"))
H, W = bench.H_IMG, bench.W_IMG
with torch.no_grad():
    x = torch.rand(B, 3, H, W, device=dev)
    M = bench.make_homographies(B, 1000).to(dev)
    report("cfg2 warp_perspective fwd 1080p bilinear/zeros", t(lambda: K.warp_perspective(x, M, (H, W))), t(lambda: R.warp_perspective(x, M, (H, W)), 3), B)
    report("cfg3 gaussian_blur2d k=11 sigma=2 reflect", t(lambda: K.gaussian_blur2d(x, (11, 11), (2.0, 2.0))),
           t(lambda: R.gaussian_blur2d(x, (11, 11), (2.0, 2.0)), 3), B)
    for mode in ("nearest", "bicubic"):
        report(f"warp_perspective fwd 1080p {mode}/zeros", t(lambda: K.warp_perspective(x, M, (H, W), mode=mode)),
               t(lambda: R.warp_perspective(x, M, (H, W), mode=mode), 3), B)
    del x
Bb = max(8, B // 2)
src = torch.rand(Bb, 3, 720, 1280, device=dev)
Mb = bench.make_homographies(Bb, 7).to(dev)
cot = torch.randn(Bb, 3, 720, 1280, device=dev)


def step(impl):
    s = src.detach().requires_grad_(True)
    m = Mb.detach().requires_grad_(True)
    impl.warp_perspective(s, m, (720, 1280)).backward(cot)


report("cfg4 warp_perspective fwd+bwd 720p (d/dsrc, d/dM)", t(lambda: step(K)), t(lambda: step(R), 3), Bb)
