"""ncu driver: a few launches of the blur kernel and of the warp backward at reduced batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kornia_b200 as K
what = sys.argv[1]
if what == "blur":
    x = torch.rand(16, 3, 1080, 1920, device="cuda")
    for _ in range(3):
        y = K.gaussian_blur2d(x, (11, 11), (2.0, 2.0))
else:
    B, H, W = 32, 720, 1280
    src = torch.rand(B, 3, H, W, device="cuda", requires_grad=True)
    g = torch.Generator().manual_seed(7)
    quad = torch.tensor([[0.0, 0.0], [W - 1.0, 0.0], [W - 1.0, H - 1.0], [0.0, H - 1.0]]).expand(B, 4, 2)
    M = bench.perspective_from_quads(quad, quad + 8.0 * torch.randn(B, 4, 2, generator=g)).cuda().requires_grad_(True)
    for _ in range(3):
        out = K.warp_perspective(src, M, (H, W))
        gs, gm = torch.autograd.grad(out.sum(), [src, M])
torch.cuda.synchronize()
print("done")
