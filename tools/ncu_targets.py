"""One workload per kernel for `ncu --set full -k regex:<kernel>` captures (B=16 x 3 x 1080 x 1920, three calls):
    ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 2 -c 1 -o out python tools/ncu_targets.py remap
targets (optional: batch, homography seed): remap | remap_reflection | filter2d | ssim | grad | bicubic | reflection | fill | blur11 | blur17 | ingest"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import kornia_b200 as K  # noqa: E402

which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
x = torch.rand(B, 3, 1080, 1920, device="cuda")
M = bench.make_homographies(B, int(sys.argv[3]) if len(sys.argv) > 3 else 3).cuda()
if which in ("remap_reflection", "remap_border"):
    ys, xs = torch.meshgrid(torch.arange(1080, dtype=torch.float32, device="cuda"), torch.arange(1920, dtype=torch.float32, device="cuda"), indexing="ij")
    amp = torch.linspace(0.5, 4.0, B, device="cuda")[:, None, None]
    mx = (xs[None] + amp * torch.sin(ys / 40.0)[None]).contiguous()
    my = (ys[None] + amp * torch.cos(xs / 55.0)[None]).contiguous()
    f = lambda: K.remap(x, mx, my, padding_mode=which[6:], align_corners=True)  # noqa: E731
elif which == "remap":
    ys, xs = torch.meshgrid(torch.arange(1080, dtype=torch.float32, device="cuda"), torch.arange(1920, dtype=torch.float32, device="cuda"), indexing="ij")
    r2 = ((xs - 960) / 960) ** 2 + ((ys - 540) / 540) ** 2
    mx = (960 + (xs - 960) * (1 + 0.02 * r2))[None].expand(B, -1, -1).contiguous()
    my = (540 + (ys - 540) * (1 + 0.02 * r2))[None].expand(B, -1, -1).contiguous()
    f = lambda: K.remap(x, mx, my, align_corners=True)  # noqa: E731
elif which == "filter2d":
    k = torch.randn(1, 7, 7, device="cuda")
    f = lambda: K.filter2d(x, k)  # noqa: E731
elif which == "ssim":
    y = x.flip(-1).contiguous()
    f = lambda: K.metrics.ssim(x, y, 11)  # noqa: E731
elif which == "grad":
    f = lambda: K.filters.spatial_gradient(x, "sobel", 1)  # noqa: E731
elif which == "blur11":
    f = lambda: K.gaussian_blur2d(x, (11, 11), (2.0, 2.0))  # noqa: E731
elif which == "blur17":
    f = lambda: K.gaussian_blur2d(x, (17, 17), (3.0, 3.0))  # noqa: E731
elif which == "ingest":
    frames = torch.randint(0, 256, (B, 1080, 1920, 3), device="cuda", dtype=torch.uint8)
    f = lambda: K.geometry.transform.warp_perspective_from_uint8(frames, M, (1080, 1920))  # noqa: E731
elif which in ("bicubic",):
    f = lambda: K.warp_perspective(x, M, (1080, 1920), mode="bicubic")  # noqa: E731
else:  # reflection | fill (bilinear)
    f = lambda: K.warp_perspective(x, M, (1080, 1920), padding_mode=which, fill_value=torch.tensor([0.1, 0.5, 0.9]))  # noqa: E731
for _ in range(3):
    f()
torch.cuda.synchronize()
