"""Tiny driver for ncu captures: a few warp_perspective launches at 1080p with a small batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kornia_b200 as K
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
src = torch.rand(B, 3, 1080, 1920, device="cuda")
M = bench.make_homographies(B, 1000).cuda()
for _ in range(n):
    out = K.warp_perspective(src, M, (1080, 1920))
torch.cuda.synchronize()
print("done", float(out[0, 0, 0, 0]))
