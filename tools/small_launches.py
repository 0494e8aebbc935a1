"""Kernel list of one small-image call (32x3x256x256, the reference's published operating point): run under
ncu --metrics gpu__time_duration.sum to see where the device time of a call goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kornia_b200 as K

x = torch.rand(32, 3, 256, 256, device="cuda")
M = bench.make_homographies(32, 5).cuda()
A = M[:, :2].contiguous()
ang = torch.linspace(-30, 30, 32, device="cuda")
for _ in range(3):
    K.warp_perspective(x, M, (256, 256))
    K.warp_affine(x, A, (256, 256))
    K.geometry.transform.rotate(x, ang)
torch.cuda.synchronize()
