"""GPU-box diagnostic: where do kernel and oracle differ?  (prelude vs sampler vs device)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import kornia_b200 as K
from kornia_b200.geometry import _prelude as P
from oracle import kornia_restated as R

torch.manual_seed(0)
g = torch.Generator().manual_seed(0)
B, H, W, h, w = 2, 23, 31, 19, 27
src = torch.rand(B, 3, H, W, generator=g)
M = torch.eye(3)[None].repeat(B, 1, 1) + 0.02 * torch.randn(B, 3, 3, generator=g)
M[:, 2, :2] *= 0.01
for dt in (torch.float32, torch.float64):
    s, m_ = src.to(dt), M.to(dt)
    mc = P.inverse3x3(P.normalize_homography(m_, (H, W), (h, w)))
    mg = P.inverse3x3(P.normalize_homography(m_.cuda(), (H, W), (h, w)))
    print(dt, "prelude cpu-vs-cuda max abs", float((mc - mg.cpu()).abs().max()))
    n_src = P.pixel_to_norm(H, W, m_)
    a = (m_ @ P.inverse3x3(n_src)); b = (m_.cuda() @ P.inverse3x3(n_src.cuda()))
    print("   matmul cpu-vs-cuda", float((a - b.cpu()).abs().max()), " inv ", float((P.inverse3x3(n_src) - P.inverse3x3(n_src.cuda()).cpu()).abs().max()))
    xs, ys = P.meshgrid_axes(h, w, "cuda", dt)
    xc, yc = P.meshgrid_axes(h, w, "cpu", dt)
    print("   base grid cpu-vs-cuda", float((xs.cpu() - xc).abs().max()))
    grid_g = R.perspective_grid(mg, xs, ys)
    grid_c = R.perspective_grid(mc, xc, yc)
    print("   grid cpu-vs-cuda", float((grid_g.cpu() - grid_c).abs().max()))
    ref_g = F.grid_sample(s.cuda(), grid_g, align_corners=True)
    ref_c = F.grid_sample(s, grid_c, align_corners=True)
    torch.backends.cudnn.enabled = False
    ref_g_nocudnn = F.grid_sample(s.cuda(), grid_g, align_corners=True)
    torch.backends.cudnn.enabled = True
    ours = K.warp_perspective(s.cuda(), m_.cuda(), (h, w))
    print("   ours vs torch-cuda(cudnn)", float((ours - ref_g).abs().max()), " vs torch-cuda(aten)", float((ours - ref_g_nocudnn).abs().max()),
          " vs cpu", float((ours.cpu() - ref_c).abs().max()), " torch cuda vs cpu", float((ref_g.cpu() - ref_c).abs().max()))
    # feed the CPU-computed matrices to our kernel: isolates the sampler
    from kornia_b200._ops import WarpFunction
    ours_cpu_m = WarpFunction.apply(s.cuda(), mc.cuda(), xc.cuda(), yc.cuda(), None, h, w, True, 0, 0, True)
    print("   ours (cpu-made m) vs cpu", float((ours_cpu_m.cpu() - ref_c).abs().max()))
print("matmul precision flags:", torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
