"""Which samples of the bench batch make the 'reflection' warp slower than 'zeros'?  One sample per call (x8 copies to fill the GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kornia_b200 as K

B = 64
src = torch.rand(8, 3, 1080, 1920, device="cuda")
M = bench.make_homographies(B, 1000).cuda()


def t(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


rows = []
for b in range(B):
    Mb = M[b:b + 1].expand(8, -1, -1).contiguous()
    z = t(lambda: K.warp_perspective(src, Mb, (1080, 1920), padding_mode="zeros"))
    r = t(lambda: K.warp_perspective(src, Mb, (1080, 1920), padding_mode="reflection"))
    rows.append((r / z, b, z, r))
rows.sort(reverse=True)
for ratio, b, z, r in rows[:8]:
    print(f"sample {b:2d}: zeros {z:7.1f} us  reflection {r:7.1f} us  x{ratio:.2f}\n{M[b].cpu().numpy()}")
print("median ratio", sorted(x[0] for x in rows)[B // 2])
