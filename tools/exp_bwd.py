"""GPU experiment: where does the tiled backward spend its time?  (debug knobs skip parts; results are then wrong)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import kornia_b200 as K
dev = "cuda"
B, H, W = 128, 720, 1280
src = torch.rand(B, 3, H, W, device=dev, requires_grad=True)
g = torch.Generator().manual_seed(7)
quad = torch.tensor([[0.0, 0.0], [W - 1.0, 0.0], [W - 1.0, H - 1.0], [0.0, H - 1.0]]).expand(B, 4, 2)
M = bench.perspective_from_quads(quad, quad + 8.0 * torch.randn(B, 4, 2, generator=g)).to(dev).requires_grad_(True)
cot = torch.rand(B, 3, H, W, device=dev) - 0.5
def run(wrt):
    out = K.warp_perspective(src, M, (H, W))
    for _ in range(2): torch.autograd.grad(out, wrt, grad_outputs=cot, retain_graph=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): torch.autograd.grad(out, wrt, grad_outputs=cot, retain_graph=True)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / 5
for dbg, what in ((0, "full"), (1, "no TMA-reduce flush"), (3, "no strip adds, no flush"), (4, "no d/dM taps"), (7, "coords + gout only")):
    os.environ["KB200_BWD_DEBUG"] = str(dbg)
    print(f"debug={dbg} ({what}): d/dsrc+d/dM {run([src, M]):.3f} ms | d/dsrc only {run([src]):.3f} ms | d/dM only {run([M]):.3f} ms  (each incl. zero-fill + prelude bwd)", flush=True)
os.environ["KB200_BWD_DEBUG"] = "0"
x = torch.zeros_like(src)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): torch.zeros_like(src)
e.record(); torch.cuda.synchronize()
print(f"zero-fill alone: {s.elapsed_time(e)/5:.3f} ms")
