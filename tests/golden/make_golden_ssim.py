"""Golden vectors for ``ssim`` / ``ssim_loss`` recorded from the UNMODIFIED reference on CPU fp32
(build container only):  python tests/golden/make_golden_ssim.py  ->  tests/golden/ssim.npz
Same layout as family.npz (tensors by keyword, gradients of sum(out * cot) for ``*_grad`` cases)."""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Bag, import_reference, smooth_image  # noqa: E402


def main():
    import_reference()
    import kornia.losses as KL
    import kornia.metrics as KM

    torch.manual_seed(0)
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(99)
    bag = Bag()

    def fwd(name, fn, op, tensors, kw):
        bag.add(name, op, tensors, kw, dict(out=fn(**tensors, **kw)))

    def grad(name, fn, op, tensors, kw, wrt):
        leaves = {k: (v.clone().requires_grad_(True) if k in wrt else v) for k, v in tensors.items()}
        out = fn(**leaves, **kw)
        cot = torch.rand(out.shape, generator=gen) - 0.5
        g = torch.autograd.grad((out * cot).sum(), [leaves[k] for k in wrt])
        outs = {"out": out.detach(), "cot": cot}
        outs.update({f"grad_{k}": gi for k, gi in zip(wrt, g)})
        bag.add(name, op + "_grad", tensors, kw, outs)

    a = torch.rand(2, 3, 37, 70, generator=gen)                     # several tiles, odd height, even width
    b = (a + 0.1 * torch.randn(2, 3, 37, 70, generator=gen)).clamp(0, 1)
    sa = smooth_image(1, 2, 40, 33, gen)                            # odd width: scalar stores
    sb = (sa + 0.05 * torch.randn(1, 2, 40, 33, generator=gen)).clamp(0, 1)
    for ws in (3, 5, 7, 9, 11):
        for pad in ("same", "valid"):
            fwd(f"ssim_noise_{ws}_{pad}", KM.ssim, "ssim", dict(img1=a, img2=b), dict(window_size=ws, padding=pad))
            fwd(f"ssim_smooth_{ws}_{pad}", KM.ssim, "ssim", dict(img1=sa, img2=sb), dict(window_size=ws, padding=pad))
    fwd("ssim_identical", KM.ssim, "ssim", dict(img1=a, img2=a.clone()), dict(window_size=11))
    fwd("ssim_maxval255", KM.ssim, "ssim", dict(img1=a * 255, img2=b * 255), dict(window_size=7, max_val=255.0))
    fwd("ssim_eps", KM.ssim, "ssim", dict(img1=sa, img2=sb), dict(window_size=5, eps=1e-3))
    fwd("ssim_window13", KM.ssim, "ssim", dict(img1=a, img2=b), dict(window_size=13))   # beyond the fused kernel: composed path
    tiny = torch.rand(1, 1, 6, 6, generator=gen)
    fwd("ssim_tiny", KM.ssim, "ssim", dict(img1=tiny, img2=tiny.flip(-1).contiguous()), dict(window_size=5))
    for red in ("mean", "sum", "none"):
        fwd(f"loss_{red}", KL.ssim_loss, "ssim_loss", dict(img1=a, img2=b), dict(window_size=5, reduction=red))
    fwd("loss_valid", KL.ssim_loss, "ssim_loss", dict(img1=sa, img2=sb), dict(window_size=11, padding="valid"))
    grad("ssim_grad", KM.ssim, "ssim", dict(img1=sa, img2=sb), dict(window_size=5), ["img1", "img2"])
    grad("loss_grad", KL.ssim_loss, "ssim_loss", dict(img1=a, img2=b), dict(window_size=7), ["img1"])
    bag.save(os.path.join(HERE, "ssim.npz"))


if __name__ == "__main__":
    main()
