"""Golden vectors for the remaining callers of SURVEY.md 8f (rows 2-4): the pyramid functions (pyrdown, pyrup,
build_pyramid, build_laplacian_pyramid), the resize family (resize, rescale, resize_to_be_divisible) and the lens
functions (distort_points, undistort_image) -- recorded by running the UNMODIFIED reference on CPU fp32.  Build
container only (needs /root/reference):

    python tests/golden/make_golden_wider.py        ->  tests/golden/wider.npz

Same layout as family.npz (tensors by keyword, other keyword arguments as JSON); functions that return a list
store ``out0``, ``out1``, ...; ``*_grad`` cases store a cotangent and the reference's autograd gradients of
sum(out * cot).
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Bag, import_reference, smooth_image  # noqa: E402


def main():
    import_reference()
    import kornia.geometry.calibration as KC
    import kornia.geometry.calibration.distort  # noqa: F401
    import kornia.geometry.transform as KT

    torch.manual_seed(0)
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(777)
    bag = Bag()

    def tup(kw):
        return {k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items()}

    def fwd(name, fn, op, tensors, kw):
        out = fn(**tensors, **tup(kw))
        outs = {f"out{i}": o for i, o in enumerate(out)} if isinstance(out, (list, tuple)) else dict(out=out)
        bag.add(name, op, tensors, kw, outs)

    def grad(name, fn, op, tensors, kw, wrt):
        leaves = {k: (v.clone().requires_grad_(True) if k in wrt else v) for k, v in tensors.items()}
        out = fn(**leaves, **tup(kw))
        cot = torch.rand(out.shape, generator=gen) - 0.5
        g = torch.autograd.grad((out * cot).sum(), [leaves[k] for k in wrt])
        outs = {"out": out.detach(), "cot": cot}
        outs.update({f"grad_{k}": gi for k, gi in zip(wrt, g)})
        bag.add(name, op + "_grad", tensors, kw, outs)

    noise = torch.rand(2, 3, 18, 26, generator=gen)
    odd = torch.rand(1, 2, 17, 23, generator=gen)
    smooth = smooth_image(2, 3, 32, 48, gen)
    pow2 = smooth_image(1, 2, 32, 64, gen)
    wide = torch.rand(1, 1, 12, 140, generator=gen)  # crosses one 128-wide tile of the 5x5 kernel

    # ------------------------------------------------------------------ pyramids
    for border in ("reflect", "replicate", "constant", "circular"):
        for ac in (False, True):
            fwd(f"pyrdown_{border}_{int(ac)}", KT.pyrdown, "pyrdown", dict(input=noise), dict(border_type=border, align_corners=ac))
            fwd(f"pyrup_{border}_{int(ac)}", KT.pyrup, "pyrup", dict(input=odd), dict(border_type=border, align_corners=ac))
    fwd("pyrdown_odd", KT.pyrdown, "pyrdown", dict(input=odd), {})
    fwd("pyrdown_wide", KT.pyrdown, "pyrdown", dict(input=wide), {})
    fwd("pyrdown_factor3", KT.pyrdown, "pyrdown", dict(input=smooth), dict(factor=3.0))
    fwd("pyrdown_factor1p5", KT.pyrdown, "pyrdown", dict(input=odd), dict(factor=1.5))
    grad("pyrdown_grad", KT.pyrdown, "pyrdown", dict(input=smooth), {}, ["input"])
    grad("pyrup_grad", KT.pyrup, "pyrup", dict(input=noise), dict(border_type="replicate"), ["input"])
    fwd("build_pyramid_3", KT.build_pyramid, "build_pyramid", dict(input=smooth), dict(max_level=3))
    fwd("build_pyramid_1", KT.build_pyramid, "build_pyramid", dict(input=odd), dict(max_level=1))
    fwd("build_pyramid_ac", KT.build_pyramid, "build_pyramid", dict(input=noise), dict(max_level=2, border_type="replicate", align_corners=True))
    fwd("laplacian_pyr_pow2", KT.build_laplacian_pyramid, "build_laplacian_pyramid", dict(input=pow2), dict(max_level=3))
    fwd("laplacian_pyr_padded", KT.build_laplacian_pyramid, "build_laplacian_pyramid", dict(input=noise), dict(max_level=3))
    fwd("laplacian_pyr_halfpow2", KT.build_laplacian_pyramid, "build_laplacian_pyramid", dict(input=smooth), dict(max_level=2))

    # ------------------------------------------------------------------ resize family
    for interp in ("bilinear", "nearest", "bicubic", "area"):
        fwd(f"resize_{interp}_down", KT.resize, "resize", dict(input=smooth), dict(size=[13, 20], interpolation=interp))
        fwd(f"resize_{interp}_up", KT.resize, "resize", dict(input=odd), dict(size=[30, 41], interpolation=interp))
    for ac in (True, False):
        fwd(f"resize_ac{int(ac)}", KT.resize, "resize", dict(input=noise), dict(size=[9, 40], align_corners=ac))
    for side in ("short", "long", "vert", "horz"):
        fwd(f"resize_side_{side}", KT.resize, "resize", dict(input=noise), dict(size=12, side=side))
        fwd(f"resize_side_{side}_tall", KT.resize, "resize", dict(input=noise.transpose(-1, -2).contiguous()), dict(size=12, side=side))
    fwd("resize_aa_down", KT.resize, "resize", dict(input=smooth), dict(size=[11, 16], antialias=True))
    fwd("resize_aa_down_x_only", KT.resize, "resize", dict(input=smooth), dict(size=[40, 12], antialias=True))
    fwd("resize_aa_up_noop", KT.resize, "resize", dict(input=odd), dict(size=[34, 46], antialias=True))
    fwd("resize_aa_big_factor", KT.resize, "resize", dict(input=smooth), dict(size=[4, 5], antialias=True, interpolation="bicubic", align_corners=True))
    fwd("resize_hw", KT.resize, "resize", dict(input=noise[0, 0]), dict(size=[7, 9]))
    fwd("resize_chw", KT.resize, "resize", dict(input=noise[0]), dict(size=[7, 9], antialias=True))
    fwd("resize_5d", KT.resize, "resize", dict(input=torch.rand(2, 2, 3, 10, 12, generator=gen)), dict(size=[5, 6], antialias=True))
    grad("resize_aa_grad", KT.resize, "resize", dict(input=smooth), dict(size=[11, 16], antialias=True), ["input"])
    fwd("rescale_float", KT.rescale, "rescale", dict(input=noise), dict(factor=0.5, antialias=True))
    fwd("rescale_pair", KT.rescale, "rescale", dict(input=noise), dict(factor=[2.0, 0.4], antialias=True))
    fwd("rescale_up", KT.rescale, "rescale", dict(input=odd), dict(factor=1.7, interpolation="bicubic"))
    fwd("divisible_8", KT.resize_to_be_divisible, "resize_to_be_divisible", dict(input=noise), dict(divisible_factor=8))
    fwd("divisible_5_chw", KT.resize_to_be_divisible, "resize_to_be_divisible", dict(input=odd[0]), dict(divisible_factor=5, antialias=True))

    # ------------------------------------------------------------------ lens model
    img = smooth_image(2, 3, 30, 40, gen)
    cam = torch.tensor([[[38.0, 0.0, 19.5], [0.0, 36.0, 14.5], [0.0, 0.0, 1.0]], [[45.0, 0.0, 21.0], [0.0, 44.0, 13.0], [0.0, 0.0, 1.0]]])
    coefs = {
        4: torch.tensor([[-0.25, 0.08, 0.002, -0.003], [0.15, -0.05, -0.004, 0.001]]),
        5: torch.tensor([[-0.25, 0.08, 0.002, -0.003, 0.01], [0.15, -0.05, -0.004, 0.001, -0.02]]),
        8: torch.tensor([[-0.25, 0.08, 0.002, -0.003, 0.01, 0.05, -0.02, 0.004], [0.15, -0.05, -0.004, 0.001, -0.02, -0.03, 0.01, 0.002]]),
        12: torch.tensor([[-0.25, 0.08, 0.002, -0.003, 0.01, 0.05, -0.02, 0.004, 0.003, -0.001, 0.002, 0.0015],
                          [0.15, -0.05, -0.004, 0.001, -0.02, -0.03, 0.01, 0.002, -0.002, 0.001, -0.0015, 0.001]]),
        14: torch.tensor([[-0.25, 0.08, 0.002, -0.003, 0.01, 0.05, -0.02, 0.004, 0.003, -0.001, 0.002, 0.0015, 0.02, -0.015],
                          [0.15, -0.05, -0.004, 0.001, -0.02, -0.03, 0.01, 0.002, -0.002, 0.001, -0.0015, 0.001, -0.01, 0.03]]),
    }
    pts = torch.rand(2, 11, 2, generator=gen) * torch.tensor([39.0, 29.0])
    for n, d in coefs.items():
        fwd(f"undistort_{n}", KC.undistort_image, "undistort_image", dict(image=img, K=cam, dist=d), {})
        fwd(f"distort_points_{n}", KC.distort_points, "distort_points", dict(points=pts, K=cam, dist=d), {})
    fwd("undistort_unbatched_K", KC.undistort_image, "undistort_image", dict(image=img[:1], K=cam[0], dist=coefs[5][0]), {})
    fwd("undistort_chw", KC.undistort_image, "undistort_image", dict(image=img[1], K=cam[1], dist=coefs[8][1]), {})
    fwd("undistort_5d", KC.undistort_image, "undistort_image",
        dict(image=img.reshape(2, 1, 3, 30, 40), K=cam.reshape(2, 1, 3, 3), dist=coefs[4].reshape(2, 1, 4)), {})
    fwd("distort_points_newK", KC.distort_points, "distort_points", dict(points=pts, K=cam, dist=coefs[14], new_K=cam.flip(0).contiguous()), {})
    fwd("distort_points_shared", KC.distort_points, "distort_points", dict(points=pts[0], K=cam, dist=coefs[5]), {})
    grad("undistort_grad", KC.undistort_image, "undistort_image", dict(image=img, K=cam, dist=coefs[8]), {}, ["image", "K", "dist"])
    grad("distort_points_grad", KC.distort_points, "distort_points", dict(points=pts, K=cam, dist=coefs[14]), {}, ["points", "K", "dist"])

    tau = dict(taux=torch.tensor([0.02, -0.3, 0.0]), tauy=torch.tensor([-0.015, 0.2, 0.0]))
    fwd("tilt_forward", KC.tilt_projection, "tilt_projection", tau, dict(return_inverse=False))
    fwd("tilt_inverse", KC.tilt_projection, "tilt_projection", tau, dict(return_inverse=True))

    bag.save(os.path.join(HERE, "wider.npz"))


if __name__ == "__main__":
    main()
