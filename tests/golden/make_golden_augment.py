"""Golden vectors for the augmentation callers of the hot path: RandomPerspective, RandomAffine and
RandomGaussianBlur of the UNMODIFIED reference run on CPU fp32 with a pinned seed.  Build container only (needs
/root/reference):

    python tests/golden/make_golden_augment.py        ->  tests/golden/augment.npz

Per case: the constructor arguments (JSON), the seed, the input, every tensor of the parameter dictionary the reference
sampled (``param_<name>``: batch_prob included), the output and -- for the geometric classes -- ``transform_matrix``.  The
tests replay the recorded parameters through kornia_b200.augmentation on the GPU (output parity) and re-sample from the same
seed on the CPU generator (the parameter stream itself)."""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Bag, import_reference, smooth_image  # noqa: E402


def main():
    import_reference()
    import kornia.augmentation as KA

    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(4242)
    bag = Bag()
    noise = torch.rand(5, 3, 40, 56, generator=gen)
    smooth = smooth_image(4, 3, 64, 96, gen)
    single = torch.rand(3, 32, 48, generator=gen)  # (C,H,W): keepdim cases

    def record(name, cls, ctor, x, seed):
        aug = getattr(KA, cls)(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in ctor.items()})
        torch.manual_seed(seed)
        out = aug(x)
        outs = {"out": out}
        for k, v in aug._params.items():
            if isinstance(v, torch.Tensor):
                outs[f"param_{k}"] = v
        if getattr(aug, "transform_matrix", None) is not None and cls != "RandomGaussianBlur":
            outs["transform_matrix"] = aug.transform_matrix
        bag.add(name, cls, {"input": x}, {"ctor": ctor, "seed": seed}, outs)

    # ---- RandomPerspective
    record("persp_default", "RandomPerspective", dict(distortion_scale=0.5, p=1.0), noise, 1)
    record("persp_gated", "RandomPerspective", dict(distortion_scale=0.3, p=0.5), noise, 2)
    record("persp_area", "RandomPerspective", dict(distortion_scale=0.4, p=1.0, sampling_method="area_preserving", align_corners=True), smooth, 3)
    record("persp_nearest_same", "RandomPerspective", dict(distortion_scale=0.6, p=1.0, resample="NEAREST", same_on_batch=True), noise, 4)
    record("persp_bicubic", "RandomPerspective", dict(distortion_scale=0.2, p=0.7, resample="BICUBIC"), smooth, 5)
    record("persp_keepdim", "RandomPerspective", dict(distortion_scale=0.5, p=1.0, keepdim=True), single, 6)
    # ---- RandomAffine
    record("affine_rot", "RandomAffine", dict(degrees=30.0, p=1.0), noise, 11)
    record("affine_full", "RandomAffine", dict(degrees=[-20.0, 40.0], translate=[0.1, 0.2], scale=[0.8, 1.3], shear=[-10.0, 10.0, -5.0, 8.0], p=1.0), smooth, 12)
    record("affine_gated", "RandomAffine", dict(degrees=15.0, translate=[0.05, 0.05], scale=[0.9, 1.1, 0.7, 1.2], shear=8.0, p=0.5), noise, 13)
    record("affine_border_ac", "RandomAffine", dict(degrees=45.0, padding_mode="BORDER", align_corners=True, p=1.0), noise, 14)
    record("affine_reflection_nearest", "RandomAffine", dict(degrees=10.0, shear=[-15.0, 15.0], padding_mode="REFLECTION", resample="NEAREST", p=1.0), smooth, 15)
    record("affine_same_on_batch", "RandomAffine", dict(degrees=25.0, translate=[0.2, 0.1], same_on_batch=True, p=1.0), noise, 16)
    record("affine_keepdim", "RandomAffine", dict(degrees=30.0, scale=[0.7, 1.4], p=1.0, keepdim=True), single, 17)
    # ---- RandomGaussianBlur
    record("blur_default", "RandomGaussianBlur", dict(kernel_size=[5, 5], sigma=[0.1, 2.0], p=1.0), noise, 21)
    record("blur_gated", "RandomGaussianBlur", dict(kernel_size=[3, 7], sigma=[0.5, 1.5], p=0.5), noise, 22)
    record("blur_k11_replicate", "RandomGaussianBlur", dict(kernel_size=11, sigma=[1.0, 3.0], border_type="replicate", p=1.0), smooth, 23)
    record("blur_same_nonsep", "RandomGaussianBlur", dict(kernel_size=[5, 5], sigma=[0.3, 1.0], separable=False, same_on_batch=True, p=1.0), smooth, 24)
    record("blur_keepdim", "RandomGaussianBlur", dict(kernel_size=[3, 3], sigma=[0.2, 0.8], p=1.0, keepdim=True), single, 25)
    bag.save(os.path.join(HERE, "augment.npz"))


if __name__ == "__main__":
    main()
