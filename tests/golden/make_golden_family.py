"""Golden vectors for the callers either side of the hot path: the filter
family (box_blur, laplacian, unsharp_mask, spatial_gradient, sobel), the affine family (affine,
rotate, translate, scale, shear), the crops and the two matrix builders -- recorded by running the
UNMODIFIED reference on CPU fp32.  Build container only (needs /root/reference):

    python tests/golden/make_golden_family.py        ->  tests/golden/family.npz

Each case stores the tensors passed BY KEYWORD (names = the reference's parameter names), the other
keyword arguments as JSON, the output and -- for ``*_grad`` cases -- a cotangent plus the reference's
autograd gradients of sum(out * cot).
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
    import kornia.filters as KF
    import kornia.geometry.transform as KT

    torch.manual_seed(0)
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(4321)
    bag = Bag()

    def tup(kw):
        return {k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items()}

    def fwd(name, fn, op, tensors, kw):
        bag.add(name, op, tensors, kw, dict(out=fn(**tensors, **tup(kw))))

    def grad(name, fn, op, tensors, kw, wrt):
        leaves = {k: (v.clone().requires_grad_(True) if k in wrt else v) for k, v in tensors.items()}
        out = fn(**leaves, **tup(kw))
        cot = torch.rand(out.shape, generator=gen) - 0.5
        g = torch.autograd.grad((out * cot).sum(), [leaves[k] for k in wrt])
        outs = {"out": out.detach(), "cot": cot}
        outs.update({f"grad_{k}": gi for k, gi in zip(wrt, g)})
        bag.add(name, op + "_grad", tensors, kw, outs)

    noise = torch.rand(2, 3, 17, 23, generator=gen)
    smooth = smooth_image(2, 3, 24, 36, gen)
    thin = torch.rand(1, 2, 1, 9, generator=gen)
    tiny = torch.rand(3, 1, 2, 2, generator=gen)
    wide = torch.rand(1, 1, 9, 132, generator=gen)  # crosses one 128-wide block, W % 4 == 0

    # ------------------------------------------------------------------ filter family
    for border in ("reflect", "replicate", "constant", "circular"):
        for ks in (3, [3, 5], [5, 3], 7):
            for sep in (False, True):
                tag = f"{ks}".replace(" ", "").replace("[", "").replace("]", "").replace(",", "x")
                fwd(f"box_{border}_{tag}_{'sep' if sep else '2d'}", KF.box_blur, "box_blur", dict(input=noise),
                    dict(kernel_size=ks, border_type=border, separable=sep))
    grad("box_grad_2d", KF.box_blur, "box_blur", dict(input=smooth), dict(kernel_size=[3, 5], border_type="reflect", separable=False),
         ["input"])
    grad("box_grad_sep", KF.box_blur, "box_blur", dict(input=smooth), dict(kernel_size=5, border_type="replicate", separable=True),
         ["input"])

    for border in ("reflect", "replicate", "constant", "circular"):
        for ks in (3, 5, [3, 7]):
            for norm in (True, False):
                tag = f"{ks}".replace(" ", "").replace("[", "").replace("]", "").replace(",", "x")
                fwd(f"lap_{border}_{tag}_{int(norm)}", KF.laplacian, "laplacian", dict(input=noise),
                    dict(kernel_size=ks, border_type=border, normalized=norm))
    grad("lap_grad", KF.laplacian, "laplacian", dict(input=smooth), dict(kernel_size=5, border_type="reflect", normalized=True),
         ["input"])

    for border in ("reflect", "constant"):
        fwd(f"unsharp_{border}", KF.unsharp_mask, "unsharp_mask", dict(input=noise),
            dict(kernel_size=[3, 3], sigma=[1.5, 1.5], border_type=border))
    fwd("unsharp_5x7", KF.unsharp_mask, "unsharp_mask", dict(input=smooth), dict(kernel_size=[5, 7], sigma=[0.8, 2.0], border_type="replicate"))
    fwd("unsharp_sigma_tensor", KF.unsharp_mask, "unsharp_mask", dict(input=noise, sigma=torch.tensor([[1.0, 2.0], [0.5, 0.7]])),
        dict(kernel_size=[5, 5], border_type="reflect"))
    grad("unsharp_grad", KF.unsharp_mask, "unsharp_mask", dict(input=smooth), dict(kernel_size=[5, 5], sigma=[1.2, 1.2], border_type="reflect"),
         ["input"])

    for img_name, img in (("noise", noise), ("smooth", smooth), ("thin", thin), ("tiny", tiny), ("wide", wide)):
        for mode in ("sobel", "diff"):
            for order in (1, 2):
                for norm in (True, False):
                    fwd(f"sg_{img_name}_{mode}_{order}_{int(norm)}", KF.spatial_gradient, "spatial_gradient", dict(input=img),
                        dict(mode=mode, order=order, normalized=norm))
        for norm in (True, False):
            fwd(f"sobel_{img_name}_{int(norm)}", KF.sobel, "sobel", dict(input=img), dict(normalized=norm, eps=1e-6))
    fwd("sobel_eps", KF.sobel, "sobel", dict(input=noise), dict(normalized=True, eps=1e-3))
    for mode in ("sobel", "diff"):
        for order in (1, 2):
            grad(f"sg_grad_{mode}_{order}", KF.spatial_gradient, "spatial_gradient", dict(input=smooth),
                 dict(mode=mode, order=order, normalized=True), ["input"])
    grad("sg_grad_tiny", KF.spatial_gradient, "spatial_gradient", dict(input=tiny), dict(mode="sobel", order=2, normalized=False), ["input"])
    grad("sg_grad_thin", KF.spatial_gradient, "spatial_gradient", dict(input=thin), dict(mode="sobel", order=1, normalized=True), ["input"])
    grad("sobel_grad", KF.sobel, "sobel", dict(input=smooth), dict(normalized=True, eps=1e-6), ["input"])

    # ------------------------------------------------------------------ matrix builders
    centers = torch.tensor([[11.0, 8.0], [3.5, 20.25], [0.0, 0.0]])
    angles = torch.tensor([30.0, -112.5, 90.0])
    scales = torch.tensor([[1.0, 1.0], [0.7, 1.4], [2.0, 0.5]])
    fwd("rotmat", KT.get_rotation_matrix2d, "get_rotation_matrix2d", dict(center=centers, angle=angles, scale=scales), {})
    grad("rotmat_grad", KT.get_rotation_matrix2d, "get_rotation_matrix2d", dict(center=centers, angle=angles, scale=scales), {},
         ["center", "angle", "scale"])
    quad = torch.tensor([[0.0, 0.0], [35.0, 0.0], [35.0, 23.0], [0.0, 23.0]]).expand(4, 4, 2)
    quad_to = quad + 3.0 * torch.randn(4, 4, 2, generator=gen)
    fwd("persp_pts", KT.get_perspective_transform, "get_perspective_transform", dict(points_src=quad.contiguous(), points_dst=quad_to), {})
    fwd("persp_pts_rev", KT.get_perspective_transform, "get_perspective_transform", dict(points_src=quad_to, points_dst=quad.contiguous()), {})
    grad("persp_pts_grad", KT.get_perspective_transform, "get_perspective_transform", dict(points_src=quad.contiguous(), points_dst=quad_to),
         {}, ["points_src", "points_dst"])

    # ------------------------------------------------------------------ affine family
    two_angles = torch.tensor([30.0, -75.0])
    for mode in ("bilinear", "nearest"):
        for pad in ("zeros", "border", "reflection"):
            for ac in (True, False):
                fwd(f"rotate_{mode}_{pad}_{int(ac)}", KT.rotate, "rotate", dict(tensor=smooth, angle=two_angles),
                    dict(mode=mode, padding_mode=pad, align_corners=ac))
    fwd("rotate_center", KT.rotate, "rotate", dict(tensor=noise, angle=two_angles, center=torch.tensor([[4.0, 6.0], [12.0, 3.0]])), {})
    fwd("rotate_90", KT.rotate, "rotate", dict(tensor=torch.rand(1, 3, 4, 4, generator=gen), angle=torch.tensor([90.0])), {})
    fwd("rotate_bicubic", KT.rotate, "rotate", dict(tensor=smooth, angle=two_angles), dict(mode="bicubic"))
    grad("rotate_grad", KT.rotate, "rotate", dict(tensor=smooth, angle=two_angles), {}, ["tensor", "angle"])
    for ac in (True, False):
        fwd(f"translate_{int(ac)}", KT.translate, "translate", dict(tensor=smooth, translation=torch.tensor([[1.0, 0.0], [-3.25, 2.5]])),
            dict(align_corners=ac))
        fwd(f"scale_{int(ac)}", KT.scale, "scale", dict(tensor=smooth, scale_factor=torch.tensor([[2.0, 2.0], [0.6, 1.3]])),
            dict(align_corners=ac))
        fwd(f"shear_{int(ac)}", KT.shear, "shear", dict(tensor=smooth, shear=torch.tensor([[0.5, 0.0], [-0.2, 0.3]])), dict(align_corners=ac))
    fwd("scale_iso_center", KT.scale, "scale", dict(tensor=noise, scale_factor=torch.tensor([1.5]), center=torch.tensor([[5.0, 5.0]])), {})
    grad("translate_grad", KT.translate, "translate", dict(tensor=smooth, translation=torch.tensor([[1.5, -0.5], [-3.25, 2.5]])), {},
         ["tensor", "translation"])
    grad("scale_grad", KT.scale, "scale", dict(tensor=smooth, scale_factor=torch.tensor([[1.2, 0.9], [0.6, 1.3]])), {}, ["tensor", "scale_factor"])
    mats = KT.get_rotation_matrix2d(torch.tensor([[17.5, 11.5]]).expand(2, 2), two_angles, torch.tensor([[1.1, 0.9]]).expand(2, 2))
    fwd("affine_batch", KT.affine, "affine", dict(tensor=smooth, matrix=mats), {})
    fwd("affine_one_image", KT.affine, "affine", dict(tensor=smooth[:1], matrix=mats), dict(padding_mode="border"))
    fwd("affine_chw", KT.affine, "affine", dict(tensor=smooth[0], matrix=mats[:1]), dict(align_corners=False))

    # ------------------------------------------------------------------ crops
    boxes = torch.tensor([[[1.0, 1.0], [14.0, 1.0], [14.0, 10.0], [1.0, 10.0]], [[4.0, 2.0], [30.0, 5.0], [28.0, 20.0], [3.0, 18.0]]])
    for mode in ("bilinear", "nearest"):
        for ac in (True, False):
            fwd(f"crop_resize_{mode}_{int(ac)}", KT.crop_and_resize, "crop_and_resize", dict(input_tensor=smooth, boxes=boxes),
                dict(size=[9, 13], mode=mode, align_corners=ac))
            fwd(f"center_crop_{mode}_{int(ac)}", KT.center_crop, "center_crop", dict(input_tensor=smooth), dict(size=[10, 18], mode=mode, align_corners=ac))
    fwd("center_crop_odd", KT.center_crop, "center_crop", dict(input_tensor=noise), dict(size=[5, 7]))
    dst = torch.tensor([[[0.0, 0.0], [11.0, 0.0], [11.0, 7.0], [0.0, 7.0]]]).expand(2, 4, 2).contiguous()
    fwd("crop_boxes", KT.crop_by_boxes, "crop_by_boxes", dict(input_tensor=smooth, src_box=boxes, dst_box=dst), {})
    fwd("crop_boxes_border", KT.crop_by_boxes, "crop_by_boxes", dict(input_tensor=smooth, src_box=boxes, dst_box=dst),
        dict(padding_mode="border", align_corners=False))
    hom = KT.get_perspective_transform(boxes, dst)
    for ac in (True, False):
        fwd(f"crop_mat3_{int(ac)}", KT.crop_by_transform_mat, "crop_by_transform_mat", dict(input_tensor=smooth, transform=hom),
            dict(out_size=[8, 12], align_corners=ac))
        fwd(f"crop_mat2_{int(ac)}", KT.crop_by_transform_mat, "crop_by_transform_mat", dict(input_tensor=smooth, transform=mats),
            dict(out_size=[8, 12], align_corners=ac))
    fwd("crop_mat3_shared_1px", KT.crop_by_transform_mat, "crop_by_transform_mat", dict(input_tensor=smooth, transform=hom[:1]),
        dict(out_size=[1, 12], align_corners=False))
    # corners off the pixel grid: bilinear sampling is not differentiable at integer coordinates (one-sided derivatives
    # differ), and the destination corners map exactly onto the source corners
    soft = torch.tensor([[[1.3, 1.6], [14.2, 1.4], [14.4, 10.7], [1.1, 10.3]], [[4.3, 2.6], [30.4, 5.2], [28.7, 20.4], [3.2, 18.6]]])
    grad("crop_resize_grad", KT.crop_and_resize, "crop_and_resize", dict(input_tensor=smooth, boxes=soft), dict(size=[9, 13]), ["input_tensor", "boxes"])

    bag.save(os.path.join(HERE, "family.npz"))


if __name__ == "__main__":
    main()
