"""Golden vectors for the uint8 ingest warps: the UNMODIFIED reference's three steps
``image_to_tensor`` (kornia/image/utils.py:27) -> ``_to_float32`` (kornia/io/io.py:108-111) -> ``warp_perspective`` /
``warp_affine`` (imgwarp.py:69,177) on CPU fp32, recorded against the interleaved uint8 input.  Build container only
(needs /root/reference):

    python tests/golden/make_golden_ingest.py        ->  tests/golden/ingest.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Bag, import_reference, jitter_homography  # noqa: E402


def main():
    kornia = import_reference()
    import kornia as K
    import kornia.geometry.transform as KT
    from kornia.io.io import _to_float32

    torch.manual_seed(0)
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(4242)
    bag = Bag()

    def frames(B, H, W, C):
        return torch.randint(0, 256, (B, H, W, C), generator=gen, dtype=torch.uint8)

    def reference(image, M, kw, affine):
        x = K.image.image_to_tensor(image.numpy(), keepdim=False)   # (B,H,W,C) -> (B,C,H,W); (H,W,C) -> (1,C,H,W)
        x = _to_float32(x) if kw.get("normalize", True) else x.float()
        fn = KT.warp_affine if affine else KT.warp_perspective
        args = {k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items() if k != "normalize"}
        return fn(x, M, **args)

    def add(name, image, M, kw, affine, fill=None):
        tensors = dict(image=image, M=M)
        kw2 = dict(kw)
        if fill is not None:
            tensors["fill_value"] = fill
        out = reference(image, M, dict(kw2, **({"fill_value": fill} if fill is not None else {})), affine)
        bag.add(name, "warp_affine_from_uint8" if affine else "warp_perspective_from_uint8", tensors, kw2, dict(out=out))

    B, H, W = 2, 37, 52
    img3 = frames(B, H, W, 3)
    Hm = jitter_homography(kornia, B, H, W, gen, 3.0)
    wild = Hm.clone()
    wild[1] = torch.tensor([[0.8, -0.5, 14.0], [0.45, 0.9, -9.0], [2e-3, -1e-3, 1.0]])  # leaves the image: exercises the padding
    for mode in ("bilinear", "nearest", "bicubic"):
        for pad in ("zeros", "border", "reflection"):
            for ac in (True, False):
                add(f"persp_{mode}_{pad}_{int(ac)}", img3, wild, dict(dsize=[H, W], mode=mode, padding_mode=pad, align_corners=ac), False)
        add(f"persp_{mode}_fill", img3, wild, dict(dsize=[H, W], mode=mode, padding_mode="fill", align_corners=True), False,
            fill=torch.tensor([0.25, 0.5, 0.75]))
    add("persp_resample", img3, Hm, dict(dsize=[29, 61], mode="bilinear", padding_mode="zeros", align_corners=True), False)
    add("persp_raw_bytes", img3, Hm, dict(dsize=[H, W], mode="bilinear", padding_mode="border", align_corners=True, normalize=False), False)
    add("persp_gray", frames(B, H, W, 1), Hm, dict(dsize=[H, W], mode="bilinear", padding_mode="zeros", align_corners=True), False)
    add("persp_rgba", frames(B, H, W, 4), Hm, dict(dsize=[H, W], mode="bicubic", padding_mode="reflection", align_corners=False), False)
    add("persp_single_hwc", img3[0], Hm[:1], dict(dsize=[H, W], mode="bilinear", padding_mode="zeros", align_corners=True), False)
    add("persp_all_levels", torch.arange(256, dtype=torch.uint8).reshape(1, 16, 16, 1), torch.eye(3)[None],
        dict(dsize=[16, 16], mode="nearest", padding_mode="zeros", align_corners=True), False)  # every byte value through / 255

    rot = KT.get_rotation_matrix2d(torch.tensor([[W / 2, H / 2]]).expand(B, 2), torch.tensor([30.0, -12.0]), torch.ones(B, 2))
    for mode in ("bilinear", "nearest", "bicubic"):
        for ac in (True, False):
            add(f"affine_{mode}_{int(ac)}", img3, rot, dict(dsize=[H, W], mode=mode, padding_mode="zeros", align_corners=ac), True)
    add("affine_shared", img3, rot[:1], dict(dsize=[41, 33], mode="bilinear", padding_mode="border", align_corners=True), True)
    add("affine_fill_scalar", img3, rot, dict(dsize=[H, W], mode="bilinear", padding_mode="fill", align_corners=True), True, fill=torch.tensor(0.5))
    add("affine_fill_gray", frames(B, H, W, 1), rot, dict(dsize=[H, W], mode="bilinear", padding_mode="fill", align_corners=False), True,
        fill=torch.tensor([0.3]))
    # ------------------------------------------------------------------ undistort_image on decoder bytes
    import kornia.geometry.calibration as KC

    def undist(name, image, cam, coef, normalize=True):
        x = K.image.image_to_tensor(image.numpy(), keepdim=False)
        x = _to_float32(x) if normalize else x.float()
        n = x.shape[0]
        out = KC.undistort_image(x, cam if cam.dim() == 3 else cam.expand(n, 3, 3), coef if coef.dim() == 2 else coef.expand(n, coef.shape[-1]))
        bag.add(name, "undistort_image_from_uint8", dict(image=image, K=cam, dist=coef), {} if normalize else dict(normalize=False), dict(out=out))

    Hu, Wu = 30, 40
    fr = frames(2, Hu, Wu, 3)
    cam = torch.tensor([[[34.0, 0.0, 19.5], [0.0, 33.0, 14.0], [0.0, 0.0, 1.0]], [[30.0, 0.0, 21.0], [0.0, 31.0, 15.5], [0.0, 0.0, 1.0]]])
    coefs = {4: torch.tensor([[-0.25, 0.08, 0.002, -0.003], [0.15, -0.05, -0.004, 0.001]]),
             5: torch.tensor([[-0.25, 0.08, 0.002, -0.003, 0.01], [0.15, -0.05, -0.004, 0.001, -0.02]]),
             8: torch.tensor([[-0.25, 0.08, 0.002, -0.003, 0.01, 0.04, -0.02, 0.004], [0.15, -0.05, -0.004, 0.001, -0.02, 0.03, 0.01, -0.002]]),
             12: torch.tensor([[-0.25, 0.08, 0.002, -0.003, 0.01, 0.04, -0.02, 0.004, 0.003, -0.001, 0.002, 0.0015],
                               [0.15, -0.05, -0.004, 0.001, -0.02, 0.03, 0.01, -0.002, -0.002, 0.001, 0.001, -0.001]])}
    coefs[14] = torch.cat([coefs[12], torch.tensor([[0.02, -0.015], [-0.01, 0.02]])], -1)   # tilt terms: the composition path
    for k, c in coefs.items():
        undist(f"undistort_u8_{k}", fr, cam, c)
    undist("undistort_u8_raw", fr, cam, coefs[5], normalize=False)
    undist("undistort_u8_gray", frames(2, Hu, Wu, 1), cam, coefs[8])
    undist("undistort_u8_single_hwc", fr[0], cam[:1], coefs[5][:1])
    undist("undistort_u8_shared_camera", fr, cam[0], coefs[5][0])
    undist("undistort_u8_odd_width", frames(1, 21, 37, 3), cam[:1], coefs[4][:1])    # W % 4 != 0: the composition path
    bag.save(os.path.join(HERE, "ingest.npz"))


if __name__ == "__main__":
    main()
