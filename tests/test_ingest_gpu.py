"""GPU parity of the uint8 ingest warps (csrc/warp_u8.cuh) through the C ABI.  The kernel was written
after the round-1 GPU budget was spent: it has run on the host emulator only, nothing else in the library calls it, and
these tests are skipped unless KB200_RUN_UNVERIFIED=1 (tools/r2_first_call.sh runs them):

    KB200_RUN_UNVERIFIED=1 python -m pytest tests/test_ingest_gpu.py -m gpu -q
"""
import pytest
import torch

import kornia_b200 as K
from conftest import golden
from helpers import run_family_case

pytestmark = [pytest.mark.gpu]
DEV = "cuda"
ING = golden("ingest")
KT, KC = K.geometry.transform, K.geometry.calibration
WARPS = [n for n in ING.names() if ING.meta[n]["op"].startswith("warp_")]
UNDISTORT = [n for n in ING.names() if ING.meta[n]["op"].startswith("undistort")]


@pytest.mark.parametrize("name", WARPS)
def test_ingest_matches_reference(name):
    """Golden vectors of image_to_tensor + _to_float32 + warp_* recorded from the reference on CPU: 1e-4 rel (north_star)."""
    op, kw, ins, outs = ING.case(name)
    before = K._ops.launch_count
    got = run_family_case(KT, op, kw, ins, device=DEV)
    assert K._ops.launch_count == before + 1 + (1 if ins["M"].shape[0] >= 2 else 0)  # the warp (+ the one-launch prelude)
    assert got.device.type == torch.device(DEV).type and got.dtype == torch.float32 and got.shape == outs["out"].shape and got.is_contiguous()
    if kw["mode"] == "nearest":
        # a rounding tie of the index may fall on the other side between host and device
        bad = (got.cpu() - outs["out"]).abs() > (1e-5 + 1e-4 * outs["out"].abs())
        assert bad.float().mean().item() <= 0.02
    else:
        torch.testing.assert_close(got.cpu(), outs["out"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", WARPS)
def test_ingest_is_bit_identical_to_the_three_steps_on_device(name):
    """Same device, same sampler arithmetic: permute + .float() / 255 + the fp32 warp of this library == one kernel."""
    op, kw, ins, outs = ING.case(name)
    got = run_family_case(KT, op, kw, ins, device=DEV)
    img = ins["image"].to(DEV)
    img = img.unsqueeze(0) if img.dim() == 3 else img
    x = img.permute(0, 3, 1, 2).float()
    x = x / 255.0 if kw.get("normalize", True) else x
    fn = KT.warp_affine if op.startswith("warp_affine") else KT.warp_perspective
    args = {k: (tuple(v) if isinstance(v, list) else v) for k, v in kw.items() if k != "normalize"}
    if "fill_value" in ins:
        args["fill_value"] = ins["fill_value"].to(DEV)
    want = fn(x.contiguous(), ins["M"].to(DEV), **args)
    assert torch.equal(got, want), float((got - want).abs().max())


def test_ingest_full_size_properties():
    """1080p: the identity homography returns byte / 255 -- exactly under 'nearest'; the normalise / invert / unnormalise
    chain leaves the coordinates ~1e-4 px off the integers (as in the reference), so 'bilinear' is close, not equal; a
    pure integer translation moves the image and fills with zeros; frames with every byte value."""
    B, H, W = 4, 1080, 1920
    frames = torch.randint(0, 256, (B, H, W, 3), device=DEV, dtype=torch.uint8)
    frames[0, :16, :16, 0] = torch.arange(256, device=DEV, dtype=torch.uint8).reshape(16, 16)
    eye = torch.eye(3, device=DEV).expand(B, 3, 3).contiguous()
    want = frames.permute(0, 3, 1, 2).float() / 255.0
    assert torch.equal(KT.warp_perspective_from_uint8(frames, eye, (H, W), mode="nearest"), want)
    torch.testing.assert_close(KT.warp_perspective_from_uint8(frames, eye, (H, W)), want, rtol=0, atol=1e-3)
    exact = KT.warp_perspective_from_uint8(frames[:1], eye[:1], (H, W), mode="nearest", normalize="exact")  # torch's CPU form: a true division
    assert torch.equal(exact.cpu(), frames[:1].cpu().permute(0, 3, 1, 2).float() / 255.0)
    raw = KT.warp_perspective_from_uint8(frames, eye, (H, W), mode="nearest", normalize=False)
    assert torch.equal(raw, frames.permute(0, 3, 1, 2).float())
    shift = eye.clone()
    shift[:, 0, 2], shift[:, 1, 2] = 7.0, -3.0
    got = KT.warp_perspective_from_uint8(frames, shift, (H, W), mode="nearest")
    assert torch.equal(got[..., :-3, 7:], want[..., 3:, :-7]) and float(got[..., -3:, :].abs().max()) == 0.0 and float(got[..., :7].abs().max()) == 0.0
    one = KT.warp_affine_from_uint8(frames[0], eye[:1, :2], (H, W), mode="nearest")
    assert one.shape == (1, 3, H, W) and torch.equal(one[0], want[0])


def test_ingest_headline_homographies_match_the_fp32_path():
    """The bench's jittered-corner homographies at 1080p: equal to the TMA-tiled fp32 warp of the converted frames."""
    import bench

    B, H, W = 4, 1080, 1920
    frames = torch.randint(0, 256, (B, H, W, 3), device=DEV, dtype=torch.uint8)
    M = bench.make_homographies(B, 5).to(DEV)
    got = KT.warp_perspective_from_uint8(frames, M, (H, W))
    want = KT.warp_perspective((frames.permute(0, 3, 1, 2).float() / 255.0).contiguous(), M, (H, W))
    assert torch.equal(got, want), float((got - want).abs().max())


@pytest.mark.parametrize("pad", ["zeros", "border", "reflection", "fill"])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_tiled_ingest_kernel_is_bit_identical_to_the_per_tap_kernel(monkeypatch, pad, channels):
    """warp_u8_tiled_kernel (window staged in shared memory, each byte converted once) against warp_fwd_u8hwc (KB200_U8_SIMPLE=1):
    headline-like homographies, rotations that push tiles to the exact path, a horizon inside the image, partial tiles."""
    import bench
    from test_parity_gpu import _wild_matrices

    B, H, W = 6, 270, 480
    frames = torch.randint(0, 256, (B, H, W, channels), device=DEV, dtype=torch.uint8)
    quad = torch.tensor([[0.0, 0.0], [W - 1.0, 0.0], [W - 1.0, H - 1.0], [0.0, H - 1.0]]).expand(B, 4, 2)
    g = torch.Generator().manual_seed(3)
    M = bench.perspective_from_quads(quad, quad + 6.0 * torch.randn(B, 4, 2, generator=g)).to(DEV)
    wild = _wild_matrices(H, W).to(DEV)
    fill = {} if pad != "fill" else dict(fill_value=torch.tensor([0.2, 0.4, 0.6, 0.8][:channels]))
    cases = ((M, (H, W)), (M, (201, 333)), (wild, (H, W))) if not (pad == "fill" and channels != 3) else ()  # warp_perspective fills RGB only
    for mats, size in cases:
        img = frames[: mats.shape[0]] if mats.shape[0] <= B else frames[:1].expand(mats.shape[0], H, W, channels).contiguous()
        for ac in (True, False):
            K.config.set("u8_tiled", 0)
            want = KT.warp_perspective_from_uint8(img, mats, size, padding_mode=pad, align_corners=ac, **fill)
            K.config.set("u8_tiled", 1)
            got = KT.warp_perspective_from_uint8(img, mats, size, padding_mode=pad, align_corners=ac, **fill)
            assert torch.equal(got, want), float((got - want).abs().max())
    rot = KT.get_rotation_matrix2d(torch.tensor([[W / 2, H / 2]], device=DEV).expand(B, 2), torch.linspace(-40, 40, B, device=DEV), torch.ones(B, 2, device=DEV))
    K.config.set("u8_tiled", 0)
    want = KT.warp_affine_from_uint8(frames, rot, (H, W), padding_mode=pad, **fill)
    K.config.set("u8_tiled", 1)
    assert torch.equal(KT.warp_affine_from_uint8(frames, rot, (H, W), padding_mode=pad, **fill), want)


@pytest.mark.parametrize("name", UNDISTORT)
def test_undistort_from_bytes_matches_reference_and_the_fp32_path(name):
    """Golden vectors of image_to_tensor + _to_float32 + undistort_image from the reference (1e-4 rel); on the device: close to
    convert + undistort_image through the default maps + remap path, and EQUAL to convert + the fused fp32 undistort
    (switch fused_undistort), whose lens arithmetic the byte kernel shares."""
    op, kw, ins, outs = ING.case(name)
    K.config.set("fused_undistort", 0)
    before = K._ops.launch_count
    got = run_family_case(KC, op, kw, ins, device=DEV)
    img = ins["image"]
    one_kernel = ins["dist"].shape[-1] != 14 and img.shape[-1] in (1, 3) and img.shape[-2] % 4 == 0
    # one launch of this library: the byte kernel, or (tilt terms / odd widths) the remap kernel after torch built the maps
    assert before + (1 if one_kernel else 0) <= K._ops.launch_count <= before + 1
    assert got.device.type == torch.device(DEV).type and got.dtype == torch.float32 and got.shape == outs["out"].shape and got.is_contiguous()
    torch.testing.assert_close(got.cpu(), outs["out"], rtol=1e-4, atol=1e-5)
    x = img.to(DEV)
    x = (x.unsqueeze(0) if x.dim() == 3 else x).permute(0, 3, 1, 2).float()
    x = (x / 255.0 if kw.get("normalize", True) else x).contiguous()
    n = x.shape[0]
    cam, d = ins["K"].to(DEV), ins["dist"].to(DEV)
    cam, d = (cam if cam.dim() == 3 else cam.expand(n, 3, 3).contiguous()), (d if d.dim() == 2 else d.expand(n, d.shape[-1]).contiguous())
    torch.testing.assert_close(got, KC.undistort_image(x, cam, d), rtol=1e-4, atol=1e-5)
    if one_kernel:
        K.config.set("fused_undistort", 1)
        assert torch.equal(got, KC.undistort_image(x, cam, d))


def test_undistort_from_bytes_full_size():
    """1080p: zero coefficients return byte / 255 to the accuracy of the (p - c) / f * f + c round trip; a real lens equals the
    fp32 path on the converted frames to 1e-5 rel-L2; gradients w.r.t. the camera take the differentiable composition."""
    from helpers import rel_l2

    B, H, W = 2, 1080, 1920
    frames = torch.randint(0, 256, (B, H, W, 3), device=DEV, dtype=torch.uint8)
    cam = torch.tensor([[1500.0, 0.0, 960.0], [0.0, 1500.0, 540.0], [0.0, 0.0, 1.0]], device=DEV).expand(B, 3, 3).contiguous()
    x = (frames.permute(0, 3, 1, 2).float() / 255.0).contiguous()
    torch.testing.assert_close(KC.undistort_image_from_uint8(frames, cam, torch.zeros(B, 4, device=DEV)), x, rtol=0, atol=2e-3)
    dist = torch.tensor([[-0.2, 0.05, 0.001, -0.002, 0.01]], device=DEV).expand(B, 5).contiguous()
    before = K._ops.launch_count
    got = KC.undistort_image_from_uint8(frames, cam, dist)
    assert K._ops.launch_count == before + 1
    assert rel_l2(got, KC.undistort_image(x, cam, dist)) < 1e-5
    camg = cam.clone().requires_grad_(True)
    out = KC.undistort_image_from_uint8(frames[:, :64, :64].contiguous(), camg, dist)
    (g,) = torch.autograd.grad(out.sum(), [camg])
    assert g.shape == cam.shape and bool(torch.isfinite(g).all())
