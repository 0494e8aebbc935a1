"""uint8 ingest warps on CPU: the oracle's restatement against golden vectors recorded from the
reference's three steps (tests/golden/make_golden_ingest.py), the host logic of the product functions with the device
call swapped for the oracle, the validation contract, and the byte -> [0,1] conversion the kernel uses."""
import inspect

import numpy as np
import pytest
import torch

import kornia_b200 as K
from kornia_b200 import _ops
from kornia_b200.geometry.transform import ingest
from conftest import golden
from helpers import run_family_case
from oracle import kornia_restated as R

ING = golden("ingest")
FP32 = dict(rtol=1e-4, atol=1e-5)


def product_module(op):
    return K.geometry.calibration if op.startswith("undistort") else K.geometry.transform


@pytest.mark.parametrize("name", ING.names())
def test_oracle_matches_reference(name):
    op, kw, ins, outs = ING.case(name)
    got = run_family_case(R, op, kw, ins)
    assert got.dtype == torch.float32 and got.shape == outs["out"].shape
    torch.testing.assert_close(got, outs["out"], rtol=1e-6, atol=1e-6)


@pytest.fixture()
def device_call_on_cpu(monkeypatch):
    """_ops.warp_u8hwc replaced by the same contract on CPU: (B,H,W,C) bytes + the prelude's sampling matrix and axes ->
    grid_sample on the converted image.  What runs for real is everything above the C call."""
    import torch.nn.functional as F

    def mirror(image, m, bx, by, fill, h, w, projective, interp, pad, align, normalize):
        assert image.dtype == torch.uint8 and image.dim() == 4 and m.dtype == torch.float32 and normalize in (0, 1, 2)
        x = R.image_to_float(image, normalize != 0)
        grid = R.perspective_grid(m, bx, by) if projective else R.affine_grid(m, bx, by)
        if grid.shape[0] == 1 and x.shape[0] > 1:
            grid = grid.expand(x.shape[0], -1, -1, -1)
        mode = {0: "bilinear", 1: "nearest", 2: "bicubic"}[interp]
        if pad == 3:
            fv = fill.to(x).reshape(-1)
            return R.fill_and_sample(x, grid, mode, align, fv)
        return F.grid_sample(x, grid, mode=mode, padding_mode={0: "zeros", 1: "border", 2: "reflection"}[pad], align_corners=align)

    monkeypatch.setattr(_ops, "warp_u8hwc", mirror)

    def undistort_mirror(image, lens, normalize):
        """kb200_undistort_u8hwc_forward's contract on CPU: (B,16) lens rows -> camera matrix + 12 coefficients -> the oracle."""
        assert image.dtype == torch.uint8 and image.dim() == 4 and lens.shape == (image.shape[0], 16) and normalize in (0, 1, 2)
        cam = torch.zeros(lens.shape[0], 3, 3)
        cam[:, 0, 0], cam[:, 1, 1], cam[:, 0, 2], cam[:, 1, 2], cam[:, 2, 2] = lens[:, 0], lens[:, 1], lens[:, 2], lens[:, 3], 1.0
        if image.shape[-1] not in (1, 3) or image.shape[2] % 4 != 0:
            raise _ops._lib.Unsupported("outside the tiled kernel's envelope")
        return R.undistort_image(R.image_to_float(image, normalize != 0), cam, lens[:, 4:].contiguous())

    from kornia_b200.geometry.calibration import undistort

    monkeypatch.setattr(_ops, "undistort_u8hwc", undistort_mirror)
    monkeypatch.setattr(_ops, "undistort_fused", lambda *a, **k: (_ for _ in ()).throw(_ops._lib.Unsupported("no device here")))  # fp32 fused kernel
    monkeypatch.setattr(undistort, "remap", R.remap)                              # the composition path (tilt, odd widths)

    def take_the_kernel_path():  # undistort only: its host code asks image.is_cuda before calling the C entry
        monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))

    return take_the_kernel_path


@pytest.mark.parametrize("name", ING.names())
def test_product_host_logic_on_cpu(device_call_on_cpu, name):
    op, kw, ins, outs = ING.case(name)
    if op.startswith("undistort"):
        device_call_on_cpu()
    got = run_family_case(product_module(op), op, kw, ins)
    torch.testing.assert_close(got, outs["out"], **FP32)


def test_signatures_extend_the_warps():
    for ours, base in ((ingest.warp_perspective_from_uint8, K.geometry.transform.warp_perspective),
                       (ingest.warp_affine_from_uint8, K.geometry.transform.warp_affine)):
        a, b = inspect.signature(ours), inspect.signature(base)
        names = list(a.parameters)
        assert names[0] == "image" and names[1:-1] == list(b.parameters)[1:] and names[-1] == "normalize"
        for n in names[1:-1]:
            assert a.parameters[n].default == b.parameters[n].default


def test_validation_runs_before_device_work():
    f = ingest.warp_perspective_from_uint8
    img, M = torch.zeros(2, 8, 9, 3, dtype=torch.uint8), torch.eye(3).expand(2, 3, 3)
    with pytest.raises(TypeError):
        f(img.float(), M, (8, 9))
    with pytest.raises(TypeError):
        f(img.numpy(), M, (8, 9))
    with pytest.raises(ValueError):
        f(img[0, 0], M, (8, 9))
    with pytest.raises(ValueError):
        f(img, M[:, :2], (8, 9))
    with pytest.raises(ValueError):
        f(img, M, (8, 9), mode="cubic")
    with pytest.raises(ValueError):
        f(img, M, (8, 9), padding_mode="fill", fill_value=torch.zeros(2))
    with pytest.raises(RuntimeError, match="same batch size"):
        f(img, M[:1], (8, 9))
    with pytest.raises(ValueError, match="normalize"):
        f(img, M, (8, 9), normalize="255")
    with pytest.raises(RuntimeError, match="forward-only"):
        f(img, M.clone().requires_grad_(True), (8, 9))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        f(img, M, (8, 9))  # valid request, CPU tensors: the engine has no CPU path
    g = ingest.warp_affine_from_uint8
    with pytest.raises(ValueError):
        g(img, M, (8, 9))
    with pytest.raises(RuntimeError, match="same batch size"):
        g(img, torch.zeros(3, 2, 3), (8, 9))


def test_byte_conversions():
    """csrc/warp_u8.cuh: normalize=2 (unit_from_byte: q0 = u * RN(1/255), e = fma(-255, q0, u), q = fma(e, r, q0)) equals
    float(u) / 255.0f, torch's CPU division, for all 256 bytes (the products are exact in double, so numpy reproduces the
    fmas); normalize=1 is q0 itself, torch's CUDA form of the same expression, within one ulp of it."""
    u = np.arange(256, dtype=np.float32)
    r = np.float32(1.0) / np.float32(255.0)
    assert r.view(np.uint32) == 0x3B808081  # the literal in the kernel
    q0 = (u * r).astype(np.float32)
    e = (np.float64(-255.0) * q0.astype(np.float64) + u.astype(np.float64)).astype(np.float32)
    q = (e.astype(np.float64) * np.float64(r) + q0.astype(np.float64)).astype(np.float32)
    want = (torch.arange(256, dtype=torch.uint8).float() / 255.0).numpy()
    assert np.array_equal(q, want) and not np.array_equal(q0, want)
    assert np.abs(q0.view(np.int32) - want.view(np.int32)).max() == 1


def test_undistort_from_uint8_contract(monkeypatch):
    f = K.geometry.calibration.undistort_image_from_uint8
    img, cam, d = torch.zeros(2, 8, 12, 3, dtype=torch.uint8), torch.eye(3).expand(2, 3, 3), torch.zeros(2, 5)
    with pytest.raises(TypeError):
        f(img.float(), cam, d)
    with pytest.raises(ValueError):
        f(img, cam[:, :2], d)
    with pytest.raises(ValueError):
        f(img, cam, torch.zeros(2, 6))
    with pytest.raises(ValueError, match="normalize"):
        f(img, cam, d, normalize=255)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        f(img, cam, d)  # valid request, CPU tensors: no CPU path
    a, b = inspect.signature(f), inspect.signature(K.geometry.calibration.undistort_image)
    assert list(a.parameters)[:-1] == list(b.parameters) and list(a.parameters)[-1] == "normalize"
