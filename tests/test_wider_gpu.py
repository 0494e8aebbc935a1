"""GPU parity of the remaining callers of SURVEY.md 8f (pyramids, resize family, undistort_image) through the C ABI:
golden vectors recorded from the reference (fp32 CPU), the oracle in fp64, and full-size properties.  These functions
are host compositions over kernels with their own parity tests (filter2d 5x5 tiled, the one-pass separable blur, the
tiled remap); what is checked here is the composition on the device (devices / dtypes / strides of the
intermediates)."""
import pytest
import torch

import kornia_b200 as K
from conftest import golden
from helpers import family_grads, rel_l2, run_family_case
from oracle import kornia_restated as R

pytestmark = pytest.mark.gpu
WID = golden("wider")
FP32 = dict(rtol=1e-4, atol=1e-5)
DEV = "cuda"
FWD = [n for n in WID.names() if not WID.meta[n]["op"].endswith("_grad")]
GRAD = [n for n in WID.names() if WID.meta[n]["op"].endswith("_grad")]


def product_module(op):
    base = op[:-5] if op.endswith("_grad") else op
    return K.geometry.calibration if hasattr(K.geometry.calibration, base) else K.geometry.transform


def as_list(out, outs):
    if isinstance(out, (list, tuple)):
        return list(out), [outs[f"out{i}"] for i in range(len(out))]
    return [out], [outs["out"]]


@pytest.mark.parametrize("name", FWD)
def test_wider_forward_matches_reference(name):
    op, kw, ins, outs = WID.case(name)
    got, want = as_list(run_family_case(product_module(op), op, kw, ins, device=DEV), outs)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.device.type == torch.device(DEV).type and g.dtype == w.dtype and g.shape == w.shape
        if kw.get("interpolation") == "nearest":
            # index rounding of F.interpolate('nearest') may fall on the other side of a tie between host and device
            bad = (g.cpu() - w).abs() > (1e-5 + 1e-4 * w.abs())
            assert bad.float().mean().item() <= 0.02
        else:
            torch.testing.assert_close(g.cpu(), w, **FP32)


@pytest.mark.parametrize("name", GRAD)
def test_wider_grads_match_reference(name):
    op, kw, ins, outs = WID.case(name)
    got = family_grads(product_module(op), op, kw, ins, outs, device=DEV)
    for key, want in outs.items():
        if key == "cot":
            continue
        # image gradients to the path's 1e-4; the few-parameter gradients (K, dist, points) are sums over every pixel
        # whose order differs between host and device: 5e-4
        tol = 1e-4 if key in ("out", "grad_input", "grad_image") else 5e-4
        assert rel_l2(got[key].cpu(), want) < tol, (key, rel_l2(got[key].cpu(), want))


@pytest.mark.parametrize("name", ["pyrdown_reflect_0", "pyrup_replicate_1", "laplacian_pyr_padded", "resize_aa_down", "rescale_pair",
                                  "undistort_8", "undistort_14", "undistort_5d"])
def test_wider_fp64_matches_oracle(name):
    op, kw, ins, outs = WID.case(name)
    got, _ = as_list(run_family_case(product_module(op), op, kw, ins, device=DEV, dtype=torch.float64), outs)
    want, _ = as_list(run_family_case(R, op, kw, ins, dtype=torch.float64), outs)
    for g, w in zip(got, want):
        assert g.dtype == torch.float64
        torch.testing.assert_close(g.cpu(), w, rtol=1e-8, atol=1e-9)


def test_pyrdown_full_size_properties():
    """1080p: a constant image stays constant under every border but 'constant'; the blur-then-average of a linear ramp
    stays the ramp in the interior (the binomial stencil and the 2x2 average are both symmetric)."""
    x = torch.full((2, 3, 1080, 1920), 0.375, device=DEV)
    for border in ("reflect", "replicate", "circular"):
        out = K.geometry.transform.pyrdown(x, border)
        assert out.shape == (2, 3, 540, 960)
        torch.testing.assert_close(out, torch.full_like(out, 0.375), rtol=1e-6, atol=1e-6)
    ramp = torch.arange(1920, device=DEV, dtype=torch.float32).expand(1, 1, 1080, 1920).contiguous() / 1920
    out = K.geometry.transform.pyrdown(ramp)
    want = (torch.arange(960, device=DEV, dtype=torch.float32) * 2 + 0.5) / 1920
    torch.testing.assert_close(out[0, 0, 100, 4:-4], want[4:-4], rtol=1e-5, atol=1e-6)
    pyr = K.geometry.transform.build_pyramid(x[:1], 4)
    assert [tuple(p.shape[-2:]) for p in pyr] == [(1080, 1920), (540, 960), (270, 480), (135, 240)]


def test_laplacian_pyramid_reconstructs_the_image():
    """Collapsing the Laplacian pyramid (upsample + add, coarse to fine) returns the input: a size-independent identity."""
    KT = K.geometry.transform
    x = torch.rand(2, 3, 256, 512, device=DEV)
    bands = KT.build_laplacian_pyramid(x, 4)
    cur = bands[-1]
    for band in reversed(bands[:-1]):
        cur = band + KT.pyrup(cur)
    torch.testing.assert_close(cur, x, rtol=1e-5, atol=1e-5)


def test_undistort_identity_and_full_size():
    """Zero coefficients: the maps are the pixel grid and the output is the input (bilinear at integer positions); at
    1080p the product equals the oracle's composition run on the same device with torch's grid_sample."""
    KC = K.geometry.calibration
    img = torch.rand(2, 3, 270, 480, device=DEV)
    cam = torch.tensor([[400.0, 0.0, 240.0], [0.0, 400.0, 135.0], [0.0, 0.0, 1.0]], device=DEV).expand(2, 3, 3).contiguous()
    out = KC.undistort_image(img, cam, torch.zeros(2, 4, device=DEV))
    # (p - c) / f * f + c returns p to ~3e-5 px at p = 480; on white noise that is up to ~5e-5 in value (measured on the oracle)
    torch.testing.assert_close(out, img, rtol=0, atol=2e-4)
    big = torch.rand(1, 3, 1080, 1920, device=DEV)
    cam = torch.tensor([[[1500.0, 0.0, 960.0], [0.0, 1500.0, 540.0], [0.0, 0.0, 1.0]]], device=DEV)
    dist = torch.tensor([[-0.2, 0.05, 0.001, -0.002, 0.01]], device=DEV)
    got = KC.undistort_image(big, cam, dist)
    want = R.undistort_image(big, cam, dist)
    assert got.shape == big.shape and got.is_contiguous()
    assert rel_l2(got, want) < 1e-5


def test_resize_antialias_runs_the_blur_kernel_and_modules_work():
    KT = K.geometry.transform
    before = K._ops.launch_count
    x = torch.rand(2, 3, 240, 320, device=DEV)
    out = KT.resize(x, (60, 80), antialias=True)
    assert K._ops.launch_count > before and out.shape == (2, 3, 60, 80)
    want = R.resize(x, (60, 80), antialias=True)
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-5)
    assert KT.Resize((30, 40), antialias=True)(x).shape == (2, 3, 30, 40)
    assert KT.Rescale(0.5)(x).shape == (2, 3, 120, 160)
    assert KT.PyrDown()(x).shape == (2, 3, 120, 160) and KT.PyrUp()(x).shape == (2, 3, 480, 640)
