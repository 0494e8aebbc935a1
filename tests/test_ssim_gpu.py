"""GPU parity of the fused SSIM kernel through the C ABI: vectors recorded from the
reference, the oracle on the same device, the library's own differentiable composition, and
size-independent properties at 1080p."""
import pytest
import torch

import kornia_b200 as K
from conftest import golden
from helpers import family_grads, rel_l2, run_family_case
from oracle import kornia_restated as R

pytestmark = pytest.mark.gpu
G = golden("ssim")
DEV = "cuda"
# the index divides two small quantities where the images are flat: the reference's own CPU/CUDA outputs differ by
# ~1e-5 absolute there, so the element-wise bar is the reference's fp32 test tolerance with the atol of an index in [-1, 1]
TOL = dict(rtol=1e-4, atol=2e-5)
FWD = [n for n in G.names() if not G.meta[n]["op"].endswith("_grad")]
GRAD = [n for n in G.names() if G.meta[n]["op"].endswith("_grad")]


def _impl(op):
    return K.losses if op.startswith("ssim_loss") else K.metrics


@pytest.mark.parametrize("name", FWD)
def test_ssim_forward_matches_reference(name):
    op, kw, ins, outs = G.case(name)
    before = K._ops.launch_count
    got = run_family_case(_impl(op), op, kw, ins, device=DEV)
    want = outs["out"].reshape(got.shape)  # reduced losses are 0-d; the .npz writer stores them as one-element arrays
    assert got.is_cuda and got.dtype == want.dtype
    if kw.get("reduction", "none") == "sum":
        assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want))
    else:
        torch.testing.assert_close(got.cpu(), want, **TOL)
    # one kernel for odd windows up to 11 taps on rows of a multiple of 4 floats (TMA rows are 16-byte aligned); the
    # composition (five one-pass blurs + torch elementwise ops) otherwise
    fused = kw["window_size"] <= 11 and ins["img1"].shape[-1] % 4 == 0
    assert (K._ops.launch_count - before == 1) == fused, (K._ops.launch_count - before, fused)


@pytest.mark.parametrize("name", GRAD)
def test_ssim_grads_match_reference(name):
    op, kw, ins, outs = G.case(name)
    got = family_grads(_impl(op), op, kw, ins, outs, device=DEV)
    for key, want in outs.items():
        if key != "cot":
            assert rel_l2(got[key].cpu(), want) < 1e-4, (key, rel_l2(got[key].cpu(), want))


@pytest.mark.parametrize("ws", [3, 5, 7, 9, 11])
def test_fused_ssim_equals_composed_path(ws):
    """One kernel vs this library's differentiable composition (five one-pass separable blurs + torch
    elementwise ops): same taps, same tap order, one rounding per op -> the same map."""
    a = torch.rand(3, 3, 101, 136, device=DEV)
    b = (a + 0.2 * torch.randn_like(a)).clamp(0, 1)
    fused = K.metrics.ssim(a, b, ws)
    composed = K.metrics.ssim(a.clone().requires_grad_(True), b, ws).detach()
    torch.testing.assert_close(fused, composed, rtol=1e-5, atol=1e-6)
    print(f"ssim ws={ws}: {(fused == composed).float().mean().item() * 100:.2f}% bit-identical, max abs diff {(fused - composed).abs().max().item():.2e}")
    torch.testing.assert_close(fused, R.ssim(a, b, ws), **TOL)
    torch.testing.assert_close(K.metrics.ssim(a, b, ws, padding="valid"), R.ssim(a, b, ws, padding="valid"), **TOL)


def test_ssim_full_size_properties():
    """1080p: ssim(x, x) = 1 everywhere; symmetric in its arguments; in [-1, 1]; the loss of identical images is 0;
    against the oracle on the same device."""
    x = torch.rand(4, 3, 1080, 1920, device=DEV)
    y = (x + 0.1 * torch.randn_like(x)).clamp(0, 1)
    same = K.metrics.ssim(x, x, 11)
    torch.testing.assert_close(same, torch.ones_like(same), rtol=0, atol=1e-4)
    xy, yx = K.metrics.ssim(x, y, 11), K.metrics.ssim(y, x, 11)
    torch.testing.assert_close(xy, yx, rtol=1e-5, atol=1e-6)
    assert float(xy.max()) <= 1.0 + 1e-5 and float(xy.min()) >= -1.0 - 1e-5
    assert float(K.losses.ssim_loss(x, x, 11)) < 1e-4
    torch.testing.assert_close(xy, R.ssim(x, y, 11), **TOL)
    assert abs(float(K.losses.ssim_loss(x, y, 11)) - float(R.ssim_loss(x, y, 11))) < 1e-5


def test_ssim_odd_shapes_and_offsets():
    buf = torch.rand(2 * 1 * 45 * 67 + 1, device=DEV)
    a = buf[1:].view(2, 1, 45, 67)                       # odd width, 4-byte aligned only
    b = a.flip(-2).contiguous()
    torch.testing.assert_close(K.metrics.ssim(a, b, 7), R.ssim(a, b, 7), **TOL)
    nc = torch.rand(2, 3, 40, 96, device=DEV)[:, :, :, ::2]   # non-contiguous input
    torch.testing.assert_close(K.metrics.ssim(nc, nc.flip(-1), 5), R.ssim(nc.contiguous(), nc.flip(-1).contiguous(), 5), **TOL)
    d = torch.rand(2, 2, 12, 14, device=DEV, dtype=torch.float64)  # fp64: composed path
    torch.testing.assert_close(K.metrics.ssim(d, d.flip(-1), 5).cpu(), R.ssim(d.cpu(), d.flip(-1).cpu(), 5), rtol=1e-9, atol=1e-10)
