"""Pin the oracle (CPU, no GPU needed).

1. oracle/kornia_restated.py (torch-op restatement of the Kornia composition) must reproduce the
   golden vectors that tests/golden/make_golden.py recorded from the imported reference.
2. oracle/aten_restated.py (numpy restatement of ATen grid_sampler_2d / pad+conv2d) must agree
   with torch's own CPU kernels, forward and (bilinear) backward, and with the goldens end to end.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden
from helpers import run_case
from oracle import aten_restated as A
from oracle import kornia_restated as R

WARP = golden("warp")
FILT = golden("filter")
TIGHT = dict(rtol=1e-5, atol=2e-6)  # same ATen kernels; slack only for BLAS differences across hosts


@pytest.mark.parametrize("name", WARP.names("warp_perspective") + WARP.names("warp_affine") + WARP.names("remap"))
def test_restated_warps_match_reference(name):
    op, kw, ins, outs = WARP.case(name)
    torch.testing.assert_close(run_case(R, op, kw, ins), outs["out"], **TIGHT)


@pytest.mark.parametrize("name", FILT.names("filter2d") + FILT.names("filter2d_separable") + FILT.names("gaussian_blur2d"))
def test_restated_filters_match_reference(name):
    op, kw, ins, outs = FILT.case(name)
    torch.testing.assert_close(run_case(R, op, kw, ins), outs["out"], **TIGHT)


@pytest.mark.parametrize("name", FILT.names("gaussian_taps"))
def test_restated_gaussian_taps(name):
    _, kw, _, outs = FILT.case(name)
    got = R.gaussian_taps(kw["kernel_size"], torch.tensor([[kw["sigma"]]]))
    torch.testing.assert_close(got, outs["out"], **TIGHT)


def test_gaussian_taps_reference_literals():
    # literal pins from the reference's docstrings (kornia/filters/kernels.py:572-579)
    torch.testing.assert_close(R.gaussian_taps(3, torch.tensor([[2.5]])), torch.tensor([[0.3243, 0.3513, 0.3243]]), atol=1e-4, rtol=0)
    torch.testing.assert_close(R.gaussian_taps(5, torch.tensor([[1.5], [0.7]])),
                               torch.tensor([[0.1201, 0.2339, 0.2921, 0.2339, 0.1201], [0.0096, 0.2054, 0.5699, 0.2054, 0.0096]]),
                               atol=1e-4, rtol=0)


GRAD_CASES = [n for op in ("warp_perspective_grad", "warp_affine_grad", "remap_grad") for n in WARP.names(op)]


@pytest.mark.parametrize("name", GRAD_CASES)
def test_restated_warp_grads_match_reference(name):
    op, kw, ins, outs = WARP.case(name)
    got = run_case(R, op, kw, ins)
    for k, v in outs.items():
        torch.testing.assert_close(got[k], v, rtol=1e-4, atol=1e-5)


FGRAD_CASES = [n for op in ("filter2d_grad", "filter2d_separable_grad", "gaussian_blur2d_grad") for n in FILT.names(op)]


@pytest.mark.parametrize("name", FGRAD_CASES)
def test_restated_filter_grads_match_reference(name):
    op, kw, ins, outs = FILT.case(name)
    got = run_case(R, op, kw, ins)
    for k, v in outs.items():
        torch.testing.assert_close(got[k], v, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ numpy restatement of ATen
MODES = {"bilinear": A.BILINEAR, "nearest": A.NEAREST, "bicubic": A.BICUBIC}
PADS = {"zeros": A.ZEROS, "border": A.BORDER, "reflection": A.REFLECTION}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("pad", list(PADS))
@pytest.mark.parametrize("ac", [True, False])
def test_numpy_grid_sample_matches_torch(mode, pad, ac):
    g = torch.Generator().manual_seed(7)
    img = torch.rand(2, 3, 9, 13, generator=g, dtype=torch.float64)
    grid = torch.rand(2, 8, 11, 2, generator=g, dtype=torch.float64) * 3.4 - 1.7
    want = F.grid_sample(img, grid, mode=mode, padding_mode=pad, align_corners=ac)
    got = A.grid_sample(img.numpy(), grid.numpy(), MODES[mode], PADS[pad], ac)
    np.testing.assert_allclose(got, want.numpy(), rtol=1e-10, atol=1e-12)
    # fp32 too (the eager op order matters here)
    want32 = F.grid_sample(img.float(), grid.float(), mode=mode, padding_mode=pad, align_corners=ac)
    got32 = A.grid_sample(img.float().numpy(), grid.float().numpy(), MODES[mode], PADS[pad], ac)
    if mode == "nearest":
        # a rounding tie may flip a tap between the two fp32 evaluations of the coordinate; allow a handful
        assert (np.abs(got32 - want32.numpy()) > 1e-6).mean() < 0.01
    else:
        np.testing.assert_allclose(got32, want32.numpy(), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("pad", list(PADS))
@pytest.mark.parametrize("ac", [True, False])
def test_numpy_bilinear_backward_matches_torch(pad, ac):
    g = torch.Generator().manual_seed(11)
    img = torch.rand(2, 3, 7, 9, generator=g, dtype=torch.float64, requires_grad=True)
    grid = (torch.rand(2, 6, 8, 2, generator=g, dtype=torch.float64) * 3.0 - 1.5).requires_grad_(True)
    cot = torch.rand(2, 3, 6, 8, generator=g, dtype=torch.float64)
    out = F.grid_sample(img, grid, mode="bilinear", padding_mode=pad, align_corners=ac)
    gi, gg = torch.autograd.grad((out * cot).sum(), [img, grid])
    ni, ng = A.grid_sample_bilinear_backward(cot.numpy(), img.detach().numpy(), grid.detach().numpy(), PADS[pad], ac)
    np.testing.assert_allclose(ni, gi.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ng, gg.numpy(), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("name", [n for n in FILT.names("filter2d") if "_corr" in n][::3])
def test_numpy_filter2d_matches_reference(name):
    _, kw, ins, outs = FILT.case(name)
    got = A.filter2d(ins["input"].numpy(), ins["kernel"].numpy(), kw["border_type"], kw["normalized"], kw["padding"], kw["behaviour"])
    np.testing.assert_allclose(got, outs["out"].numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", [n for n in WARP.names("warp_perspective") if "fill" not in n])
def test_numpy_chain_matches_reference_warp(name):
    """End to end: reference prelude (torch) -> numpy perspective grid -> numpy sampler == golden."""
    _, kw, ins, outs = WARP.case(name)
    src, M = ins["src"], ins["M"]
    h, w = kw["dsize"]
    m = R.inv3x3(R.normalized_homography(M, src.shape[-2:], (h, w))).numpy()
    xs, ys = R.meshgrid_axes(h, w, "cpu")
    bx, by = xs.numpy()[None, None, :], ys.numpy()[None, :, None]
    mm = lambda i, j: m[:, i, j, None, None]  # noqa: E731
    den = mm(2, 0) * bx + mm(2, 1) * by + mm(2, 2)
    gx = (mm(0, 0) * bx + mm(0, 1) * by + mm(0, 2)) / den
    gy = (mm(1, 0) * bx + mm(1, 1) * by + mm(1, 2)) / den
    got = A.grid_sample(src.numpy(), np.stack([gx, gy], -1), MODES[kw["mode"]], PADS[kw["padding_mode"]], kw["align_corners"])
    if kw["mode"] == "nearest":
        assert (np.abs(got - outs["out"].numpy()) > 1e-6).mean() < 0.01
    else:
        np.testing.assert_allclose(got, outs["out"].numpy(), rtol=2e-5, atol=2e-6)
