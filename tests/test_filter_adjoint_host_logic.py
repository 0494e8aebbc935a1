"""Band arithmetic of the fast separable adjoint (kornia_b200/filters/_adjoint.py, opt-in with KB200_FAST_FILTER_BWD=1) on CPU:
the two primitives are swapped for torch ops (the oracle's filter and its autograd), so what runs is exactly the host logic
-- flipped taps under a zero border for the image-sized pass, exact recomputation of the four border bands on crops --
against the autograd adjoint of the whole image."""
import pytest
import torch

from kornia_b200.filters._adjoint import separable_adjoint
from oracle import kornia_restated as R

BORDERS = {"constant": 0, "reflect": 1, "replicate": 2}


def exact_adjoint_of(border):
    def fn(g, kx, ky):
        x = torch.zeros_like(g, requires_grad=True)  # the filter is linear: the adjoint does not depend on x
        (gx,) = torch.autograd.grad(R.filter2d_separable(x, kx, ky, border), [x], g)
        return gx
    return fn


def forward_constant(g, kx, ky):
    return R.filter2d_separable(g, kx, ky, "constant")


@pytest.mark.parametrize("border", ["constant", "reflect", "replicate"])
@pytest.mark.parametrize("ksize", [3, 5, 11])
@pytest.mark.parametrize("shape", [(2, 3, 40, 52), (1, 1, 36, 36), (3, 2, 64, 48)])
def test_fast_adjoint_equals_autograd_adjoint(border, ksize, shape):
    g = torch.Generator().manual_seed(ksize * 7 + len(border))
    gout = torch.rand(*shape, generator=g, dtype=torch.float64) - 0.5
    kx = torch.rand(shape[0], ksize, generator=g, dtype=torch.float64)    # asymmetric, per-sample taps: the flip matters
    ky = torch.rand(1, ksize, generator=g, dtype=torch.float64)
    exact = exact_adjoint_of(border)
    calls = []

    def counted_exact(gg, a, b):
        calls.append(tuple(gg.shape[-2:]))
        return exact(gg, a, b)

    got = separable_adjoint(gout, kx, ky, BORDERS[border], forward_constant, counted_exact)
    want = exact(gout, kx, ky)
    torch.testing.assert_close(got, want, rtol=1e-12, atol=1e-12)
    h = (ksize - 1) // 2
    if min(shape[-2:]) <= 3 * h + 2:
        assert calls == [tuple(shape[-2:])]                      # too small for the band argument: the exact form
    elif border == "constant":
        assert calls == []                                       # one fast pass, nothing else
    else:
        crop = 3 * h + 2                                         # four crops, cost proportional to the perimeter
        assert calls == [(crop, shape[-1])] * 2 + [(shape[-2], crop)] * 2


def test_small_images_take_the_exact_form():
    gout = torch.rand(1, 1, 8, 40, dtype=torch.float64)   # H == 3h + 2
    k = torch.rand(1, 5, dtype=torch.float64)
    exact = exact_adjoint_of("reflect")
    torch.testing.assert_close(separable_adjoint(gout, k, k, 1, forward_constant, exact), exact(gout, k, k), rtol=0, atol=0)
