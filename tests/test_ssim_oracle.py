"""CPU checks for ssim / ssim_loss: the oracle against the vectors recorded from the reference, and the host-side
contract of the product functions."""
import inspect

import pytest
import torch

import kornia_b200 as K
from conftest import golden
from helpers import family_grads, rel_l2, run_family_case
from oracle import kornia_restated as R

G = golden("ssim")


@pytest.mark.parametrize("name", G.names())
def test_oracle_ssim_matches_reference(name):
    op, kw, ins, outs = G.case(name)
    if op.endswith("_grad"):
        got = family_grads(R, op, kw, ins, outs)
        for key, want in outs.items():
            if key != "cot":
                assert rel_l2(got[key], want) < 2e-6, key
    else:
        got = run_family_case(R, op, kw, ins)
        # reduced losses are 0-d; the .npz writer stores them as one-element arrays
        torch.testing.assert_close(got, outs["out"].reshape(got.shape), rtol=1e-5, atol=1e-6)


def test_ssim_contract():
    sig = lambda fn: "(" + ", ".join(p.name if p.default is inspect._empty else f"{p.name}={p.default!r}"
                                     for p in inspect.signature(fn).parameters.values()) + ")"
    # kornia/metrics/ssim.py:34-41, kornia/losses/ssim.py:26-34
    assert sig(K.metrics.ssim) == "(img1, img2, window_size, max_val=1.0, eps=1e-12, padding='same')"
    assert sig(K.losses.ssim_loss) == "(img1, img2, window_size, max_val=1.0, eps=1e-12, reduction='mean', padding='same')"
    a = torch.rand(1, 2, 8, 8)
    with pytest.raises(TypeError):
        K.metrics.ssim(1.0, a, 5)
    with pytest.raises(TypeError):
        K.metrics.ssim(a, a, 5, max_val=1)
    with pytest.raises(ValueError, match="BxCxHxW"):
        K.metrics.ssim(a[0], a, 5)
    with pytest.raises(ValueError, match="must be the same"):
        K.metrics.ssim(a, a[:, :1], 5)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        K.metrics.ssim(a, a, 5)
    assert K.metrics.SSIM(5).window_size == 5 and K.losses.SSIMLoss(7, reduction="sum").reduction == "sum"
