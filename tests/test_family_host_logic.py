"""Host logic of the caller wrappers on CPU: the six core functions they sit on are swapped for the oracle's
CPU restatements, so what is exercised is exactly the code of kornia_b200/{geometry/transform/affwarp,crop2d},
filters/{blur,laplacian,unsharp} and metrics/ssim (matrix builders, broadcasting, squeezing, tap generators, crop
arithmetic, the SSIM composition) -- compared with the vectors recorded from the reference."""
import pytest
import torch

import kornia_b200 as K
from conftest import golden
from helpers import family_grads, rel_l2, run_family_case
from oracle import kornia_restated as R

FAM = golden("family")
SSIM = golden("ssim")
HOST_OPS = ("affine", "rotate", "translate", "scale", "shear", "crop_and_resize", "center_crop", "crop_by_boxes", "crop_by_transform_mat",
            "box_blur", "laplacian", "unsharp_mask")


@pytest.fixture
def core_on_cpu(monkeypatch):
    """Route the wrappers' calls into the core functions to the oracle (CPU); nothing else is touched."""
    from importlib import import_module  # several submodules are shadowed by the function of the same name

    blur, lap, unsharp = (import_module("kornia_b200.filters." + m) for m in ("blur", "laplacian", "unsharp"))
    affwarp, crop2d = (import_module("kornia_b200.geometry.transform." + m) for m in ("affwarp", "crop2d"))
    ssim_mod = import_module("kornia_b200.metrics.ssim")

    monkeypatch.setattr(affwarp, "warp_affine", R.warp_affine)
    monkeypatch.setattr(crop2d, "warp_affine", R.warp_affine)
    monkeypatch.setattr(crop2d, "warp_perspective", R.warp_perspective)
    monkeypatch.setattr(blur, "filter2d", R.filter2d)
    monkeypatch.setattr(blur, "filter2d_separable", R.filter2d_separable)
    monkeypatch.setattr(lap, "filter2d", R.filter2d)
    monkeypatch.setattr(unsharp, "gaussian_blur2d", R.gaussian_blur2d)
    monkeypatch.setattr(ssim_mod, "filter2d_separable", R.filter2d_separable)
    monkeypatch.setattr(K._ops, "_require_cuda", lambda t, what: None)


def _impl(op):
    base = op[:-5] if op.endswith("_grad") else op
    return K.filters if hasattr(K.filters, base) else K.geometry.transform


@pytest.mark.parametrize("name", [n for n in FAM.names() if FAM.meta[n]["op"] in HOST_OPS])
def test_wrapper_forward_on_cpu(core_on_cpu, name):
    op, kw, ins, outs = FAM.case(name)
    got = run_family_case(_impl(op), op, kw, ins)
    torch.testing.assert_close(got, outs["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", [n for n in FAM.names() if FAM.meta[n]["op"].endswith("_grad") and FAM.meta[n]["op"][:-5] in HOST_OPS])
def test_wrapper_grads_on_cpu(core_on_cpu, name):
    op, kw, ins, outs = FAM.case(name)
    got = family_grads(_impl(op), op, kw, ins, outs)
    for key, want in outs.items():
        if key != "cot":
            assert rel_l2(got[key], want) < 2e-6, (key, rel_l2(got[key], want))


@pytest.mark.parametrize("name", SSIM.names())
def test_ssim_composition_on_cpu(core_on_cpu, name):
    """metrics.ssim's differentiable composition (taken whenever a gradient is needed) and losses.ssim_loss."""
    op, kw, ins, outs = SSIM.case(name)
    impl = K.losses if op.startswith("ssim_loss") else K.metrics
    if op.endswith("_grad"):
        got = family_grads(impl, op, kw, ins, outs)
        for key, want in outs.items():
            if key != "cot":
                assert rel_l2(got[key], want) < 2e-6, key
        return
    leaves = {k: (v.clone().requires_grad_(True) if k == "img1" else v) for k, v in ins.items()}  # forces the composed path
    got = run_family_case(impl, op, kw, leaves).detach()
    torch.testing.assert_close(got, outs["out"].reshape(got.shape), rtol=1e-5, atol=1e-6)
