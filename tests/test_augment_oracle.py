"""kornia_b200.augmentation against golden vectors of the reference's RandomPerspective / RandomAffine / RandomGaussianBlur
(tests/golden/make_golden_augment.py).  CPU part (this file's non-gpu tests): the host logic with the six core functions
swapped for the oracle -- (a) the parameter stream: same seed, same draws as the reference, bit for bit; (b) replaying the
recorded parameters reproduces the recorded outputs and transform matrices.  GPU part: the same replay through the CUDA
kernels, and sampling on the device."""
import importlib

import pytest
import torch

from conftest import golden
from oracle import kornia_restated as R

A = importlib.import_module("kornia_b200.augmentation")
AUG = golden("augment")
NAMES = AUG.names()


def _build(name):
    cls, kw, ins, outs = AUG.case(name)
    ctor = {k: (tuple(v) if isinstance(v, list) else v) for k, v in kw["ctor"].items()}
    return getattr(A, cls)(**ctor), kw["seed"], ins["input"], outs


def _params(outs, device=None):
    p = {k[len("param_"):]: v for k, v in outs.items() if k.startswith("param_")}
    return {k: (v.to(device) if device is not None and k not in ("forward_input_shape",) else v) for k, v in p.items()}


@pytest.fixture()
def on_cpu(monkeypatch):
    monkeypatch.setattr(A, "warp_perspective", R.warp_perspective)
    monkeypatch.setattr(A, "warp_affine", R.warp_affine)
    monkeypatch.setattr(A, "gaussian_blur2d", R.gaussian_blur2d)
    monkeypatch.setattr(A, "get_perspective_transform", R.get_perspective_transform)
    monkeypatch.setattr(A, "get_rotation_matrix2d", R.get_rotation_matrix2d)


@pytest.mark.parametrize("name", NAMES)
def test_same_seed_same_parameters_as_the_reference(on_cpu, name):
    aug, seed, x, outs = _build(name)
    torch.manual_seed(seed)
    got = aug.forward_parameters(tuple(x.reshape((1,) * (4 - x.dim()) + tuple(x.shape)).shape))
    want = _params(outs)
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k, w in want.items():
        assert got[k].shape == w.shape and got[k].dtype == w.dtype, (k, got[k].shape, w.shape, got[k].dtype, w.dtype)
        assert torch.equal(got[k], w), (k, got[k], w)


@pytest.mark.parametrize("name", NAMES)
def test_replayed_parameters_reproduce_the_reference_on_cpu(on_cpu, name):
    aug, seed, x, outs = _build(name)
    out = aug(x, params=_params(outs))
    assert out.shape == outs["out"].shape
    torch.testing.assert_close(out, outs["out"], rtol=1e-4, atol=1e-5)
    if "transform_matrix" in outs:
        torch.testing.assert_close(aug.transform_matrix, outs["transform_matrix"], rtol=1e-4, atol=1e-4)
    torch.manual_seed(seed)  # and end to end from the seed
    torch.testing.assert_close(aug(x), outs["out"], rtol=1e-4, atol=1e-5)


def test_constructor_contract():
    with pytest.raises(NotImplementedError):
        A.RandomPerspective(sampling_method="bogus")
    with pytest.raises(TypeError):
        A.RandomGaussianBlur((3, 3), (2.0, 1.0))
    with pytest.raises(ValueError):
        A.RandomAffine(10.0, scale=(1.0, 2.0, 3.0))._ranges()
    aug = A.RandomPerspective(p=1.0)
    with pytest.raises(TypeError):
        aug(torch.zeros(2, 3, 8, 8, dtype=torch.int32))
    with pytest.raises(TypeError):
        aug([1, 2, 3])


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_replayed_parameters_reproduce_the_reference_on_the_device(name):
    aug, seed, x, outs = _build(name)
    out = aug(x.cuda(), params=_params(outs, "cuda"))
    assert out.is_cuda and out.shape == outs["out"].shape
    cls = AUG.case(name)[0]
    if "nearest" in name:  # a tap on a rounding boundary may flip between CPU and GPU coordinate chains: compare where they agree
        close = torch.isclose(out.cpu(), outs["out"], rtol=1e-4, atol=1e-5)
        assert close.float().mean() > 0.98, float(close.float().mean())
    else:
        torch.testing.assert_close(out.cpu(), outs["out"], rtol=1e-4, atol=2e-5)
    if "transform_matrix" in outs:
        torch.testing.assert_close(aug.transform_matrix.cpu(), outs["transform_matrix"], rtol=1e-4, atol=1e-4)
    assert cls in ("RandomPerspective", "RandomAffine", "RandomGaussianBlur")


@pytest.mark.gpu
def test_sampling_on_the_device_and_one_launch_homographies():
    import kornia_b200 as K

    x = torch.rand(32, 3, 256, 256, device="cuda")
    for aug in (A.RandomPerspective(0.5, p=1.0), A.RandomAffine(30.0, translate=(0.1, 0.1), scale=(0.8, 1.2), shear=10.0, p=1.0),
                A.RandomGaussianBlur((5, 5), (0.1, 2.0), p=1.0)):
        aug = aug.to("cuda")
        torch.manual_seed(0)
        before = K._ops.launch_count
        a = aug(x)
        launches = K._ops.launch_count - before
        assert a.shape == x.shape and a.is_cuda and torch.isfinite(a).all()
        assert all(v.is_cuda for k, v in aug._params.items() if k != "forward_input_shape"), "parameters are sampled on the device"
        b = aug(x, params=aug._params)
        assert torch.equal(a, b), "replaying the sampled parameters is deterministic"
        # points -> homography and the matrix chain are one launch each, the image work one (or two tile-shape) launches
        assert launches <= 5, launches
