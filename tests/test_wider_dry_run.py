"""The GPU tests of the pyramid / resize / undistort callers (tests/test_wider_gpu.py) executed on CPU with the three core
functions swapped for the oracle: checks the TEST code and its tolerances (golden lookups, helper plumbing, full-size
property expectations) so that a first run on hardware can only fail on device arithmetic.  One test here per group of
GPU tests; the launch-count assertion of the resize test is the only line that needs the device."""
import pytest
import torch

from kornia_b200.geometry.calibration import undistort
from kornia_b200.geometry.transform import affwarp, pyramid
from oracle import kornia_restated as R

import test_wider_gpu as T


@pytest.fixture()
def on_cpu(monkeypatch):
    monkeypatch.setattr(pyramid, "filter2d", R.filter2d)
    monkeypatch.setattr(affwarp, "gaussian_blur2d", R.gaussian_blur2d)
    monkeypatch.setattr(undistort, "remap", R.remap)
    monkeypatch.setattr(T, "DEV", "cpu")


def test_golden_forward_cases(on_cpu):
    for name in T.FWD:
        T.test_wider_forward_matches_reference(name)


def test_golden_gradient_cases(on_cpu):
    for name in T.GRAD:
        T.test_wider_grads_match_reference(name)


def test_fp64_cases(on_cpu):
    marks = [m for m in T.test_wider_fp64_matches_oracle.pytestmark if m.name == "parametrize"]
    for name in marks[0].args[1]:
        T.test_wider_fp64_matches_oracle(name)


def test_property_tests(on_cpu):
    torch.manual_seed(0)
    T.test_pyrdown_full_size_properties()
    T.test_laplacian_pyramid_reconstructs_the_image()
    T.test_undistort_identity_and_full_size()
