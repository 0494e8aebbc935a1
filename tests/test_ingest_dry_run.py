"""The gated GPU tests of the uint8 ingest path (tests/test_ingest_gpu.py) executed on CPU: the two C entry points are
replaced by the oracle's composition and the fp32 product functions they are compared with by the oracle, so what runs is
the TEST code -- golden lookups, launch-count bookkeeping, tolerances, the full-size property expectations.  A first run on
hardware can then only fail on device arithmetic.  (torch evaluates ``x / 255.0`` as a division on CPU and as a multiply
by the fp32 reciprocal on CUDA; the mirror and the "three steps" both use the CUDA form here, as on the device.)"""
import pytest
import torch
import torch.nn.functional as F

from kornia_b200 import _lib, _ops, config
from kornia_b200.geometry.calibration import undistort
from oracle import kornia_restated as R

import test_ingest_gpu as T

RCP = torch.tensor(1.0 / 255.0, dtype=torch.float32)


def _levels(image, normalize):
    x = R.image_to_float(image, False)
    return x * RCP if normalize == 1 else x / 255.0 if normalize == 2 else x


def _warp_mirror(image, m, bx, by, fill, h, w, projective, interp, pad, align, normalize):
    x = _levels(image, normalize)
    grid = R.perspective_grid(m, bx, by) if projective else R.affine_grid(m, bx, by)
    if grid.shape[0] == 1 and x.shape[0] > 1:
        grid = grid.expand(x.shape[0], -1, -1, -1)
    mode = {0: "bilinear", 1: "nearest", 2: "bicubic"}[interp]
    _ops.launch_count += 1 + (1 if m.shape[0] >= 2 else 0)  # the kernel (+ the one-launch prelude the device would have run)
    if pad == 3:
        return R.fill_and_sample(x, grid, mode, align, fill.to(x).reshape(-1)).contiguous()
    return F.grid_sample(x, grid, mode=mode, padding_mode={0: "zeros", 1: "border", 2: "reflection"}[pad], align_corners=align).contiguous()


def _undistort_mirror(image, lens, normalize):
    if image.shape[-1] not in (1, 3) or image.shape[2] % 4 != 0:
        raise _lib.Unsupported("outside the tiled kernel's envelope")
    cam = torch.zeros(lens.shape[0], 3, 3)
    cam[:, 0, 0], cam[:, 1, 1], cam[:, 0, 2], cam[:, 1, 2], cam[:, 2, 2] = lens[:, 0], lens[:, 1], lens[:, 2], lens[:, 3], 1.0
    _ops.launch_count += 1
    return R.undistort_image(_levels(image, normalize), cam, lens[:, 4:].contiguous()).contiguous()


def _unsupported(*a, **k):
    raise _lib.Unsupported("dry run")


@pytest.fixture()
def on_cpu(monkeypatch):
    monkeypatch.setattr(_ops, "warp_u8hwc", _warp_mirror)
    monkeypatch.setattr(_ops, "undistort_u8hwc", _undistort_mirror)
    monkeypatch.setattr(_ops, "undistort_fused", _unsupported)
    monkeypatch.setattr(undistort, "remap", R.remap)
    for name in ("warp_perspective", "warp_affine", "get_rotation_matrix2d"):
        monkeypatch.setattr(T.KT, name, getattr(R, name))
    monkeypatch.setattr(T, "DEV", "cpu")
    config.set("torch_prelude", 1)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    true_div = torch.Tensor.__truediv__

    def cuda_div(self, other):  # what `tensor / 255.0` is on a CUDA device
        if isinstance(other, float) and other == 255.0 and self.dtype == torch.float32:
            return self * RCP
        return true_div(self, other)

    monkeypatch.setattr(torch.Tensor, "__truediv__", cuda_div)
    return monkeypatch


def test_golden_warp_cases(on_cpu):
    for name in T.WARPS:
        T.test_ingest_matches_reference(name)
        T.test_ingest_is_bit_identical_to_the_three_steps_on_device(name)


def test_golden_undistort_cases(on_cpu):
    for name in T.UNDISTORT:
        T.test_undistort_from_bytes_matches_reference_and_the_fp32_path(name)


def test_property_tests(on_cpu):
    import bench

    torch.manual_seed(0)
    on_cpu.setattr(bench, "H_IMG", 1080)
    on_cpu.setattr(bench, "W_IMG", 1920)
    T.test_ingest_full_size_properties()
    T.test_ingest_headline_homographies_match_the_fp32_path()
    T.test_undistort_from_bytes_full_size()
    T.test_tiled_ingest_kernel_is_bit_identical_to_the_per_tap_kernel(on_cpu, "reflection", 3)
    T.test_tiled_ingest_kernel_is_bit_identical_to_the_per_tap_kernel(on_cpu, "fill", 3)
    T.test_tiled_ingest_kernel_is_bit_identical_to_the_per_tap_kernel(on_cpu, "fill", 1)
    T.test_tiled_ingest_kernel_is_bit_identical_to_the_per_tap_kernel(on_cpu, "border", 4)
    T.test_tiled_ingest_kernel_is_bit_identical_to_the_per_tap_kernel(on_cpu, "fill", 4)
