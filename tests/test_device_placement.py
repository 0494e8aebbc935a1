"""Device placement of the host compositions, checked without a GPU: every golden case is run on the ``meta`` device
through the real autograd bridges of kornia_b200/_ops.py with the C calls stubbed out.  ``meta`` has CUDA's mixing
rules (a non-scalar CPU tensor meeting a device tensor raises "Expected all tensors to be on the same device"), so a
tap table, grid or matrix built on the default device instead of the input's -- the kind of bug the CPU suite, which
swaps the core functions for the oracle, cannot see -- fails here; shapes and dtypes of outputs and gradients are compared
with the golden vectors.  No arithmetic runs: this says nothing about values (the ``-m gpu`` tests do)."""
import contextlib

import pytest
import torch

import kornia_b200 as K
from kornia_b200 import _lib, _ops
from conftest import golden
from helpers import family_grads, run_family_case

CASES = [(g, n) for g in ("family", "wider", "ingest") for n in golden(g).names()]
# host logic that reads a value back from the device (a crop size, the reference's own `.item()`): meta cannot answer;
# these ran on hardware in round 1 (tests/test_family_gpu.py)
READS_BACK = ("crop_", "center_crop")


class _Workspace:
    def __getattr__(self, name):
        return lambda *a: 64


@pytest.fixture()
def stubbed_device(monkeypatch):
    monkeypatch.setattr(_lib, "call", lambda name, *a: None)
    monkeypatch.setattr(_lib, "last_warp_launches", lambda: 1)
    monkeypatch.setattr(_lib, "load", lambda: _Workspace())
    monkeypatch.setattr(_ops, "_require_cuda", lambda t, what: None)

    def device_pointer(t):  # every pointer handed to the C ABI must be device memory
        assert t is None or t.is_meta, f"host tensor {tuple(t.shape)} passed to a CUDA entry point"
        return 0

    monkeypatch.setattr(_ops, "_ptr", device_pointer)
    monkeypatch.setattr(_ops, "_stream", lambda t: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    # `if torch.any(tilt != 0):` and `bool((sigma > 0).all())` are device->host reads in the reference too: answer them
    # the way the golden inputs would (no tilt terms is decided per case below, sigmas are positive)
    real_bool, real_item = torch.Tensor.__bool__, torch.Tensor.item
    answer = {"bool": True}
    monkeypatch.setattr(torch.Tensor, "__bool__", lambda self: answer["bool"] if self.is_meta else real_bool(self))
    monkeypatch.setattr(torch.Tensor, "item", lambda self: answer["bool"] if self.is_meta and self.dtype == torch.bool else real_item(self))
    return answer


def _module(op):
    base = op[:-5] if op.endswith("_grad") else op
    for m in (K.geometry.calibration, K.geometry.transform, K.filters, K.losses, K.metrics, K.geometry):
        if hasattr(m, base):
            return m
    raise KeyError(base)


@pytest.mark.parametrize("gname,name", CASES)
def test_composition_stays_on_the_input_device(stubbed_device, gname, name):
    if name.startswith(READS_BACK):
        pytest.skip("reads a size back from the device")
    op, kw, ins, outs = golden(gname).case(name)
    dist = ins.get("dist")
    tilt = dist is not None and dist.shape[-1] == 14 and bool((dist[..., 12:] != 0).any())
    stubbed_device["bool"] = tilt if op.startswith(("undistort", "distort")) else True
    if op.endswith("_grad"):
        got = family_grads(_module(op), op, kw, ins, outs, device="meta")
        pairs = [(got[k], w) for k, w in outs.items() if k != "cot"]
    else:
        out = run_family_case(_module(op), op, kw, ins, device="meta")
        if isinstance(out, (list, tuple)):
            pairs = [(g, outs[f"out{i}"]) for i, g in enumerate(out)]
        else:
            pairs = [(out, outs["out"])]
    for g, w in pairs:
        assert g.device.type == "meta" and g.shape == w.shape and g.dtype == w.dtype
