"""Host logic of kornia_b200/streaming.py that can be checked without a GPU: NUMA lookup against a fake sysfs tree, the
cpulist parser, the ring arithmetic; and (GPU) the pipeline against the direct call."""
import os

import pytest
import torch

from kornia_b200 import streaming as S


def test_cpulist_parser():
    assert S.parse_cpulist("0-3,8-9,12") == [0, 1, 2, 3, 8, 9, 12]
    assert S.parse_cpulist("0-31,64-95")[-1] == 95 and len(S.parse_cpulist("0-31,64-95")) == 64
    assert S.parse_cpulist("") == []


def test_numa_node_lookup_in_a_fake_sysfs(tmp_path, monkeypatch):
    bus = "0000:1b:00.0"
    (tmp_path / "bus/pci/devices" / bus).mkdir(parents=True)
    (tmp_path / "bus/pci/devices" / bus / "numa_node").write_text("1\n")
    (tmp_path / "devices/system/node/node1").mkdir(parents=True)
    (tmp_path / "devices/system/node/node1/cpulist").write_text("32-63,96-127\n")
    monkeypatch.setattr(S, "pci_bus_id", lambda i: bus)
    assert S.device_numa_node(0, sysfs=str(tmp_path)) == 1
    assert len(S.node_cpus(1, sysfs=str(tmp_path))) == 64
    (tmp_path / "bus/pci/devices" / bus / "numa_node").write_text("-1\n")  # what a single-node box or a VM reports
    assert S.device_numa_node(0, sysfs=str(tmp_path)) is None
    monkeypatch.setattr(S, "pci_bus_id", lambda i: None)
    assert S.device_numa_node(0, sysfs=str(tmp_path)) is None


def test_binding_is_a_no_op_when_the_platform_gives_no_answer(monkeypatch):
    monkeypatch.setattr(S, "pci_bus_id", lambda i: None)
    S._BOUND.clear()
    before = os.sched_getaffinity(0)
    info = S.bind_to_device_numa_node(5)
    assert info == {"node": None, "cpus": 0, "affinity": False, "mempolicy": False}
    assert os.sched_getaffinity(0) == before
    S._BOUND.clear()


def test_ring_arithmetic():
    per = 2 * 3 * 1080 * 1920 * 4
    assert S.host_ring_samples(256, 16, per, available_bytes=3 * 256 * per, local_ranks=1) == 256
    n = S.host_ring_samples(256, 16, per, available_bytes=64 << 30, local_ranks=8)
    assert n % 16 == 0 and 16 <= n < 256 and n * per * 3 * 8 <= (64 << 30)
    assert S.host_ring_samples(8, 16, per, available_bytes=1 << 30, local_ranks=8) == 8


def test_pipeline_refuses_a_cpu_device_and_device_inputs():
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.HostPipeline("cpu")


@pytest.mark.gpu
def test_host_pipeline_equals_the_direct_call():
    import kornia_b200 as K

    g = torch.Generator().manual_seed(0)
    B = 37  # not a multiple of the chunk: the last chunk is ragged
    src = torch.rand(B, 3, 48, 80, generator=g)
    M = torch.eye(3)[None].repeat(B, 1, 1) + 0.01 * torch.randn(B, 3, 3, generator=g) * torch.tensor([[1, 1, 30.0], [1, 1, 30.0], [1e-3, 1e-3, 0]])
    want = K.warp_perspective(src.cuda(), M.cuda(), (40, 64)).cpu()
    pinned = S.pinned_empty(src.shape, torch.float32, 0).copy_(src)
    got = S.warp_perspective_host(pinned, M, (40, 64), chunk=8)
    assert got.is_pinned() and torch.equal(got, want)
    again = S.warp_perspective_host(src, M, (40, 64), chunk=16, out=torch.empty_like(want))  # pageable host memory works too
    assert torch.equal(again, want)
    # a pinned ring shorter than the logical batch: chunks cycle through it
    ring = S.pinned_empty((16, 3, 48, 80), torch.float32, 0).copy_(src[:16])
    out = S.pinned_empty((16, 3, 40, 64), torch.float32, 0)
    S.warp_perspective_host(ring, M[:16].repeat(3, 1, 1), (40, 64), out=out, chunk=8, logical_batch=48)
    assert torch.equal(out, want[:16])
    # back-to-back passes without joining the caller's stream in between (what a stream of batches does), joined once at the end
    outs = [S.pinned_empty((B, 3, 40, 64), torch.float32, 0) for _ in range(3)]
    for o in outs:
        S.warp_perspective_host(pinned, M, (40, 64), out=o, chunk=8, synchronize=False, join=False)
    S.join("cuda", 8)
    torch.cuda.current_stream().synchronize()
    assert all(torch.equal(o, want) for o in outs)
    # decoder bytes in: 3 B/pixel over PCIe, converted and warped in one kernel
    frames = (torch.rand(B, 48, 80, 3, generator=g) * 255).to(torch.uint8)
    ref = K.geometry.transform.warp_perspective_from_uint8(frames.cuda(), M.cuda(), (40, 64)).cpu()
    assert torch.equal(S.warp_perspective_host(frames, M, (40, 64), chunk=8), ref)
