#!/usr/bin/env python
"""bench.py -- headline benchmark of the warp / filter engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one ``warp_perspective`` call on a synthetic B x 3 x 1080 x 1920 fp32 batch
(BASELINE.json configs[1]: bilinear, zeros padding, align_corners=True), B per GPU fixed (weak
scaling: the batch dimension shards with no data-path collective).

One JSON line on stdout (rank 0):
  value     Mpix/s, whole job, inputs resident in HBM, through the public Python API
            (prelude + kernel), CUDA events around exactly K steps, max over ranks
  roofline  the fused warp kernel alone: CUDA events around each launch inside the timed region,
            algorithmic bytes = 24 B/pixel (read 3 fp32 + write 3 fp32)
  e2e       same metric with HOST buffers: pinned src -> H2D -> kernel -> D2H of the full output,
            chunked and pipelined over three streams, copies inside the timed region
  cpu_baseline  the oracle's torch-op port of the reference composition on the host cores,
            bounded sample (rank 0, N=1 only)
``--impl reference`` times that CPU port alone (the reference is pure Python and cannot travel to
the GPU box; the port issues the same ATen calls: oracle/kornia_restated.py).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H_IMG, W_IMG, C_IMG = 1080, 1920, 3
BYTES_PER_PIX = 24.0  # algorithmic: 3 channels x 4 B read + 3 x 4 B written per output pixel
METRIC = "Mpix/s warp_perspective Bx3x1080x1920 fwd bilinear fp32"


# ------------------------------------------------------------------------------------------ inputs
def perspective_from_quads(src_q: torch.Tensor, dst_q: torch.Tensor) -> torch.Tensor:
    """DLT: the (B,3,3) homography mapping 4 source corners to 4 destination corners (what
    kornia.geometry.get_perspective_transform returns; used by the reference's flagship benchmark,
    benchmarks/geometry/flagship.py:101-107).  Solved in float64 on the host."""
    s, d = src_q.double(), dst_q.double()
    B = s.shape[0]
    A = torch.zeros(B, 8, 8, dtype=torch.float64)
    b = torch.zeros(B, 8, dtype=torch.float64)
    for i in range(4):
        x, y, u, v = s[:, i, 0], s[:, i, 1], d[:, i, 0], d[:, i, 1]
        A[:, 2 * i, 0], A[:, 2 * i, 1], A[:, 2 * i, 2] = x, y, 1.0
        A[:, 2 * i, 6], A[:, 2 * i, 7] = -u * x, -u * y
        A[:, 2 * i + 1, 3], A[:, 2 * i + 1, 4], A[:, 2 * i + 1, 5] = x, y, 1.0
        A[:, 2 * i + 1, 6], A[:, 2 * i + 1, 7] = -v * x, -v * y
        b[:, 2 * i], b[:, 2 * i + 1] = u, v
    hvec = torch.linalg.solve(A, b)
    return torch.cat([hvec, torch.ones(B, 1, dtype=torch.float64)], 1).view(B, 3, 3).float()


def make_homographies(B: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    quad = torch.tensor([[0.0, 0.0], [W_IMG - 1.0, 0.0], [W_IMG - 1.0, H_IMG - 1.0], [0.0, H_IMG - 1.0]]).expand(B, 4, 2)
    return perspective_from_quads(quad, quad + 8.0 * torch.randn(B, 4, 2, generator=g))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        return False

    def hold_load(self, fn, sync, min_samples: int = 2, max_seconds: float = 1.0) -> int:
        """A timed region of a few tens of milliseconds can end before nvidia-smi's first 100 ms report.  Keep the very
        same load running -- untimed, after the stop event -- until a few samples exist, so that the clocks line
        describes the GPU under this load.  Returns the number of extra (untimed) steps."""
        self.extra = 0
        if self.proc is None:
            return 0
        deadline = time.time() + max_seconds
        try:
            while len(self.rows) < min_samples and time.time() < deadline:
                for _ in range(4):
                    fn()
                sync()
                self.extra += 4
        except Exception:
            pass
        return self.extra

    def summary(self):
        out = self._summary()
        out["untimed_steps_for_sampling"] = getattr(self, "extra", 0)
        return out

    def _summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU port
def cpu_reference_run(steps: int, warmup: int, sample_b: int):
    """Time the oracle's torch-op port of the reference on the host cores.  The thread count that serves
    the reference best is picked by a one-shot calibration (ATen's CPU sampler parallelises over the batch,
    the elementwise ops over elements; 128 threads on an 8-image batch oversubscribe badly)."""
    from oracle import kornia_restated as R

    cores = os.cpu_count() or 1
    g = torch.Generator().manual_seed(0)
    src = torch.rand(sample_b, C_IMG, H_IMG, W_IMG, generator=g)
    M = make_homographies(sample_b, 0)

    def once():
        t0 = time.perf_counter()
        R.warp_perspective(src, M, (H_IMG, W_IMG))
        return time.perf_counter() - t0

    best_t, best_n = None, cores
    for n in sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, sample_b)}, reverse=True):
        torch.set_num_threads(n)
        once()
        t = once()
        if best_t is None or t < best_t:
            best_t, best_n = t, n
    torch.set_num_threads(best_n)
    for _ in range(warmup):
        once()
    dt = sum(once() for _ in range(steps)) / steps
    mpix = sample_b * H_IMG * W_IMG / dt / 1e6
    return mpix, dt * 1e3, best_n


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 32
    mpix, ms, cores = cpu_reference_run(args.steps, max(args.warmup, 1), sample_b)
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"warp_perspective fwd B={args.batch}x3x1080x1920 bilinear zeros align_corners=True",
                   "per_step_sample": f"B={sample_b} of the batch (CPU per-image throughput is batch independent)"},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": f"B={sample_b}x3x1080x1920 per step, torch CPU ops (oracle/kornia_restated.py), {cores} of {os.cpu_count()} threads (best of a calibration)"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ ours
def host_ring_samples(B: int, chunk: int, available_bytes=None, local_ranks=None) -> int:
    """How many samples of the batch each rank keeps in pinned host memory (source and destination each); the arithmetic
    lives in the library (kornia_b200/streaming.py:host_ring_samples)."""
    from kornia_b200.streaming import host_ring_samples as ring

    return ring(B, chunk, 2 * C_IMG * H_IMG * W_IMG * 4, available_bytes, local_ranks)


def e2e_run(K, M_dev, B, steps, warmup, chunk, dev):
    """Host-buffer throughput through the library's own host pipeline (kornia_b200.streaming.warp_perspective_host: pinned
    src -> device -> warp -> pinned dst, chunked over 3 streams, the process bound to the GPU's NUMA node)."""
    from kornia_b200 import streaming

    n_el = B * C_IMG * H_IMG * W_IMG
    numa = streaming.bind_to_device_numa_node(dev.index)  # before the pinned allocations: their pages follow the policy
    # Host side of the step: the whole batch in pinned memory (2 x 6.37 GB per rank at B=256).  When the box cannot
    # spare that for every local rank (8 ranks would lock 102 GB), the batch is streamed through a shorter pinned ring
    # of whole chunks instead: the bytes crossing PCIe per step are the same, the note says which form ran.
    HB = host_ring_samples(B, chunk)
    try:
        src_h = streaming.pinned_empty((HB, C_IMG, H_IMG, W_IMG), torch.float32, dev.index)
        dst_h = streaming.pinned_empty((HB, C_IMG, H_IMG, W_IMG), torch.float32, dev.index)
    except RuntimeError as e:  # not enough lockable host memory
        return None, f"pinned allocation failed: {e}"
    # cheap deterministic fill (content does not affect timing); touching every page also places it
    src_h.view(-1)[: 1 << 20].uniform_()
    src_h.view(-1)[1 << 20:] = 0.5
    dst_h.zero_()

    def one_step():
        streaming.warp_perspective_host(src_h, M_dev, (H_IMG, W_IMG), out=dst_h, device=dev, chunk=chunk, logical_batch=B, synchronize=False)

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize(dev)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        one_step()
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1) / steps
    del src_h, dst_h
    host = "whole batch pinned" if HB == B else f"pinned ring of {HB} samples reused {B / HB:.1f}x per step (host memory per local rank)"
    return ms, (f"kornia_b200.streaming.warp_perspective_host: pinned host buffers ({host}), chunk={chunk} samples, 3 streams (H2D / kernel / D2H), "
                f"{n_el * 4} B each way per step; NUMA binding {numa}")


def max_over_ranks_or_none(dist, ms, note, device):
    """Max over ranks of a per-rank time that some ranks may not have (None, e.g. a pinned allocation that failed): EVERY rank
    takes part in the one all_reduce -- a missing value travels as +inf -- and all ranks return the same (ms | None, note)."""
    t = torch.tensor([ms if ms is not None else float("inf")], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = float(t.item())
    if math.isfinite(worst):
        return worst, note
    return None, note if ms is None else "unavailable on another rank"


def run_ours(args) -> None:
    import kornia_b200 as K
    from kornia_b200 import _lib, _ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; kornia_b200 has no CPU path (use --impl reference for the CPU port)")
    _lib.load()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    # per-rank shard generated in place: no scatter needed, seed = 1000 + rank
    torch.manual_seed(1000 + rank)
    src = torch.rand(B, C_IMG, H_IMG, W_IMG, device=dev)
    M = make_homographies(B, 1000 + rank).to(dev)
    dsize = (H_IMG, W_IMG)

    def barrier():
        if dist is not None:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        out = K.warp_perspective(src, M, dsize)
    variant = _lib.last_warp_variant()
    barrier()
    _ops.kernel_events = []
    launches0 = _ops.launch_count
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        t0.record()
        for _ in range(args.steps):
            out = K.warp_perspective(src, M, dsize)
        t1.record()
        barrier()
        timed_events, _ops.kernel_events = _ops.kernel_events, None
        launches = _ops.launch_count - launches0
        clk.hold_load(lambda: K.warp_perspective(src, M, dsize), lambda: torch.cuda.synchronize(dev))
    total_ms = t0.elapsed_time(t1)
    kern_ms = [s.elapsed_time(e) for (_, s, e) in timed_events]
    checksum = float(out[0, :, ::97, ::89].sum())  # touch the result
    del out
    if dist is not None:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_step = total_ms / args.steps
    pix_step = world * B * H_IMG * W_IMG
    value = pix_step / (ms_step * 1e-3) / 1e6

    # ---------------------------------------------------------------- e2e (host buffers)
    e2e_ms, e2e_note = e2e_run(K, M, B, steps=max(2, min(args.steps, 3)), warmup=1, chunk=args.e2e_chunk, dev=dev)
    if dist is not None:
        e2e_ms, e2e_note = max_over_ranks_or_none(dist, e2e_ms, e2e_note, dev)
    barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak = float(peaks["hbm_gbs"])
        peak_src = "measured MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        peak = 6650.0
    k_ms = statistics.mean(kern_ms) if kern_ms else float("nan")
    achieved = BYTES_PER_PIX * B * H_IMG * W_IMG / (k_ms * 1e-3) / 1e9
    traffic = None
    try:  # per-launch DRAM bytes of the same kernel from the committed ncu capture, if any
        traffic = json.load(open(os.path.join(ROOT, "profiles", "warp_fwd_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        mpix, ms, cores = cpu_reference_run(steps=3, warmup=1, sample_b=32)
        cpu = {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": f"3 steps of B=32x3x1080x1920 ({ms:.0f} ms each) with torch CPU ops, {cores} of {os.cpu_count()} threads "
                         "(best of a calibration): oracle/kornia_restated.py"}

    bytes_step = B * C_IMG * H_IMG * W_IMG * 4
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"warp_perspective fwd B={B}x3x1080x1920 per GPU, bilinear, zeros, align_corners=True (BASELINE.json configs[1])",
                   "global_batch": B * world, "parallelism": f"batch-sharded x{world}, no data-path collective",
                   "l2": "inputs (6.37 GB/GPU) exceed the 126 MB L2; no explicit flush", "kernel_variant": variant,
                   "homographies": "corner quad jittered by 8*randn px (benchmarks/geometry/flagship.py recipe), seed 1000+rank"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel_ms": k_ms, "kernel_launches_timed": len(kern_ms), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_PIX * B * H_IMG * W_IMG},
        "cpu_baseline": cpu,
        "e2e": ({"value": pix_step / (e2e_ms * 1e-3) / 1e6, "unit": "Mpix/s", "h2d_bytes_per_step": bytes_step,
                 "d2h_bytes_per_step": bytes_step, "ms_per_step": e2e_ms, "note": e2e_note}
                if e2e_ms is not None else {"value": None, "unit": "Mpix/s", "note": e2e_note}),
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "checksum": checksum,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_extra(args) -> None:
    """Secondary BASELINE.json configs (not the driver's headline line): ``--workload blur`` = configs[2]
    gaussian_blur2d k=11 B=256x3x1080x1920; ``--workload warp_bwd`` = configs[3] warp_perspective fwd+bwd
    (grad wrt image and H) B=128x3x720x1280; ``--workload ingest`` = warp_perspective_from_uint8 on B x 1080 x 1920 x 3 decoder
    bytes (15 algorithmic bytes per pixel).  Single GPU, inputs resident, CUDA events, one JSON line."""
    import kornia_b200 as K
    from kornia_b200 import _lib, _ops

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    _lib.load()
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    torch.manual_seed(1000)
    if args.workload == "blur":
        B = args.batch
        x = torch.rand(B, 3, H_IMG, W_IMG, device=dev)
        step = lambda: K.gaussian_blur2d(x, (11, 11), (2.0, 2.0), "reflect", True)  # noqa: E731
        pix, bytes_per_pix = B * H_IMG * W_IMG, 24.0
        name = f"gaussian_blur2d k=11 sigma=2 reflect separable B={B}x3x1080x1920"
        tag = "sepfilter_forward"
    elif args.workload == "ingest":
        B = args.batch
        frames = torch.randint(0, 256, (B, H_IMG, W_IMG, 3), device=dev, dtype=torch.uint8)
        M = make_homographies(B, 1000).to(dev)
        step = lambda: K.geometry.transform.warp_perspective_from_uint8(frames, M, (H_IMG, W_IMG))  # noqa: E731
        pix, bytes_per_pix = B * H_IMG * W_IMG, 15.0
        name = f"warp_perspective_from_uint8 (decoder bytes HWC -> warped fp32 NCHW) B={B}x1080x1920x3, bilinear, zeros"
        tag = "warp_u8hwc_forward"
    else:
        B, Hh, Ww = min(args.batch, 128), 720, 1280
        src = torch.rand(B, 3, Hh, Ww, device=dev, requires_grad=True)
        g = torch.Generator().manual_seed(7)
        quad = torch.tensor([[0.0, 0.0], [Ww - 1.0, 0.0], [Ww - 1.0, Hh - 1.0], [0.0, Hh - 1.0]]).expand(B, 4, 2)
        M = perspective_from_quads(quad, quad + 8.0 * torch.randn(B, 4, 2, generator=g)).to(dev).requires_grad_(True)
        cot = torch.rand(B, 3, Hh, Ww, device=dev) - 0.5  # fixed upstream gradient: no loss glue in the timed region

        def step():
            out = K.warp_perspective(src, M, (Hh, Ww))
            return torch.autograd.grad(out, [src, M], grad_outputs=cot)

        pix, bytes_per_pix = B * Hh * Ww, 60.0
        name = f"warp_perspective fwd+bwd (d/dsrc, d/dM), fixed cotangent, B={B}x3x720x1280"
        tag = "warp_backward"
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    _ops.kernel_events = [] if tag else None
    launches0 = _ops.launch_count
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index or 0) as clk:
        t0.record()
        for _ in range(args.steps):
            step()
        t1.record()
        torch.cuda.synchronize(dev)
        timed_events, _ops.kernel_events = (_ops.kernel_events or []), None
        launches = _ops.launch_count - launches0
        clk.hold_load(step, lambda: torch.cuda.synchronize(dev))
    ms = t0.elapsed_time(t1) / args.steps
    kern = [s.elapsed_time(e) for (tg, s, e) in timed_events if tg == tag]
    k_ms = statistics.mean(kern) if kern else None
    line = {"metric": "Mpix/s " + name, "value": pix / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "note": "whole step through the public API (for warp_bwd: forward, zero-fill of d/dsrc, backward, d/dM reduction, torch prelude autograd)"},
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes_per_pixel": bytes_per_pix,
                         "achieved_step": bytes_per_pix * pix / (ms * 1e-3) / 1e9, "frac_step": bytes_per_pix * pix / (ms * 1e-3) / 1e9 / peak,
                         "kernel_ms": k_ms, "achieved": (bytes_per_pix * pix / (k_ms * 1e-3) / 1e9) if k_ms else None,
                         "frac": (bytes_per_pix * pix / (k_ms * 1e-3) / 1e9 / peak) if k_ms else None},
            "gpu_launches": launches, "clocks": clk.summary()}
    if args.workload == "warp_bwd":  # the timed kernel is the backward alone: 36 B/pixel (read gout + src, write gsrc)
        line["roofline"].update({"kernel": "warp_backward (+ d/dM reduction)", "kernel_bytes_per_pixel": 36.0,
                                 "achieved": (36.0 * pix / (k_ms * 1e-3) / 1e9) if k_ms else None,
                                 "frac": (36.0 * pix / (k_ms * 1e-3) / 1e9 / peak) if k_ms else None})
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=256, help="samples per GPU")
    ap.add_argument("--e2e-chunk", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", choices=["warp", "blur", "warp_bwd", "ingest"], default="warp",
                    help="warp = the headline (BASELINE.json configs[1]); blur / warp_bwd = configs[2] / configs[3]; ingest = the uint8 wire-format warp "
                         "(SURVEY 8f row 4, not a BASELINE config); single GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "warp":
        run_extra(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
