"""The kernels as registered PyTorch operators (kornia_b200/_ops.py; SURVEY.md section 8b "C-ABI equivalent" row, the
reference's dynamo tests tests/geometry/transform/test_imgwarp.py:103,287,538).  CPU part: schemas, fake / meta kernels,
whole-graph capture with fake tensors, the CUDA-only contract, the inference-mode cache.  GPU part: torch.library.opcheck
on every operator and torch.compile(fullgraph=True) against eager."""
import pytest
import torch

import kornia_b200 as K
from kornia_b200 import _ops

ops = torch.ops.kornia_b200


def test_every_operator_is_registered_with_a_fake_kernel():
    for name in _ops.OPERATORS:
        op = getattr(ops, name).default
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "CUDA"), name
        assert torch._C._dispatch_has_kernel_for_dispatch_key(op.name(), "Meta") or torch._library.utils.has_fake_kernel(op), name
    for name in ("warp_fwd", "warp_prelude", "remap_fwd", "filter2d_fwd", "sepfilter_fwd", "spatial_gradient_fwd"):
        assert torch._C._dispatch_has_kernel_for_dispatch_key(getattr(ops, name).default.name(), "Autograd"), name


class _Pipeline(torch.nn.Module):
    def forward(self, src, M, mx, my):
        a = K.warp_perspective(src * 2, M, (20, 30)).sum() + K.gaussian_blur2d(src, (5, 5), (1.5, 1.5)).mean()
        b = K.remap(src, mx, my) + K.filter2d(src, torch.ones(1, 3, 3, device=src.device))
        c = K.warp_affine(src, M[:, :2], (16, 24), padding_mode="border")
        d = K.filters.sobel(src).mean() + K.filters.spatial_gradient(src).mean()
        return a + b.sum() + c.mean() + d


def test_public_functions_trace_as_one_graph():
    """torch.export runs the Python of the six functions under fake tensors: no graph break, the kernels appear as
    opaque kornia_b200::* nodes with the right output shapes."""
    src, M = torch.rand(2, 3, 16, 24), torch.eye(3)[None].repeat(2, 1, 1)
    ep = torch.export.export(_Pipeline(), (src, M, torch.rand(2, 16, 24), torch.rand(2, 16, 24)))
    seen = [str(n.target) for n in ep.graph.nodes if "kornia_b200" in str(n.target)]
    assert seen == ["kornia_b200.warp_fwd.default", "kornia_b200.sepfilter_fwd.default", "kornia_b200.remap_fwd.default",
                    "kornia_b200.filter2d_fwd.default", "kornia_b200.warp_fwd.default", "kornia_b200.spatial_gradient_fwd.default",
                    "kornia_b200.spatial_gradient_fwd.default"]
    shapes = [tuple(n.meta["val"].shape) for n in ep.graph.nodes if "kornia_b200" in str(n.target)]
    assert shapes[0] == (2, 3, 20, 30) and shapes[1] == (2, 3, 16, 24) and shapes[4] == (2, 3, 16, 24) and shapes[6] == (2, 3, 2, 16, 24)


def test_meta_tensors_run_forward_and_backward_through_the_operators():
    src = torch.rand(2, 3, 16, 24, device="meta", requires_grad=True)
    M = torch.eye(3, device="meta")[None].repeat(2, 1, 1).requires_grad_(True)
    out = K.warp_perspective(src, M, (20, 30))
    assert out.shape == (2, 3, 20, 30) and out.device.type == "meta" and out.grad_fn is not None
    gs, gm = torch.autograd.grad(out.sum(), [src, M])
    assert gs.shape == src.shape and gm.shape == M.shape
    kx = torch.rand(1, 5, device="meta", requires_grad=True)
    ky = torch.rand(1, 7, device="meta", requires_grad=True)
    y = K.filter2d_separable(src, kx, ky, padding="valid")
    assert y.shape == (2, 3, 10, 20)
    g = torch.autograd.grad(y.sum(), [src, kx, ky])
    assert [t.shape for t in g] == [src.shape, kx.shape, ky.shape]
    mx = torch.rand(1, 8, 9, device="meta", requires_grad=True)
    r = K.remap(src, mx, mx)
    assert torch.autograd.grad(r.sum(), [mx])[0].shape == (1, 8, 9)


def test_backward_operators_have_no_autograd_formula_so_double_backward_is_loud():
    src = torch.rand(1, 1, 8, 8, device="meta", requires_grad=True)
    M = torch.eye(3, device="meta")[None].requires_grad_(True)
    out = K.warp_perspective(src, M, (8, 8))
    (gs,) = torch.autograd.grad(out.sum(), [src], create_graph=True)
    with pytest.raises(RuntimeError, match="double backward"):
        torch.autograd.grad(gs.sum(), [M])


def test_cpu_tensors_raise_the_cuda_only_error():
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        K.warp_perspective(x, torch.eye(3)[None], (8, 8))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        K.gaussian_blur2d(x, (3, 3), (1.0, 1.0))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        K.remap(x, torch.zeros(1, 8, 8), torch.zeros(1, 8, 8))


def test_constants_cached_during_inference_can_be_saved_for_backward():
    """the base-grid axes and the Gaussian taps are cached per shape; an eval pass under inference_mode must not
    poison a later training step with inference tensors (save_for_backward rejects them)."""
    import importlib

    gaussian = importlib.import_module("kornia_b200.filters.gaussian")
    _prelude = importlib.import_module("kornia_b200.geometry._prelude")

    _prelude._AXES_CACHE.clear()
    gaussian._TAPS_CACHE.clear()
    M = torch.eye(3, device="meta")[None].repeat(2, 1, 1)
    with torch.inference_mode():
        K.warp_perspective(torch.rand(2, 3, 12, 12, device="meta"), M, (9, 11))
        K.warp_affine(torch.rand(2, 3, 12, 12, device="meta"), M[:, :2], (9, 11))
        K.gaussian_blur2d(torch.rand(2, 3, 12, 12, device="meta"), (3, 3), (1.0, 1.0))
    assert _prelude._AXES_CACHE and gaussian._TAPS_CACHE
    for t in list(_prelude._AXES_CACHE.values()) + list(gaussian._TAPS_CACHE.values()):
        assert not any(v.is_inference() for v in t)
    src = torch.rand(2, 3, 12, 12, device="meta", requires_grad=True)
    out = K.warp_perspective(src, M, (9, 11)).sum() + K.warp_affine(src, M[:, :2], (9, 11)).sum() + K.gaussian_blur2d(src, (3, 3), (1.0, 1.0)).sum()
    out.backward()
    assert src.grad.shape == src.shape


# ----------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _inputs(B=3, C=3, H=40, W=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.rand(B, C, H, W, generator=g).cuda()
    M = (torch.eye(3)[None] + 0.01 * torch.randn(B, 3, 3, generator=g) * torch.tensor([[1, 1, 30.0], [1, 1, 30.0], [1e-3, 1e-3, 0]])).cuda()
    return src, M


@gpu
def test_opcheck_core_operators():
    from torch.library import opcheck

    src, M = _inputs()
    src.requires_grad_(True)
    m = K.geometry._prelude.sampling_matrix(M, (40, 64), (32, 48), False).detach().requires_grad_(True)
    bx, by = K.geometry._prelude.meshgrid_axes(32, 48, src.device, src.dtype)
    opcheck(ops.warp_fwd, (src, m, bx, by, None, 32, 48, True, 0, 0, True))
    opcheck(ops.warp_fwd, (src, m, bx, by, torch.rand(3, device="cuda"), 32, 48, True, 0, 3, True))
    opcheck(ops.warp_bwd, (torch.rand(3, 3, 32, 48, device="cuda"), src.detach(), m.detach(), bx, by, None, 32, 48, True, 0, 0, True, True, True))
    opcheck(ops.warp_prelude, (M.clone().requires_grad_(True), 40, 64, 32, 48, False))
    opcheck(ops.warp_prelude, (M[:, :2].clone().requires_grad_(True), 40, 64, 32, 48, True))
    mx = (torch.rand(3, 32, 48, device="cuda") * 60).requires_grad_(True)
    my = (torch.rand(3, 32, 48, device="cuda") * 38).requires_grad_(True)
    opcheck(ops.remap_fwd, (src, mx, my, False, 0, 0, False))
    k = torch.rand(1, 3, 5, device="cuda", requires_grad=True)
    opcheck(ops.filter2d_fwd, (src, k, 1, True))
    opcheck(ops.filter2d_fwd, (src, k, 0, False))
    kx = torch.rand(1, 5, device="cuda", requires_grad=True)
    ky = torch.rand(3, 5, device="cuda", requires_grad=True)
    opcheck(ops.sepfilter_fwd, (src, kx, ky, 1, True))
    opcheck(ops.sepfilter_fwd, (src, kx, ky, 2, False))


@gpu
def test_opcheck_caller_operators():
    from torch.library import opcheck

    src, M = _inputs()
    import importlib

    taps, nout, k = importlib.import_module("kornia_b200.filters.sobel")._host_taps("sobel", 1, True, torch.float32)
    opcheck(ops.spatial_gradient_fwd, (src.clone().requires_grad_(True), list(taps), nout, k, False, 0.0))
    opcheck(ops.spatial_gradient_fwd, (src, list(taps), nout, k, True, 1e-6), test_utils=("test_schema", "test_faketensor"))
    kern = K.filters.get_gaussian_kernel1d(7, 1.5, device="cuda")
    opcheck(ops.ssim_fwd, (src, src.flip(0), kern, 1e-4, 9e-4, 1e-12))
    kx = K.filters.get_gaussian_kernel1d(5, 1.0, device="cuda")
    opcheck(ops.sepfilter_lerp_fwd, (src, kx, kx, 1, 2.0))
    c = torch.rand(4, 2, device="cuda") * 50
    opcheck(ops.rotation_matrix2d, (c, torch.rand(4, device="cuda") * 90, torch.ones(4, 2, device="cuda")))
    quad = torch.tensor([[0.0, 0], [63, 0], [63, 39], [0, 39]], device="cuda")[None].repeat(4, 1, 1)
    opcheck(ops.perspective_from_points, (quad, quad + torch.randn(4, 4, 2, device="cuda")))
    u8 = (torch.rand(3, 40, 64, 3, device="cuda") * 255).to(torch.uint8)
    m = K.geometry._prelude.sampling_matrix(M, (40, 64), (32, 48), False)
    bx, by = K.geometry._prelude.meshgrid_axes(32, 48, src.device, src.dtype)
    opcheck(ops.warp_u8hwc_fwd, (u8, m, bx, by, None, 32, 48, True, 0, 0, True, 1))


def _pipeline(src, M, mx, my):
    a = K.warp_perspective(src, M, (32, 48))
    b = K.gaussian_blur2d(a, (5, 5), (1.5, 1.5))
    c = K.remap(src, mx, my, align_corners=True)
    d = K.warp_affine(src, M[:, :2], (32, 48), padding_mode="border")
    return (a * b).sum() + c.mean() + (d * d).sum()


@gpu
@pytest.mark.parametrize("backend", ["aot_eager", "inductor"])
def test_torch_compile_fullgraph_matches_eager(backend):
    torch._dynamo.reset()
    src, M = _inputs()
    mx = torch.rand(3, 32, 48, device="cuda") * 60
    my = torch.rand(3, 32, 48, device="cuda") * 38

    def run(fn):
        s, m = src.clone().requires_grad_(True), M.clone().requires_grad_(True)
        out = fn(s, m, mx, my)
        gs, gm = torch.autograd.grad(out, [s, m])
        return out.detach(), gs, gm

    want = run(_pipeline)
    try:
        compiled = torch.compile(_pipeline, fullgraph=True, backend=backend)
        got = run(compiled)
    except Exception as e:  # an inductor toolchain problem on the box is not a defect of the operators: aot_eager must pass
        if backend == "inductor" and "kornia_b200" not in str(e):
            pytest.skip(f"inductor unavailable here: {type(e).__name__}: {str(e)[:200]}")
        raise
    # the kernels are the same launches in both modes (bit-identical); the torch glue around them may fuse differently
    # (d/dM are sums over ~1500 pixels of terms of both signs: the fused glue moves their last bits)
    for g, w, tol in zip(got, want, (1e-5, 1e-5, 2e-4)):
        torch.testing.assert_close(g, w, rtol=tol, atol=tol * max(1.0, float(w.abs().max())))


@gpu
def test_inference_then_training_on_the_device():
    src, M = _inputs(seed=3)
    with torch.inference_mode():
        K.warp_perspective(src, M, (24, 40))
        K.gaussian_blur2d(src, (7, 7), (1.2, 1.2))
    s = src.clone().requires_grad_(True)
    (K.warp_perspective(s, M, (24, 40)).sum() + K.gaussian_blur2d(s, (7, 7), (1.2, 1.2)).sum()).backward()
    assert torch.isfinite(s.grad).all()


@gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_images_pass_through(dtype):
    """fp16 / bf16 are in the reference's test matrix (testing/base.py:35-36) though in no BASELINE configuration: the kernels
    compute in fp32 and hand back the image's dtype.  Against the reference composition in the same half precision on the same
    GPU (its coordinates are rounded to half precision, so the comparison is half-grade) and against this library's fp32 result."""
    from oracle import kornia_restated as R

    g = torch.Generator().manual_seed(5)
    src32 = torch.rand(2, 3, 24, 32, generator=g).cuda()
    M32 = (torch.eye(3)[None] + 0.02 * torch.randn(2, 3, 3, generator=g) * torch.tensor([[1, 1, 5.0], [1, 1, 5.0], [1e-3, 1e-3, 0]])).cuda()
    src, M = src32.to(dtype), M32.to(dtype)
    tol = dict(rtol=0, atol=0.02 if dtype == torch.float16 else 0.25)  # the bf16 reference rounds COORDINATES to 8 bits
    out = K.warp_perspective(src, M, (20, 28))
    assert out.dtype == dtype and out.shape == (2, 3, 20, 28)
    torch.testing.assert_close(out.float(), K.warp_perspective(src.float(), M.float(), (20, 28)), **tol)
    torch.testing.assert_close(out.float(), R.warp_perspective(src, M, (20, 28)).float(), **tol)
    blur = K.gaussian_blur2d(src, (5, 5), (1.2, 1.2))
    assert blur.dtype == dtype
    torch.testing.assert_close(blur.float(), K.gaussian_blur2d(src.float(), (5, 5), (1.2, 1.2)), rtol=0, atol=0.01)
    k = torch.rand(1, 3, 3, device="cuda", dtype=dtype)
    assert K.filter2d(src, k).dtype == dtype
    mx = (torch.rand(2, 20, 28, device="cuda") * 30).to(dtype)
    assert K.remap(src, mx, mx * 0.7).dtype == dtype
    s = src.clone().requires_grad_(True)
    K.warp_perspective(s, M, (20, 28)).float().sum().backward()
    assert s.grad is not None and s.grad.dtype == dtype


@gpu
def test_cuda_graph_replay_of_a_small_warp_and_blur_chain():
    """kornia_b200.graphs.GraphedCall: prelude + both warp launches + the blur capture into one CUDA graph (no allocation inside
    the C ABI, launches on the capturing stream) and replay bit-identically on new data."""
    g = torch.Generator().manual_seed(11)
    x = torch.rand(8, 3, 96, 128, generator=g).cuda()
    M = (torch.eye(3)[None] + 0.01 * torch.randn(8, 3, 3, generator=g) * torch.tensor([[1, 1, 20.0], [1, 1, 20.0], [1e-3, 1e-3, 0]])).cuda()

    def chain(img, H):
        return K.gaussian_blur2d(K.warp_perspective(img, H, (96, 128)), (5, 5), (1.2, 1.2))

    graphed = K.graphs.GraphedCall(chain, x, M)
    for seed in (1, 2):
        g2 = torch.Generator().manual_seed(seed)
        x2 = torch.rand(8, 3, 96, 128, generator=g2).cuda()
        M2 = M.flip(0).contiguous()
        assert torch.equal(graphed(x2, M2), chain(x2, M2))
    with pytest.raises(RuntimeError, match="captured for"):
        graphed(x[:4], M[:4])
