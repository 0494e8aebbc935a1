"""Shared helpers for the parity tests: dispatch a golden case onto an implementation module."""
import torch


def _tup(v):
    return tuple(v) if isinstance(v, list) else v


def run_case(impl, op, kwargs, ins, device=None, dtype=None):
    """Call ``impl.<op>`` (oracle or product: same signatures) on a golden case; returns the output.
    For ``*_grad`` ops returns dict(out=..., grad_<name>=...)."""
    kw = dict(kwargs)
    t = {}
    for k, v in ins.items():
        if dtype is not None and v.is_floating_point():
            v = v.to(dtype)
        t[k] = v.to(device) if device is not None else v
    base = op[:-5] if op.endswith("_grad") else op
    cot = t.pop("cot", None)

    def call(tt):
        if base in ("warp_perspective", "warp_affine"):
            return getattr(impl, base)(tt["src"], tt["M"], _tup(kw["dsize"]), mode=kw["mode"], padding_mode=kw["padding_mode"],
                                       align_corners=kw["align_corners"], fill_value=tt.get("fill_value"))
        if base == "remap":
            return impl.remap(tt["image"], tt["map_x"], tt["map_y"], **kw)
        if base == "filter2d":
            return impl.filter2d(tt["input"], tt["kernel"], **kw)
        if base == "filter2d_separable":
            return impl.filter2d_separable(tt["input"], tt["kernel_x"], tt["kernel_y"], **kw)
        if base == "gaussian_blur2d":
            sigma = tt["sigma"] if "sigma" in tt else _tup(kw["sigma"])
            return impl.gaussian_blur2d(tt["input"], _tup(kw["kernel_size"]), sigma, kw["border_type"], kw["separable"])
        raise KeyError(op)

    if not op.endswith("_grad"):
        return call(t)
    wrt = [k for k in t if k != "fill_value"]
    leaves = dict(t)
    for k in wrt:
        leaves[k] = t[k].clone().requires_grad_(True)
    out = call(leaves)
    grads = torch.autograd.grad((out * cot).sum(), [leaves[k] for k in wrt], allow_unused=True)
    res = {"out": out.detach()}
    for k, g in zip(wrt, grads):
        res[f"grad_{k}"] = g if g is not None else torch.zeros_like(t[k])
    return res


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
