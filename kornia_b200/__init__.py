"""kornia_b200 -- B200-native (sm_100a) drop-in for Kornia's warp / filter hot path.

    kornia.geometry.transform.{warp_perspective, warp_affine, remap}
    kornia.filters.{filter2d, filter2d_separable, gaussian_blur2d}

Same function names, signatures, defaults, validation and autograd behaviour; the image work runs
in hand-written CUDA kernels behind the C ABI of ``include/kornia_b200.h``.  CUDA-only: there is
no CPU path and no fallback -- a missing library or a CPU tensor raises.
"""
from __future__ import annotations

from . import augmentation, config, core, filters, geometry, graphs, losses, metrics, streaming
from .filters import filter2d, filter2d_separable, gaussian_blur2d
from .geometry.transform import remap, warp_affine, warp_perspective

__version__ = "0.1.0"

_PATCHED = {}


def install(kornia_module=None) -> None:
    """Rebind the hot-path functions (and the four callers that have a fused kernel of their own here:
    ``get_perspective_transform``, ``spatial_gradient``, ``sobel``, ``ssim``) on an imported ``kornia`` package.

    Every module of the package that holds a reference to one of the originals is patched: the
    defining modules and re-export sites (geometry/transform/__init__.py:30, geometry/__init__.py:40,
    filters/__init__.py:37-38) and every ``from ... import warp_perspective``-style importer, so the
    callers either side of the path -- ``RandomPerspective`` / ``RandomAffine`` / ``RandomGaussianBlur``
    (augmentation/_2d/geometric/perspective.py:108, affine.py:154, _2d/intensity/gaussian_blur.py:108),
    ``affine`` / ``rotate`` / ``crop_by_transform_mat``, ``GaussianBlur2d``, ``unsharp_mask``, SSIM ... --
    run on the CUDA kernels without being rewritten.  ``uninstall()`` restores the originals."""
    import importlib
    import sys

    k = kornia_module or importlib.import_module("kornia")
    prefix = k.__name__ + "."
    defining = {
        "warp_perspective": ("geometry.transform.imgwarp", warp_perspective),
        "warp_affine": ("geometry.transform.imgwarp", warp_affine),
        "remap": ("geometry.transform.imgwarp", remap),
        "filter2d": ("filters.filter", filter2d),
        "filter2d_separable": ("filters.filter", filter2d_separable),
        "gaussian_blur2d": ("filters.gaussian", gaussian_blur2d),
        # callers with a kernel of their own here (the rest of the family reaches the kernels through the six above)
        "get_perspective_transform": ("geometry.transform.imgwarp", geometry.transform.get_perspective_transform),
        "spatial_gradient": ("filters.sobel", filters.spatial_gradient),
        "sobel": ("filters.sobel", filters.sobel),
        "ssim": ("metrics.ssim", metrics.ssim),
    }
    originals = {}
    for name, (mod, fn) in defining.items():
        m = importlib.import_module(prefix + mod)
        originals[name] = (_PATCHED.get((m, name), getattr(m, name)), fn)
    for modname, m in list(sys.modules.items()):
        if m is None or not (modname == k.__name__ or modname.startswith(prefix)):
            continue
        for name, (orig, fn) in originals.items():
            if m.__dict__.get(name) is orig:
                _PATCHED.setdefault((m, name), orig)
                setattr(m, name, fn)


def uninstall() -> None:
    for (m, name), orig in _PATCHED.items():
        setattr(m, name, orig)
    _PATCHED.clear()


__all__ = ["warp_perspective", "warp_affine", "remap", "filter2d", "filter2d_separable", "gaussian_blur2d", "install",
           "uninstall", "augmentation", "config", "core", "filters", "geometry", "graphs", "losses", "metrics", "streaming"]
