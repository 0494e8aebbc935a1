"""kornia_b200 -- B200-native (sm_100a) drop-in for Kornia's warp / filter hot path.

    kornia.geometry.transform.{warp_perspective, warp_affine, remap}
    kornia.filters.{filter2d, filter2d_separable, gaussian_blur2d}

Same function names, signatures, defaults, validation and autograd behaviour; the image work runs
in hand-written CUDA kernels behind the C ABI of ``include/kornia_b200.h``.  CUDA-only: there is
no CPU path and no fallback -- a missing library or a CPU tensor raises.
"""
from __future__ import annotations

from . import core, filters, geometry
from .filters import filter2d, filter2d_separable, gaussian_blur2d
from .geometry.transform import remap, warp_affine, warp_perspective

__version__ = "0.1.0"

_PATCHED = {}


def install(kornia_module=None) -> None:
    """Rebind the six hot-path functions on an imported ``kornia`` package (every re-export site:
    geometry/transform/__init__.py:30, geometry/__init__.py:40, filters/__init__.py:37-38 and the
    defining modules, so internal callers such as ``GaussianBlur2d`` / ``RandomPerspective`` pick
    them up).  ``uninstall()`` restores the originals."""
    import importlib

    k = kornia_module or importlib.import_module("kornia")
    sites = {
        "warp_perspective": (warp_perspective, ["geometry.transform.imgwarp", "geometry.transform", "geometry"]),
        "warp_affine": (warp_affine, ["geometry.transform.imgwarp", "geometry.transform", "geometry"]),
        "remap": (remap, ["geometry.transform.imgwarp", "geometry.transform", "geometry"]),
        "filter2d": (filter2d, ["filters.filter", "filters"]),
        "filter2d_separable": (filter2d_separable, ["filters.filter", "filters"]),
        "gaussian_blur2d": (gaussian_blur2d, ["filters.gaussian", "filters"]),
    }
    for name, (fn, mods) in sites.items():
        for mod in mods:
            m = importlib.import_module(f"{k.__name__}.{mod}")
            if hasattr(m, name):
                _PATCHED.setdefault((m, name), getattr(m, name))
                setattr(m, name, fn)


def uninstall() -> None:
    for (m, name), orig in _PATCHED.items():
        setattr(m, name, orig)
    _PATCHED.clear()


__all__ = ["warp_perspective", "warp_affine", "remap", "filter2d", "filter2d_separable", "gaussian_blur2d", "install",
           "uninstall", "core", "filters", "geometry"]
