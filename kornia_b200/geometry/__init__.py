from . import transform
from .transform import remap, warp_affine, warp_perspective

__all__ = ["transform", "remap", "warp_affine", "warp_perspective"]
