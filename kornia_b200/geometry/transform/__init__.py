from .imgwarp import remap, warp_affine, warp_perspective

__all__ = ["remap", "warp_affine", "warp_perspective"]
