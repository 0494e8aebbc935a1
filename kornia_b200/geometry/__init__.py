from . import transform
from .transform import (
    affine,
    center_crop,
    crop_and_resize,
    crop_by_boxes,
    crop_by_transform_mat,
    get_perspective_transform,
    get_rotation_matrix2d,
    remap,
    rotate,
    scale,
    shear,
    translate,
    warp_affine,
    warp_perspective,
)

__all__ = ["transform", "remap", "warp_affine", "warp_perspective", "affine", "rotate", "translate", "scale", "shear",
           "crop_and_resize", "center_crop", "crop_by_boxes", "crop_by_transform_mat", "get_perspective_transform",
           "get_rotation_matrix2d"]
