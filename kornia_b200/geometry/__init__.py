from . import calibration, transform
from .calibration import distort_points, undistort_image
from .transform import (
    affine,
    build_laplacian_pyramid,
    build_pyramid,
    center_crop,
    crop_and_resize,
    crop_by_boxes,
    crop_by_transform_mat,
    get_perspective_transform,
    get_rotation_matrix2d,
    pyrdown,
    pyrup,
    remap,
    rescale,
    resize,
    rotate,
    scale,
    shear,
    translate,
    warp_affine,
    warp_perspective,
)

__all__ = ["transform", "calibration", "undistort_image", "distort_points", "pyrdown", "pyrup", "build_pyramid",
           "build_laplacian_pyramid", "resize", "rescale", "remap", "warp_affine", "warp_perspective", "affine", "rotate", "translate", "scale", "shear",
           "crop_and_resize", "center_crop", "crop_by_boxes", "crop_by_transform_mat", "get_perspective_transform",
           "get_rotation_matrix2d"]
