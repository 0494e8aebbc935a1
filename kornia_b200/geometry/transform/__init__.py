from .affwarp import affine, rotate, scale, shear, translate
from .crop2d import center_crop, crop_and_resize, crop_by_boxes, crop_by_transform_mat
from .imgwarp import remap, warp_affine, warp_perspective
from .matrices import angle_to_rotation_matrix, deg2rad, get_perspective_transform, get_rotation_matrix2d

__all__ = ["remap", "warp_affine", "warp_perspective", "affine", "rotate", "translate", "scale", "shear", "crop_and_resize",
           "center_crop", "crop_by_boxes", "crop_by_transform_mat", "get_perspective_transform", "get_rotation_matrix2d",
           "angle_to_rotation_matrix", "deg2rad"]
