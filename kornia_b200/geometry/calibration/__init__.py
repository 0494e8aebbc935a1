from .distort import distort_points, tilt_projection
from .undistort import undistort_image, undistort_image_from_uint8

__all__ = ["distort_points", "tilt_projection", "undistort_image", "undistort_image_from_uint8"]
