from .distort import distort_points, tilt_projection
from .undistort import undistort_image

__all__ = ["distort_points", "tilt_projection", "undistort_image"]
