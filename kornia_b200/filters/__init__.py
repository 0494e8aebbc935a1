from .filter import filter2d, filter2d_separable
from .gaussian import GaussianBlur2d, gaussian_blur2d
from .kernels import gaussian, get_gaussian_kernel1d, get_gaussian_kernel2d, normalize_kernel2d

__all__ = ["filter2d", "filter2d_separable", "gaussian_blur2d", "GaussianBlur2d", "gaussian", "get_gaussian_kernel1d",
           "get_gaussian_kernel2d", "normalize_kernel2d"]
