from .check import check, check_is_tensor, check_shape, checks_enabled, disable_checks, enable_checks
from .exceptions import BaseError, ShapeError, TypeCheckError

__all__ = ["check", "check_is_tensor", "check_shape", "checks_enabled", "disable_checks", "enable_checks",
           "BaseError", "ShapeError", "TypeCheckError"]
