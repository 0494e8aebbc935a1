"""Exception types of the validation contract.

When the real ``kornia`` package is importable its own classes are re-used, so that
``pytest.raises(kornia.core.exceptions.ShapeError)`` keeps working after the drop-in is installed
(:func:`kornia_b200.install`).  Otherwise equivalent classes with the same names, hierarchy and
attributes are defined here (reference: kornia/core/exceptions.py:34-118).
"""
from __future__ import annotations

try:  # pragma: no cover - depends on the environment
    from kornia.core.exceptions import BaseError, ShapeError, TypeCheckError  # type: ignore
except Exception:  # kornia absent (the normal case on the GPU box)

    class BaseError(Exception):
        """Root of the validation errors raised by the checked entry points."""

    class ShapeError(BaseError):
        """A tensor does not have the rank / extents the operator needs."""

        def __init__(self, message, *, actual_shape=None, expected_shape=None):
            super().__init__(message)
            self.actual_shape = actual_shape
            self.expected_shape = expected_shape

    class TypeCheckError(BaseError):
        """An argument is not of the required Python / tensor type."""

        def __init__(self, message, *, actual_type=None, expected_type=None):
            super().__init__(message)
            self.actual_type = actual_type
            self.expected_type = expected_type


__all__ = ["BaseError", "ShapeError", "TypeCheckError"]
