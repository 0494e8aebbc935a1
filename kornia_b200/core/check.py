"""Argument validation with the reference's switch and messages.

Mirrors the observable behaviour of kornia/core/check.py: the three guards used on the filter /
remap path raise ``TypeCheckError`` / ``ShapeError`` / ``BaseError`` with the message fragments the
reference's tests assert on (check.py:180,354; tests/filters/test_filters.py:104-133), and all of
them become no-ops when checks are disabled -- ``KORNIA_CHECKS=0``, ``python -O`` or
:func:`disable_checks` (check.py:63-125).  If the real ``kornia`` is importable its switch is
honoured too, so one call to ``kornia.core.check.disable_checks()`` covers both packages.
"""
from __future__ import annotations

import os
from typing import Any, Optional, Sequence

import torch

from .exceptions import BaseError, ShapeError, TypeCheckError


def _initial_state() -> bool:
    flag = os.getenv("KORNIA_CHECKS")
    if flag is not None:
        return flag.lower() in ("1", "true", "yes", "on")
    return __debug__


_enabled = _initial_state()


def checks_enabled() -> bool:
    if not _enabled:
        return False
    try:  # pragma: no cover - only when kornia is co-installed
        import sys

        mod = sys.modules.get("kornia.core.check")
        if mod is not None:
            return bool(mod.are_checks_enabled())
    except Exception:
        pass
    return True


def disable_checks() -> None:
    global _enabled
    _enabled = False


def enable_checks() -> None:
    global _enabled
    _enabled = True


def check_is_tensor(x: Any, msg: Optional[str] = None) -> None:
    if not checks_enabled() or isinstance(x, torch.Tensor):
        return
    text = f"Type mismatch: expected Tensor, got {type(x)}."
    if msg is not None:
        text += f"\n  {msg}"
    raise TypeCheckError(text, actual_type=type(x), expected_type=torch.Tensor)


def check_shape(x: torch.Tensor, spec: Sequence[str], msg: Optional[str] = None) -> None:
    """``spec`` entries are symbolic names (any extent) or digit strings (exact extent); a leading
    or trailing ``"*"`` absorbs extra dimensions (check.py:166-212)."""
    if not checks_enabled():
        return
    spec = list(spec)
    shape = list(x.shape)
    if spec and spec[0] == "*":
        want, got = spec[1:], shape[len(shape) - (len(spec) - 1):] if len(spec) > 1 else []
    elif spec and spec[-1] == "*":
        want, got = spec[:-1], shape[: len(spec) - 1]
    else:
        want, got = spec, shape
    if len(want) != len(got):
        text = (f"Shape dimension mismatch: expected {len(want)} dimensions, got {len(got)}.\n"
                f"  Expected shape: {spec}\n  Actual shape: {shape}")
        if msg is not None:
            text += f"\n  {msg}"
        raise ShapeError(text, actual_shape=shape, expected_shape=spec)
    for axis, (w, g) in enumerate(zip(want, got)):
        if w.isnumeric() and int(w) != g:
            text = (f"Shape mismatch at dimension {axis}: expected {int(w)}, got {g}.\n"
                    f"  Expected shape: {spec}\n  Actual shape: {shape}")
            if msg is not None:
                text += f"\n  {msg}"
            raise ShapeError(text, actual_shape=shape, expected_shape=spec)


def check(condition: bool, msg: Optional[str] = None) -> None:
    if checks_enabled() and not condition:
        raise BaseError("Validation condition failed" if msg is None else msg)


__all__ = ["check", "check_is_tensor", "check_shape", "checks_enabled", "disable_checks", "enable_checks",
           "BaseError", "ShapeError", "TypeCheckError"]
