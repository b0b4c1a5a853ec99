"""Mirror of ``policy.AppArmorDelta`` (/root/reference/policy/policy.go:155-167), the only
production caller of the compare path."""
from __future__ import annotations

import ctypes

from . import _native as N
from .helpers import _names


def AppArmorDelta(oldPath: str, newPath: str, prefix: str) -> tuple[dict[str, bool], dict[str, bool]]:
    pol, npol = ctypes.c_void_p(), ctypes.c_size_t()
    tpl, ntpl = ctypes.c_void_p(), ctypes.c_size_t()
    N.check(N.lib().snapgpu_apparmor_delta(N.fs(oldPath), N.fs(newPath), N.fs(prefix), ctypes.byref(pol),
                                           ctypes.byref(npol), ctypes.byref(tpl), ctypes.byref(ntpl)))
    return ({n: True for n in _names(pol, npol)}, {n: True for n in _names(tpl, ntpl)})
