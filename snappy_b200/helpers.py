"""Mirror of the reference's ``helpers`` package for the hot path (Go -> C ABI -> sm_100a).

Same names, argument meaning and error behaviour as

* ``helpers.Sha512sum``      /root/reference/helpers/helpers.go:188-201
* ``helpers.FilesAreEqual``  /root/reference/helpers/cmp.go:31-59
* ``helpers.DirUpdated``     /root/reference/helpers/cmp.go:97-114

plus the batch forms the Go shim in INTEGRATION.md calls.  All arithmetic runs in
libsnapgpu's CUDA kernels; nothing here computes a hash or compares bytes on the CPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as N


def Sha512sum(infile: str) -> str:
    """Hex SHA-512 of a file. Raises OSError where Go returns the os/io error."""
    out = ctypes.create_string_buffer(129)
    rc = N.lib().snapgpu_sha512sum_file(N.fs(infile), out)
    if rc == N.EIO:
        raise OSError(N.last_error())
    N.check(rc)
    return out.value.decode("ascii")


def FilesAreEqual(a: str, b: str) -> bool:
    """True iff both files can be read and have identical contents; every error is False."""
    return bool(N.lib().snapgpu_files_are_equal(N.fs(a), N.fs(b)))


def _names(ptr, count) -> list[str]:
    total = 0
    raw = ctypes.cast(ptr, ctypes.POINTER(ctypes.c_char))
    out = []
    for _ in range(count.value):
        s = ctypes.string_at(ctypes.addressof(raw.contents) + total)
        out.append(s.decode("utf-8", "surrogateescape"))
        total += len(s) + 1
    N.lib().snapgpu_free(ptr)
    return out


def DirUpdated(dirA: str, dirB: str, pfx: str) -> dict[str, bool]:
    """Files present in both directories whose contents differ, as ``{pfx+name: True}``."""
    ptr, count = ctypes.c_void_p(), ctypes.c_size_t()
    N.check(N.lib().snapgpu_dir_updated(N.fs(dirA), N.fs(dirB), N.fs(pfx), ctypes.byref(ptr), ctypes.byref(count)))
    return {n: True for n in _names(ptr, count)}


# ---- batch forms (what the Go shim calls once per tree / per directory pair) ---------------

def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def sha512_batch(data: np.ndarray, offsets, lengths, out: np.ndarray | None = None) -> np.ndarray:
    """Digests ``(n, 64) uint8`` of ``data[offsets[i] : offsets[i] + lengths[i]]`` (host buffers).
    ``out`` may be a caller-owned ``(n, 64) uint8`` array to be filled (a loop reuses it)."""
    N.ensure_init()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets, lengths = _u64(offsets), _u64(lengths)
    n = len(offsets)
    if out is None:
        out = np.empty((n, 64), dtype=np.uint8)
    assert out.shape == (n, 64) and out.dtype == np.uint8 and out.flags.c_contiguous
    N.check(N.lib().snapgpu_sha512_batch(data.ctypes.data, offsets.ctypes.data, lengths.ctypes.data, n,
                                          out.ctypes.data))
    return out


def cmp_batch(a: np.ndarray, b: np.ndarray, offsets, lengths) -> np.ndarray:
    """``equal[i]`` (uint8 0/1) for pairs at the same offsets of two packed host buffers."""
    N.ensure_init()
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    offsets, lengths = _u64(offsets), _u64(lengths)
    n = len(offsets)
    out = np.zeros(n, dtype=np.uint8)
    N.check(N.lib().snapgpu_cmp_batch(a.ctypes.data, b.ctypes.data, offsets.ctypes.data, lengths.ctypes.data, n,
                                       out.ctypes.data))
    return out


class Sha512Stream:
    """sha512.New(): a hash.Hash-shaped streaming digest of one long message (snapgpu_hasher_*).

    ``Write`` gathers bytes in pinned memory; every full 512 KiB piece is hashed on the GPU by a
    worker thread while the caller keeps writing.  ``Sum`` does not disturb the running state."""

    def __init__(self):
        N.ensure_init()
        self._h = N.lib().snapgpu_hasher_new()
        if not self._h:
            raise N.SnapGpuError(N.ECUDA, N.last_error())

    def Write(self, p) -> int:
        arr = np.frombuffer(p, dtype=np.uint8) if len(p) else np.zeros(1, dtype=np.uint8)
        N.check(N.lib().snapgpu_hasher_write(self._h, arr.ctypes.data, len(p)))
        return len(p)

    def Sum(self) -> bytes:
        out = (ctypes.c_uint8 * 64)()
        N.check(N.lib().snapgpu_hasher_sum(self._h, ctypes.addressof(out)))
        return bytes(out)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            N.lib().snapgpu_hasher_free(h)
