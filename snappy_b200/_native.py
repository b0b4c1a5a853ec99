"""ctypes binding of libsnapgpu.so (the C ABI declared in include/snapgpu.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C snappy_b200/csrc``.
There is no Python or CPU implementation behind it: if the shared object is missing, or no
CUDA device can be initialised, the calls below raise.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SNAPGPU_LIB") or _HERE / "libsnapgpu.so")    # SNAPGPU_LIB: an experimental build

OK, ECUDA, EINVAL, EIO, EMODE, ENAME, ENOINIT = 0, -1, -2, -3, -4, -5, -6


class SnapGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"snapgpu error {code}: {message}")
        self.code = code
        self.message = message


class Stats(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", ctypes.c_uint64),
        ("sha512_launches", ctypes.c_uint64),
        ("cmp_launches", ctypes.c_uint64),
        ("h2d_bytes", ctypes.c_uint64),
        ("d2h_bytes", ctypes.c_uint64),
        ("last_sha512_kernel_ms", ctypes.c_double),
        ("last_cmp_kernel_ms", ctypes.c_double),
        ("sha512_kernel_ms_sum", ctypes.c_double),
        ("sha512_kernel_timed", ctypes.c_uint64),
        ("cmp_kernel_ms_sum", ctypes.c_double),
        ("cmp_kernel_timed", ctypes.c_uint64),
        ("sha512_long_launches", ctypes.c_uint64),
    ]


class TreeStats(ctypes.Structure):
    _fields_ = [
        ("total_ms", ctypes.c_double),
        ("pack_ms", ctypes.c_double),
        ("gpu_tail_ms", ctypes.c_double),
        ("chain_tail_ms", ctypes.c_double),
        ("yaml_ms", ctypes.c_double),
        ("entries", ctypes.c_uint64),
        ("files_hashed", ctypes.c_uint64),
        ("files_cached", ctypes.c_uint64),
        ("batches", ctypes.c_uint64),
        ("yaml_bytes", ctypes.c_uint64),
        ("pack_threads", ctypes.c_uint),
    ]


_lib = None

_vp, _sz, _u64, _i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int
_cp = ctypes.c_char_p
_pp = ctypes.POINTER(ctypes.c_void_p)
_psz = ctypes.POINTER(ctypes.c_size_t)
_pd = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); every symbol include/snapgpu.h declares
SIGNATURES = {
    "snapgpu_init": (_i, [_vp, _i]),
    "snapgpu_shutdown": (None, []),
    "snapgpu_num_devices": (_i, []),
    "snapgpu_last_error": (_cp, []),
    "snapgpu_version": (_cp, []),
    "snapgpu_set_option": (_i, [_cp, ctypes.c_longlong]),
    "snapgpu_alloc_pinned": (_vp, [_sz]),
    "snapgpu_free_pinned": (None, [_vp]),
    "snapgpu_sha512_batch": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "snapgpu_sha512_stream": (_i, [_vp, _i, _vp, _u64, _u64, _i]),
    "snapgpu_cmp_batch": (_i, [_vp, _vp, _vp, _vp, _sz, _vp]),
    "snapgpu_sha512_batch_device": (_i, [_i, _vp, _vp, _vp, _sz, _vp, _vp]),
    "snapgpu_cmp_batch_device": (_i, [_i, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "snapgpu_sha512sum_file": (_i, [_cp, _cp]),
    "snapgpu_write_hashes": (_i, [_cp, _cp]),
    "snapgpu_hashes_yaml": (_i, [_cp, _cp, _pp, _psz]),
    "snapgpu_write_hashes_digest": (_i, [_cp, _vp]),
    "snapgpu_hashes_yaml_digest": (_i, [_cp, _vp, _pp, _psz]),
    "snapgpu_tree_stats": (_i, [ctypes.POINTER(TreeStats)]),
    "snapgpu_files_are_equal": (_i, [_cp, _cp]),
    "snapgpu_dir_updated": (_i, [_cp, _cp, _cp, _pp, _psz]),
    "snapgpu_apparmor_delta": (_i, [_cp, _cp, _cp, _pp, _psz, _pp, _psz]),
    "snapgpu_free": (None, [_vp]),
    "snapgpu_verify_hashes": (_i, [_cp, _cp, _cp, _pp, _psz]),
    "snapgpu_read_archive_sha512": (_i, [_cp, _cp, _sz]),
    "snapgpu_hasher_new": (_vp, []),
    "snapgpu_hasher_write": (_i, [_vp, _vp, _sz]),
    "snapgpu_hasher_sum": (_i, [_vp, _vp]),
    "snapgpu_hasher_free": (None, [_vp]),
    "snapgpu_copy_to_build_dir": (_i, [_cp, _cp, _i]),
    "snapgpu_warm": (_i, []),
    "snapgpu_should_exclude": (_i, [_cp]),
    "snapgpu_digest_cache_clear": (None, []),
    "snapgpu_digest_cache_stats": (None, [_psz, ctypes.POINTER(ctypes.c_uint64)]),
    "snapgpu_synth_fill_device": (_i, [_i, _vp, _vp, _vp, _sz, _u64, _u64, _vp]),
    "snapgpu_get_stats": (_i, [ctypes.POINTER(Stats)]),
    "snapgpu_reset_stats": (None, []),
    "snapgpu_pipe_microbench": (_i, [_i, _i, _i, _pd, _pd, _pd]),
    "snapgpu_h2d_probe": (_i, [_vp, _sz, _i, _pd]),
    "snapgpu_test_yaml_from_digests": (_i, [_cp, _vp, _sz, _pp, _psz]),
    "snapgpu_test_filehash_yaml": (_i, [_cp, ctypes.c_longlong, _cp, ctypes.c_uint, _pp, _psz]),
    "snapgpu_test_plan_order": (_i, [_vp, _sz, _vp]),
    "snapgpu_test_shard": (_i, [_vp, _sz, _i, _vp]),
    "snapgpu_test_split": (_i, [_vp, _sz, _i, _vp]),
    "snapgpu_test_chunks": (ctypes.c_longlong, [_vp, _vp, _sz, _u64, _i, _vp, _sz]),
    "snapgpu_test_long_bin": (_i, [_vp, _sz, _i, _i, ctypes.c_longlong, _vp, _vp]),
}


def lib() -> ctypes.CDLL:
    """Load libsnapgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SnapGpuError(ENOINIT, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                        f"g.build()'` (or make -C snappy_b200/csrc); there is no CPU fallback")
        L = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().snapgpu_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise SnapGpuError(rc, last_error())


def init(devices=None) -> int:
    """Bind to CUDA devices (list of ordinals, or None for all visible). Returns the count."""
    L = lib()
    if devices is None:
        check(L.snapgpu_init(None, 0))
    else:
        arr = (ctypes.c_int * len(devices))(*devices)
        check(L.snapgpu_init(ctypes.cast(arr, ctypes.c_void_p), len(devices)))
    return L.snapgpu_num_devices()


def ensure_init() -> None:
    if lib().snapgpu_num_devices() == 0:
        init(None)


def set_option(key: str, value: int) -> None:
    check(lib().snapgpu_set_option(key.encode(), int(value)))


def stats() -> Stats:
    s = Stats()
    check(lib().snapgpu_get_stats(ctypes.byref(s)))
    return s


def tree_stats() -> dict:
    """Phases of the calling thread's most recent write_hashes / hashes_yaml."""
    s = TreeStats()
    check(lib().snapgpu_tree_stats(ctypes.byref(s)))
    return {k: getattr(s, k) for k, _ in TreeStats._fields_}


def reset_stats() -> None:
    lib().snapgpu_reset_stats()


def take_string(ptr: ctypes.c_void_p, length: int) -> bytes:
    """Copy a malloc'd buffer handed out by the library and free it."""
    try:
        return ctypes.string_at(ptr, length)
    finally:
        lib().snapgpu_free(ptr)


def fs(path) -> bytes:
    return os.fsencode(path)
