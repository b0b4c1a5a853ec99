"""Device-resident batch calls (kernel-only path) on top of torch CUDA tensors.

torch is plumbing here: it owns the HBM allocations and the stream; the work is done by
libsnapgpu's kernels through the ``*_device`` entry points of include/snapgpu.h.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as N


def _stream_ptr(stream) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def sha512_batch_device(d_data: torch.Tensor, offsets, lengths, d_digests: torch.Tensor | None = None,
                        dev: int = 0, stream=None) -> torch.Tensor:
    """Enqueue SHA-512 of ``n`` files of the packed CUDA buffer ``d_data``; returns (n, 64) uint8."""
    assert d_data.is_cuda and d_data.dtype == torch.uint8 and d_data.is_contiguous()
    offsets, lengths = _u64(offsets), _u64(lengths)
    n = len(offsets)
    if d_digests is None:
        d_digests = torch.empty((n, 64), dtype=torch.uint8, device=d_data.device)
    N.check(N.lib().snapgpu_sha512_batch_device(dev, d_data.data_ptr(), offsets.ctypes.data, lengths.ctypes.data, n,
                                                 d_digests.data_ptr(), _stream_ptr(stream)))
    return d_digests


def cmp_batch_device(d_a: torch.Tensor, d_b: torch.Tensor, offsets, lengths, d_equal: torch.Tensor | None = None,
                     dev: int = 0, stream=None) -> torch.Tensor:
    assert d_a.is_cuda and d_b.is_cuda and d_a.dtype == torch.uint8 and d_b.dtype == torch.uint8
    offsets, lengths = _u64(offsets), _u64(lengths)
    n = len(offsets)
    if d_equal is None:
        d_equal = torch.empty(n, dtype=torch.uint8, device=d_a.device)
    N.check(N.lib().snapgpu_cmp_batch_device(dev, d_a.data_ptr(), d_b.data_ptr(), offsets.ctypes.data,
                                              lengths.ctypes.data, n, d_equal.data_ptr(), _stream_ptr(stream)))
    return d_equal


def synth_fill_device(d_data: torch.Tensor, offsets, lengths, first_index: int = 0, seed: int = 20150423,
                      dev: int = 0, stream=None) -> None:
    assert d_data.is_cuda and d_data.dtype == torch.uint8
    offsets, lengths = _u64(offsets), _u64(lengths)
    N.check(N.lib().snapgpu_synth_fill_device(dev, d_data.data_ptr(), offsets.ctypes.data, lengths.ctypes.data,
                                              len(offsets), first_index, seed, _stream_ptr(stream)))


def pipe_microbench(kind: int, warps_per_sm: int = 16, dev: int = 0) -> dict:
    import ctypes
    ipc, ms, mhz = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    N.check(N.lib().snapgpu_pipe_microbench(dev, kind, warps_per_sm, ctypes.byref(ipc), ctypes.byref(ms),
                                            ctypes.byref(mhz)))
    return {"kind": kind, "warps_per_sm": warps_per_sm, "warp_inst_per_clk_per_sm": ipc.value,
            "elapsed_ms": ms.value, "sm_clock_mhz": mhz.value}
