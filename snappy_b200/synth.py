"""Deterministic synthetic inputs of the benchmark configs.

Content: little-endian u64 word ``j`` of file ``i`` is ``splitmix64(seed + i*GOLDEN + j)``,
truncated to the file's length; the same generator exists as a CUDA kernel
(``snapgpu_synth_fill_device``) so that kernel-only runs can fill HBM in place.
Sizes: config 2 is ``clip(round(exp(N(ln 8192, 1))), 1024, 65536)`` from
``numpy.random.default_rng(20150423)``.
"""
from __future__ import annotations

import numpy as np

SEED = 20150423
GOLDEN = 0x9E3779B97F4A7C15
_M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x.astype(np.uint64) + np.uint64(GOLDEN))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def lognormal_sizes(n: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    s = np.rint(np.exp(rng.normal(np.log(8192.0), 1.0, size=n)))
    return np.clip(s, 1024, 65536).astype(np.uint64)


def layout(lengths: np.ndarray, align: int = 16) -> tuple[np.ndarray, int]:
    """Pack files back to back at ``align``-byte boundaries. Returns (offsets, total_bytes)."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    slots = (lengths + np.uint64(align - 1)) // np.uint64(align) * np.uint64(align)
    offsets = np.zeros(len(lengths), dtype=np.uint64)
    if len(lengths) > 1:
        np.cumsum(slots[:-1], out=offsets[1:])
    total = int(offsets[-1] + slots[-1]) if len(lengths) else 0
    return offsets, total + 64          # slack so 16-byte loads never leave the buffer


def blocks(lengths) -> np.ndarray:
    """128-byte SHA-512 blocks per file including padding: (L + 144) // 128."""
    return (np.asarray(lengths, dtype=np.uint64) + np.uint64(144)) // np.uint64(128)


def fill_host(buf: np.ndarray, offsets: np.ndarray, lengths: np.ndarray, first_index: int = 0,
              seed: int = SEED, group_bytes: int = 32 << 20) -> None:
    """Write the synthetic content of files ``first_index + k`` into ``buf`` (uint8)."""
    n = len(offsets)
    if n == 0:
        return
    assert all(int(o) % 8 == 0 for o in offsets[:4]), "layout must be 8-byte aligned"
    words = buf[: len(buf) // 8 * 8].view(np.uint64)
    nwords = (np.asarray(lengths, dtype=np.uint64) + np.uint64(7)) // np.uint64(8)
    woff = np.asarray(offsets, dtype=np.uint64) // np.uint64(8)
    start = 0
    with np.errstate(over="ignore"):
        while start < n:
            end, acc = start, 0
            while end < n and (acc == 0 or acc + int(nwords[end]) * 8 <= group_bytes):
                acc += int(nwords[end]) * 8
                end += 1
            nw = nwords[start:end].astype(np.int64)
            tot = int(nw.sum())
            if tot:
                file_of = np.repeat(np.arange(start, end, dtype=np.uint64), nw)
                first = np.cumsum(nw) - nw
                j = np.arange(tot, dtype=np.uint64) - np.repeat(first.astype(np.uint64), nw)
                base = np.uint64(seed) + (file_of + np.uint64(first_index)) * np.uint64(GOLDEN)
                vals = splitmix64(base + j)
                dst = np.repeat(woff[start:end], nw) + j
                words[dst.astype(np.int64)] = vals
            start = end


def make_host_batch(lengths: np.ndarray, seed: int = SEED, first_index: int = 0):
    """(data uint8, offsets, lengths) of a packed synthetic batch in ordinary host memory."""
    lengths = np.asarray(lengths, dtype=np.uint64)
    offsets, total = layout(lengths)
    data = np.zeros(total, dtype=np.uint8)
    fill_host(data, offsets, lengths, first_index, seed)
    # bytes past each file's end inside its 16-byte slot are zeroed so that host- and
    # device-generated buffers agree wherever a file's bytes are
    return data, offsets, lengths


def file_bytes(i: int, length: int, seed: int = SEED) -> bytes:
    """Content of synthetic file ``i`` (for spot checks at sizes the host cannot hold)."""
    nw = (length + 7) // 8
    with np.errstate(over="ignore"):
        base = np.uint64((seed + i * GOLDEN) & _M64)
        vals = splitmix64(base + np.arange(nw, dtype=np.uint64))
    return vals.tobytes()[:length]


def tree_names(n: int) -> list[str]:
    """Names ``d%04d/f%07d.bin``: 1000 files per directory, Walk order = index order."""
    return ["d%04d/f%07d.bin" % (i // 1000, i) for i in range(n)]
