#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck): batched SHA-512 (aligned, unaligned,
tails), the long-file kernel, the plan kernels, compare (aligned / byte path), synth fill."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O                # noqa: E402  (checker)
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import helpers               # noqa: E402

N.init([0])
rng = np.random.default_rng(5)


def pack(lengths, align):
    offs, pos = [], 0
    for l in lengths:
        pos = (pos + align - 1) // align * align
        offs.append(pos)
        pos += int(l)
    return rng.integers(0, 256, pos + 64, dtype=np.uint8), np.array(offs, dtype=np.uint64), np.array(lengths, dtype=np.uint64)


for align in (16, 1):
    lengths = list(range(0, 300)) + [4096, 65536, 65535, 200_000, 300_001]
    data, off, ln = pack(lengths, align)
    data = data[: int(off[-1] + ln[-1])].copy()        # exact-size buffer: any over-read is out of bounds
    assert np.array_equal(helpers.sha512_batch(data, off, ln), O.sha512_batch(data, off, ln, 4))
    b = data.copy()
    b[int(off[-1]) + 5] ^= 1
    assert np.array_equal(helpers.cmp_batch(data, b, off, ln), O.cmp_batch(data, b, off, ln, 4))
print("sanitize target ok")
