#!/usr/bin/env python
"""Randomised parity stress: random batches x random options (kernel variant, launch shape, long-file
bin form, staging size, alignment, pinned/pageable, device-resident/host path) against the oracle.
usage: stress.py [seconds] [seed] [devices]   -- prints one JSON summary line; exits 1 on the first mismatch.
With devices > 1 the host-buffer calls are sharded over that many GPUs in this one process."""
import ctypes
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O                # noqa: E402  (checker only)
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, helpers       # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(seed)
N.init(list(range(ndev)))
O.build()
KiB, MiB = 1 << 10, 1 << 20


def random_lengths():
    n = int(rng.choice([1, 2, 31, 32, 33, 100, 700, 3000]))
    kind = rng.integers(0, 6)
    if kind == 0:
        ln = rng.integers(0, 300, n)
    elif kind == 1:
        ln = rng.choice([0, 1, 111, 112, 127, 128, 129, 239, 240, 255, 256, 4096, 65535, 65536], n)
    elif kind == 2:
        ln = np.clip(np.round(np.exp(rng.normal(np.log(8192), 1.0, n))), 0, 200_000)
    elif kind == 3:
        ln = rng.integers(0, 70_000, n)
    elif kind == 4:                                    # a few long files among short ones
        ln = rng.integers(0, 20_000, n)
        for _ in range(int(rng.integers(1, 6))):
            ln[rng.integers(0, n)] = int(rng.integers(128 * KiB, 6 * MiB))
    else:                                              # many long files (more than the bin takes, sometimes)
        n = int(rng.choice([3, 17, 40, 300]))
        ln = rng.integers(128 * KiB, 400 * KiB, n)
    return np.asarray(ln, dtype=np.uint64)


def layout(ln, align, jitter):
    off = np.zeros(len(ln), dtype=np.uint64)
    pos = int(rng.integers(0, 16)) if jitter else 0
    for i, l in enumerate(ln):
        pos = (pos + align - 1) // align * align
        if jitter:
            pos += int(rng.integers(0, 16))
        off[i] = pos
        pos += int(l)
    return off, pos + 64


cases = mism = 0
t_end = time.time() + budget
while time.time() < t_end:
    ln = random_lengths()
    if int(ln.sum()) > 600 * MiB:
        continue
    align, jitter = (16, False) if rng.random() < 0.6 else (1, True)
    off, total = layout(ln, align, jitter)
    data = rng.integers(0, 256, total, dtype=np.uint8)
    want = O.sha512_batch(data, off, ln, 16, bool(O.lib().oracle_have_openssl()))
    opts = {"sha_variant": int(rng.integers(0, 6)), "sha_warps_per_sm": int(rng.integers(0, 4)),
            "long_kernel": int(rng.integers(0, 3)), "pair_form": int(rng.integers(0, 2)),
            "pair_files_per_cta": int(rng.choice([0, 0, 1, 2, 3, 7, 16])),
            "long_min_blocks": int(rng.choice([0, 0, 16, 64, 1024])), "two_ended": int(rng.integers(0, 3)),
            "staging_bytes": int(rng.choice([1, 3, 16, 64, 1024])) * MiB}
    for k, v in opts.items():
        N.set_option(k, v)
    mode = int(rng.integers(0, 3))
    if mode == 0:                                      # device-resident
        d = torch.from_numpy(data).cuda()
        got = device.sha512_batch_device(d, off, ln).cpu().numpy()
    elif mode == 1:                                    # pageable host buffer
        got = helpers.sha512_batch(data, off, ln)
    else:                                              # pinned host buffer
        p = N.lib().snapgpu_alloc_pinned(total)
        host = np.frombuffer((ctypes.c_uint8 * total).from_address(p), dtype=np.uint8)
        host[:] = data
        got = helpers.sha512_batch(host, off, ln)
        del host
        N.lib().snapgpu_free_pinned(p)
    cases += 1
    if not np.array_equal(got, want):
        bad = np.nonzero((got != want).any(axis=1))[0]
        print(json.dumps({"MISMATCH": True, "case": cases, "opts": opts, "mode": mode, "align": align, "n": len(ln),
                          "bad": bad[:8].tolist(), "bad_len": [int(ln[i]) for i in bad[:8]]}), flush=True)
        mism += 1
        break
    # compare: same layout, a few flipped bytes
    if rng.random() < 0.4 and len(ln):
        b = data.copy()
        flips = sorted(set(int(i) for i in rng.integers(0, len(ln), 3) if ln[int(i)] > 0))
        for i in flips:
            b[int(off[i]) + int(rng.integers(0, int(ln[i])))] ^= 1 << int(rng.integers(0, 8))
        eq = helpers.cmp_batch(data, b, off, ln)
        if np.nonzero(eq == 0)[0].tolist() != flips:
            print(json.dumps({"CMP_MISMATCH": True, "case": cases, "flips": flips,
                              "got": np.nonzero(eq == 0)[0].tolist()[:10]}), flush=True)
            mism += 1
            break
print(json.dumps({"what": "randomised parity stress", "devices": ndev, "seed": seed, "seconds": budget, "cases": cases, "mismatches": mism}))
sys.exit(1 if mism else 0)
