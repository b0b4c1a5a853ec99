#!/usr/bin/env python
"""Which launch shape (CTAs per SM) is best for which length distribution: kernel-only, device-resident."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

N.init([0])
PEAK = 148 * 64 * 1.965e9
rng = np.random.default_rng(5)
n = 400_000
cases = {
    "config 2 (100k lognormal s=1.0)": synth.lognormal_sizes(100_000),
    "config 5 shard (250k x 64K)": np.full(250_000, 65536, dtype=np.uint64),
    "lognormal s=1.0 (cfg2 x4)": synth.lognormal_sizes(n),
    "lognormal s=0.5": np.clip(np.round(np.exp(rng.normal(np.log(8192), 0.5, n))), 1024, 65536).astype(np.uint64),
    "lognormal s=0.25": np.clip(np.round(np.exp(rng.normal(np.log(8192), 0.25, n))), 1024, 65536).astype(np.uint64),
    "uniform 0..64K": rng.integers(0, 65536, n // 2).astype(np.uint64),
    "uniform 8K..16K": rng.integers(8192, 16384, n).astype(np.uint64),
    "half 4K half 64K": np.concatenate([np.full(n // 4, 4096), np.full(n // 8, 65536)]).astype(np.uint64),
    "uniform 4K": np.full(n, 4096, dtype=np.uint64),
    # narrow distributions at the depths where the automatic shape is two CTAs per SM
    "uniform 8K..16K, 100k files": rng.integers(8192, 16384, 100_000).astype(np.uint64),
    "uniform 24K..32K, 87k files": rng.integers(24576, 32768, 87_000).astype(np.uint64),
    "uniform 60K..64K, 58k files": rng.integers(61440, 65536, 58_000).astype(np.uint64),
    "lognormal s=0.25, 150k files": np.clip(np.round(np.exp(rng.normal(np.log(8192), 0.25, 150_000))), 1024, 65536).astype(np.uint64),
}
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, lengths in cases.items():
    if only and only not in name:
        continue
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    blocks = synth.blocks(lengths)
    row = {"lengths": name, "files": len(lengths), "max_over_mean_blocks": round(float(blocks.max() / blocks.mean()), 2),
           "depth": round(float(blocks.sum() / (148 * 128 * blocks.max())), 2)}
    for bal in (1, 2, 0):                  # two-ended claims: by the spread of the lengths (the default), always, never
        N.set_option("two_ended", bal)
        for r in ((0, 1, 2, 3) if bal == 1 else (0, 2, 3)):
            N.set_option("sha_warps_per_sm", r)
            dg = torch.empty((len(lengths), 64), dtype=torch.uint8, device="cuda:0")
            for _ in range(2):
                device.sha512_batch_device(d, off, lengths, dg)
            torch.cuda.synchronize()
            N.reset_stats()
            for _ in range(10):
                device.sha512_batch_device(d, off, lengths, dg)
            torch.cuda.synchronize()
            s = N.stats()
            ms = s.sha512_kernel_ms_sum / s.sha512_kernel_timed
            key = ("auto" if r == 0 else f"R{r}") + {1: "", 2: " two-ended always", 0: " one-ended"}[bal]
            row[key] = round(int(blocks.sum()) * 3568 / (ms * 1e-3) / PEAK, 4)
            print("   ", name, key, row[key], file=sys.stderr, flush=True)
    print(json.dumps(row), flush=True)
    del d
