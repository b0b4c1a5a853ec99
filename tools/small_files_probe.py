#!/usr/bin/env python
"""Kernel-only throughput on batches of SMALL files (per-file overhead against the roofline)."""
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
for size, n in ((4096, 400_000), (1024, 1_000_000), (256, 2_000_000), (100, 4_000_000), (0, 4_000_000)):
    lengths = np.full(n, size, dtype=np.uint64)
    off, total = synth.layout(lengths)
    d = torch.empty(max(total, 16), dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    dg = torch.empty((n, 64), dtype=torch.uint8, device="cuda:0")
    for r in (0, 1, 2, 3):
        N.set_option("sha_warps_per_sm", r)
        for _ in range(2):
            device.sha512_batch_device(d, off, lengths, dg)
        torch.cuda.synchronize()
        N.reset_stats()
        for _ in range(5):
            device.sha512_batch_device(d, off, lengths, dg)
        torch.cuda.synchronize()
        s = N.stats()
        ms = s.sha512_kernel_ms_sum / s.sha512_kernel_timed
        blocks = int(synth.blocks(lengths).sum())
        print(json.dumps({"file_bytes": size, "files": n, "ctas_per_sm": r or "auto", "kernel_ms": round(ms, 4),
                          "files_per_s": round(n / ms * 1e3), "frac_of_alu_peak": round(blocks * 3568 / (ms * 1e-3) / PEAK, 3)}), flush=True)
    del d, dg
