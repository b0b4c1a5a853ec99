#!/usr/bin/env python
"""End-to-end (host buffer) timing of the config 2 batch for a few staging sizes, with the
library's pipeline trace on stderr (SNAPGPU_TRACE=1)."""
import ctypes
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, helpers, synth  # noqa: E402

N.init([0])
lengths = synth.lognormal_sizes(100_000)
off, total = synth.layout(lengths)
d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
device.synth_fill_device(d, off, lengths)
p = N.lib().snapgpu_alloc_pinned(total)
host = np.frombuffer((ctypes.c_uint8 * total).from_address(p), dtype=np.uint8)
torch.from_numpy(host).copy_(d.cpu())
pageable = host.copy()
nbytes = int(lengths.sum())
for name, buf in (("pinned", host), ("pageable", pageable)):
    for mib in (64, 256, 1024, 2048):
        N.set_option("staging_bytes", mib << 20)
        helpers.sha512_batch(buf, off, lengths)
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            helpers.sha512_batch(buf, off, lengths)
            best = min(best, time.perf_counter() - t0)
        print(f"{name} staging {mib} MiB: {best * 1e3:.2f} ms  {nbytes / best / 1e9:.1f} GB/s", flush=True)
