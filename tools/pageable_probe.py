#!/usr/bin/env python
"""End-to-end time of the config 2 batch from ORDINARY (pageable) host memory -- what a Go slice handed
straight through cgo is -- over the number of bounce-buffer feeder threads, next to pinned memory."""
import ctypes
import json
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
pageable = d.cpu().numpy().copy()
nbytes = int(lengths.sum())
want = None
for feeders in (1, 2, 4, 6, 8, 0):
    N.set_option("feeders", feeders)
    got = helpers.sha512_batch(pageable, off, lengths)
    want = got if want is None else want
    assert np.array_equal(got, want)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        helpers.sha512_batch(pageable, off, lengths, out=got)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"memory": "pageable", "feeders": feeders or "auto", "ms": best * 1e3, "gb_per_s": nbytes / best / 1e9}), flush=True)
p = N.lib().snapgpu_alloc_pinned(total)
host = np.frombuffer((ctypes.c_uint8 * total).from_address(p), dtype=np.uint8)
host[:] = pageable
got = helpers.sha512_batch(host, off, lengths)
assert np.array_equal(got, want)
best = 1e9
for _ in range(4):
    t0 = time.perf_counter()
    helpers.sha512_batch(host, off, lengths, out=got)
    best = min(best, time.perf_counter() - t0)
print(json.dumps({"memory": "pinned", "ms": best * 1e3, "gb_per_s": nbytes / best / 1e9}), flush=True)
