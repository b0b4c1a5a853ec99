#!/usr/bin/env python
"""Kernel-only step loop of bench.py (config 2) with the library trace on: where does a step's time go?"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

N.init([0])
lengths = synth.lognormal_sizes(100_000)
off, total = synth.layout(lengths)
d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
device.synth_fill_device(d, off, lengths)
dg = torch.empty((len(lengths), 64), dtype=torch.uint8, device="cuda:0")
for _ in range(3):
    device.sha512_batch_device(d, off, lengths, dg)
torch.cuda.synchronize()
N.reset_stats()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(steps):
    device.sha512_batch_device(d, off, lengths, dg)
t1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
st = N.stats()
print(f"steps {steps}: device {e0.elapsed_time(e1) / steps:.3f} ms/step, host enqueue {(t1 - t0) * 1e3 / steps:.3f} ms/step, "
      f"kernel {st.sha512_kernel_ms_sum / st.sha512_kernel_timed:.3f} ms")
