#!/usr/bin/env python
"""ncu target: one 2 MiB+ file alone -> the long-file bin (sha512_pair_kernel by default; argv[1] = MiB, default 4)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N.init([0])
lengths = np.array([mib << 20], dtype=np.uint64)
off, total = synth.layout(lengths)
d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
device.synth_fill_device(d, off, lengths)
for _ in range(3):
    dg = device.sha512_batch_device(d, off, lengths)
torch.cuda.synchronize()
st = N.stats()
print(f"{mib} MiB alone: {st.sha512_kernel_ms_sum / st.sha512_kernel_timed:.3f} ms per launch")
