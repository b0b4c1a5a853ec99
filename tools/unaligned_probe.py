#!/usr/bin/env python
"""Config 2 packed densely (no 16-byte alignment): staged any-alignment kernel vs the register-load kernel."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

N.init([0])
lengths = synth.lognormal_sizes(100_000)
lengths = lengths + (np.arange(len(lengths), dtype=np.uint64) % 7)            # odd sizes
off = np.zeros(len(lengths), dtype=np.uint64)
off[1:] = np.cumsum(lengths[:-1])
off += 3                                                                       # and an odd base
total = int(off[-1] + lengths[-1]) + 64
d = torch.randint(0, 256, (total,), dtype=torch.uint8, device="cuda:0")
nblk = int(synth.blocks(lengths).sum())
ref = None
for variant, name in ((0, "staged, any alignment"), (2, "register loads + funnel shifts")):
    N.set_option("sha_variant", variant)
    for _ in range(2):
        dg = device.sha512_batch_device(d, off, lengths)
    torch.cuda.synchronize()
    N.reset_stats()
    for _ in range(5):
        dg = device.sha512_batch_device(d, off, lengths)
    torch.cuda.synchronize()
    st = N.stats()
    ms = st.sha512_kernel_ms_sum / st.sha512_kernel_timed
    h = dg.cpu().numpy().tobytes()
    ref = ref or h
    print(f"unaligned config 2, {name}: {ms:.3f} ms, frac {3568 * nblk / ms / 1e9 / 18.61248:.3f}, same digests {h == ref}", flush=True)
import hashlib
host = d.cpu().numpy()
dgn = np.frombuffer(ref, dtype=np.uint8).reshape(-1, 64)
for i in (0, 1, 77, 99_999):
    assert dgn[i].tobytes() == hashlib.sha512(host[int(off[i]):int(off[i] + lengths[i])].tobytes()).digest()
print("spot checks vs hashlib ok")
