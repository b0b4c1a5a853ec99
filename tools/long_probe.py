#!/usr/bin/env python
"""Long-file bin, mode 1 (one lane per file) against mode 2 (a lane pair per file): digests checked
against hashlib, chain time per block from the library's own kernel timing.  argv[1] = MiB per file."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N.init([0])
for nfiles in (1, 4, 16, 17, 64, 500):
    lengths = np.array([(mib << 20) + 128 * i + (i % 7) for i in range(nfiles)], dtype=np.uint64)
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    host = d.cpu().numpy()
    want = [hashlib.sha512(host[int(o):int(o) + int(l)].tobytes()).digest() for o, l in zip(off[:3], lengths[:3])]
    want_last = hashlib.sha512(host[int(off[-1]):int(off[-1]) + int(lengths[-1])].tobytes()).digest()
    for mode in (1, 2):
        N.set_option("long_kernel", mode)
        dg = torch.empty((nfiles, 64), dtype=torch.uint8, device="cuda:0")
        device.sha512_batch_device(d, off, lengths, dg)
        torch.cuda.synchronize()
        N.reset_stats()
        device.sha512_batch_device(d, off, lengths, dg)
        torch.cuda.synchronize()
        s = N.stats()
        got = dg.cpu().numpy()
        ok = all(got[i].tobytes() == w for i, w in enumerate(want)) and got[-1].tobytes() == want_last
        blocks = int(synth.blocks(lengths).max())
        print(json.dumps({"files": nfiles, "mib_each": mib, "mode": mode, "ok": ok, "kernel_ms": s.sha512_kernel_ms_sum,
                          "us_per_block": s.sha512_kernel_ms_sum * 1e3 / blocks,
                          "mb_per_s_per_chain": float(lengths.max()) / (s.sha512_kernel_ms_sum * 1e-3) / 1e6}), flush=True)
    del d
