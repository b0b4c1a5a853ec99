#!/usr/bin/env python
"""Chain time per block of the lane-pair kernel's round forms (option pair_form, sha512_pair.cuh) for
one file alone and for 16 / 64 files, digests checked against hashlib.
argv: MiB per file, forms (comma list), file counts (comma list), files per CTA (0 = the default spread)."""
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

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 8
forms = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
N.init([0])
N.set_option("long_kernel", 2)
per_cta = int(sys.argv[4]) if len(sys.argv) > 4 else 0
N.set_option("pair_files_per_cta", per_cta)
counts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 16, 64]
for nfiles in counts:
    lengths = np.array([(mib << 20) + 128 * i + (i % 7) for i in range(nfiles)], dtype=np.uint64)
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    host = d.cpu().numpy()
    want = [hashlib.sha512(host[int(o):int(o) + int(l)].tobytes()).digest() for o, l in zip(off, lengths)]
    for form in forms:
        N.set_option("pair_form", form)
        dg = torch.empty((nfiles, 64), dtype=torch.uint8, device="cuda:0")
        device.sha512_batch_device(d, off, lengths, dg)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            N.reset_stats()
            device.sha512_batch_device(d, off, lengths, dg)
            torch.cuda.synchronize()
            best = min(best, N.stats().sha512_kernel_ms_sum)
        got = dg.cpu().numpy()
        ok = all(got[i].tobytes() == w for i, w in enumerate(want))
        blocks = int(synth.blocks(lengths).max())
        print(json.dumps({"files": nfiles, "mib_each": mib, "pair_form": form, "files_per_cta": per_cta, "ok": ok, "kernel_ms": best,
                          "us_per_block": best * 1e3 / blocks, "clk_per_block_at_1965": best * 1e3 / blocks * 1965,
                          "mb_per_s_per_chain": float(lengths.max()) / (best * 1e-3) / 1e6}), flush=True)
    del d
