#!/usr/bin/env python
"""Small fixed workload for ncu captures: argv = workload variant ctas_per_sm [launches]."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

wname = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 0
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
N.init([0])
N.set_option("sha_variant", variant)
N.set_option("sha_warps_per_sm", ctas)
dev = torch.device("cuda:0")
if wname == "cmp":
    n = 2000
    cl = np.full(n, 1 << 20, dtype=np.uint64)
    co, tot = synth.layout(cl)
    da = torch.empty(tot, dtype=torch.uint8, device=dev)
    device.synth_fill_device(da, co, cl)
    db = da.clone()
    for _ in range(launches):
        eq = device.cmp_batch_device(da, db, co, cl)
    torch.cuda.synchronize()
    print("cmp all equal:", bool(eq.cpu().numpy().all()))
    sys.exit(0)
lengths = {"cfg2": synth.lognormal_sizes(100_000), "cfg5": np.full(50_000, 65536, dtype=np.uint64),
           "4k": np.full(400_000, 4096, dtype=np.uint64),
           "solo": np.full(148 * 4 * 32, 65536, dtype=np.uint64)}[wname]
off, total = synth.layout(lengths)
d = torch.empty(total, dtype=torch.uint8, device=dev)
device.synth_fill_device(d, off, lengths)
for _ in range(launches):
    dg = device.sha512_batch_device(d, off, lengths)
torch.cuda.synchronize()
st = N.stats()
print(f"{wname} v{variant} R{ctas}: kernel {st.sha512_kernel_ms_sum / max(st.sha512_kernel_timed, 1):.3f} ms")
