#!/usr/bin/env python
"""Compare kernel over CTAs/SM on config 4 (10,000 pairs x 1 MiB, all equal + 1 % differing)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

N.init([0])
dev = torch.device("cuda:0")
n = 10_000
cl = np.full(n, 1 << 20, dtype=np.uint64)
co, tot = synth.layout(cl)
da = torch.empty(tot, dtype=torch.uint8, device=dev)
device.synth_fill_device(da, co, cl)
db = da.clone()
rng = np.random.default_rng(synth.SEED)
differ = rng.choice(n, 100, replace=False)
pos = rng.integers(0, 1 << 20, 100)
db[torch.from_numpy(co[differ].astype(np.int64) + pos).to(dev)] ^= 1
algo = int((n - 100) * 2 * (1 << 20) + sum(2 * 16384 * (int(p) // 16384 + 1) for p in pos))
deq = torch.empty(n, dtype=torch.uint8, device=dev)
for ctas in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4,5,6,8,10,12,15,16,20,24,32").split(",")]:
    N.set_option("cmp_ctas_per_sm", ctas)
    for _ in range(3):
        device.cmp_batch_device(da, db, co, cl, deq)
    torch.cuda.synchronize()
    N.reset_stats()
    for _ in range(10):
        device.cmp_batch_device(da, db, co, cl, deq)
    torch.cuda.synchronize()
    st = N.stats()
    ms = st.cmp_kernel_ms_sum / st.cmp_kernel_timed
    ok = sorted(np.nonzero(deq.cpu().numpy() == 0)[0].tolist()) == sorted(differ.tolist())
    print(f"ctas/SM {ctas:2d}: {ms:.3f} ms  {algo / ms / 1e6:.0f} GB/s  frac {algo / ms / 1e6 / 6553.9:.3f}  ok={ok}", flush=True)
