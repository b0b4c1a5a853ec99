#!/usr/bin/env python
"""Per-warp start/finish times and work of one SHA-512 launch (needs the SNAPGPU_TRACE_WARPS build:
SNAPGPU_LIB=snappy_b200/libsnapgpu_trace.so).  Shows where a launch loses time at its end: per
SM sub-partition, how many blocks it processed and when its last warp finished."""
import ctypes
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

N.init([0])
raw = ctypes.CDLL(str(N.LIB_PATH))
rng = np.random.default_rng(5)
n = 400_000
cases = {
    "lognormal s=1.0 (cfg2 x4)": synth.lognormal_sizes(n),
    "lognormal s=0.5": np.clip(np.round(np.exp(rng.normal(np.log(8192), 0.5, n))), 1024, 65536).astype(np.uint64),
    "half 4K half 64K": np.concatenate([np.full(n // 4, 4096), np.full(n // 8, 65536)]).astype(np.uint64),
}
for name, lengths in cases.items():
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    for r, te in ((2, 0), (3, 0)):         # te: the opt-in "balance" mode of round 1, since removed
        N.set_option("sha_warps_per_sm", r)
        for _ in range(3):
            device.sha512_batch_device(d, off, lengths)
        torch.cuda.synchronize()
        nw = 148 * r * 4
        buf = np.zeros((nw, 4), dtype=np.uint64)
        rc = raw.snapgpu_test_warp_trace(0, buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(nw))
        assert rc == 0
        start, end = buf[:, 0].astype(np.int64), buf[:, 1].astype(np.int64)
        units, blocks = (buf[:, 2] >> np.uint64(32)).astype(np.int64), (buf[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
        smid, hw = (buf[:, 3] >> np.uint64(32)).astype(np.int64), (buf[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
        t0 = start.min()
        fin = (end - t0) / 1e6                              # ms
        span = fin.max()
        key = smid * 4 + (hw & 3)
        smsp_fin = np.array([fin[key == k].max() for k in np.unique(key)])
        smsp_blocks = np.array([blocks[key == k].sum() for k in np.unique(key)])
        rank = hw >> 2                                      # slot rank on the sub-partition: 0 = placed first
        row = {"lengths": name, "ctas_per_sm": r, "balance": te, "makespan_ms": round(float(span), 3),
               "warp_finish_ms p10/p50/p90/max": [round(float(np.percentile(fin, q)), 3) for q in (10, 50, 90, 100)],
               "smsp_finish_ms p10/p50/p90": [round(float(np.percentile(smsp_fin, q)), 3) for q in (10, 50, 90)],
               "smsp_blocks min/mean/max": [int(smsp_blocks.min()), int(smsp_blocks.mean()), int(smsp_blocks.max())],
               "idle_share": round(float(1 - fin.mean() / span), 4),
               "blocks_by_slot_rank": [int(blocks[rank == k].sum()) for k in range(r)],
               "mean_finish_by_slot_rank": [round(float(fin[rank == k].mean()), 3) for k in range(r)]}
        print(json.dumps(row), flush=True)
    del d
