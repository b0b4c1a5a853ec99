#!/usr/bin/env python
"""SHA-512 kernel variants x CTAs/SM on the benchmark shapes, plus the single-file chain
latency per variant.  argv[1] = output .jsonl, argv[2] = comma list of variants."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

out = open(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "variants.jsonl", "w")
variants = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,5,6,7,8,9,10,11").split(",")]
N.init([0])
dev = torch.device("cuda:0")


def emit(d):
    out.write(json.dumps(d) + "\n")
    out.flush()
    print(json.dumps(d), flush=True)


def time_sha(d_data, off, ln, d_dg, steps=5):
    for _ in range(2):
        device.sha512_batch_device(d_data, off, ln, d_dg)
    torch.cuda.synchronize()
    N.reset_stats()
    for _ in range(steps):
        device.sha512_batch_device(d_data, off, ln, d_dg)
    torch.cuda.synchronize()
    st = N.stats()
    return st.sha512_kernel_ms_sum / max(st.sha512_kernel_timed, 1)


workloads = {
    "cfg2": (synth.lognormal_sizes(100_000), (1, 2, 3)),
    "4k-x-400k": (np.full(400_000, 4096, dtype=np.uint64), (1, 2, 3)),
    "cfg5-shard-50k": (np.full(50_000, 65536, dtype=np.uint64), (1, 2, 3)),
    "solo-4MiB": (np.array([4 << 20], dtype=np.uint64), (1,)),
}
for wname, (lengths, shapes) in workloads.items():
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device=dev)
    device.synth_fill_device(d, off, lengths)
    dg = torch.empty((len(lengths), 64), dtype=torch.uint8, device=dev)
    nblk = int(synth.blocks(lengths).sum())
    ref = None
    for variant in variants:
        for warps in shapes:
            N.set_option("sha_variant", variant)
            N.set_option("sha_warps_per_sm", warps)
            k_ms = time_sha(d, off, lengths, dg, steps=3 if wname.startswith("solo") else 5)
            h = dg.cpu().numpy().tobytes()
            ref = ref or h
            emit({"what": "sha", "workload": wname, "variant": variant, "ctas_per_sm": warps, "step_ms": 0.0,
                  "kernel_ms": k_ms, "file_gbs": int(lengths.sum()) / k_ms / 1e6, "blocks": nblk,
                  "us_per_block": k_ms * 1e3 / nblk if wname.startswith("solo") else None,
                  "T_instr_s": 3568 * nblk / k_ms / 1e9, "frac_nominal": 3568 * nblk / k_ms / 1e9 / (148 * 64 * 1.965e-3),
                  "same_digests": h == ref})
    del d, dg
out.close()
