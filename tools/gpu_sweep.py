#!/usr/bin/env python
"""Kernel-only sweep on one B200: pipe-rate probes, then the SHA-512 kernel over
variant x warps-per-SM on configs 2 and 5 (shard), and the compare kernel over CTAs/SM.
Writes JSON lines to the path given as argv[1] (default gpurun_out/sweep.jsonl)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import device, synth         # noqa: E402

out_path = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "sweep.jsonl")
out_path.parent.mkdir(parents=True, exist_ok=True)
out = open(out_path, "w")


def emit(d):
    out.write(json.dumps(d) + "\n")
    out.flush()
    print(json.dumps(d), flush=True)


N.init([0])
dev = torch.device("cuda:0")
props = torch.cuda.get_device_properties(dev)
emit({"what": "device", "name": props.name, "sms": props.multi_processor_count,
      "mem_gb": props.total_memory / 1e9})

names = ["IADD3", "LOP3", "SHF", "IMAD", "IMAD.WIDE", "IADD3+IMAD", "LOP3+IMAD.WIDE", "SHA-mix 10:3:3",
         "IMAD.HI", "LOP3+IMAD.HI", "SHF:IMAD:IMAD.HI 2:1:1", "IADD3+IMAD.X"]
for kind in range(12):
    for w in (4, 8, 16, 32):
        r = device.pipe_microbench(kind, w)
        r.update(what="pipe", name=names[kind])
        emit(r)


def time_sha(d_data, off, ln, d_dg, steps=5):
    for _ in range(2):
        device.sha512_batch_device(d_data, off, ln, d_dg)
    torch.cuda.synchronize()
    N.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        device.sha512_batch_device(d_data, off, ln, d_dg)
    e1.record()
    torch.cuda.synchronize()
    st = N.stats()
    return e0.elapsed_time(e1) / steps, st.sha512_kernel_ms_sum / max(st.sha512_kernel_timed, 1)


workloads = {
    "cfg2": synth.lognormal_sizes(100_000),
    "cfg5-shard-50k": np.full(50_000, 65536, dtype=np.uint64),
    "cfg1": np.full(1000, 4096, dtype=np.uint64),
    "4k-x-400k": np.full(400_000, 4096, dtype=np.uint64),
    "cfg5-shard-250k": np.full(250_000, 65536, dtype=np.uint64),
}
for wname, lengths in workloads.items():
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device=dev)
    device.synth_fill_device(d, off, lengths)
    dg = torch.empty((len(lengths), 64), dtype=torch.uint8, device=dev)
    nblk = int(synth.blocks(lengths).sum())
    ref = None
    for variant in (0, 1, 2):
        for warps in (1, 2, 3):
            N.set_option("sha_variant", variant)
            N.set_option("sha_warps_per_sm", warps)
            step_ms, k_ms = time_sha(d, off, lengths, dg)
            h = dg.cpu().numpy().tobytes()
            ref = ref or h
            emit({"what": "sha", "workload": wname, "variant": variant, "ctas_per_sm": warps, "step_ms": step_ms,
                  "kernel_ms": k_ms, "file_gbs": int(lengths.sum()) / k_ms / 1e6, "blocks": nblk,
                  "T_instr_s": 3568 * nblk / k_ms / 1e9, "frac_nominal": 3568 * nblk / k_ms / 1e9 / (148 * 64 * 1.965e-3),
                  "same_digests": h == ref})
    del d, dg
N.set_option("sha_variant", 0)
N.set_option("sha_warps_per_sm", 0)

# compare kernel
npairs = 4000
cl = np.full(npairs, 1 << 20, dtype=np.uint64)
co, ctotal = synth.layout(cl)
da = torch.empty(ctotal, dtype=torch.uint8, device=dev)
device.synth_fill_device(da, co, cl)
db = da.clone()
deq = torch.empty(npairs, dtype=torch.uint8, device=dev)
for ctas in (2, 4, 8, 16):
    N.set_option("cmp_ctas_per_sm", ctas)
    for _ in range(2):
        device.cmp_batch_device(da, db, co, cl, deq)
    torch.cuda.synchronize()
    N.reset_stats()
    for _ in range(5):
        device.cmp_batch_device(da, db, co, cl, deq)
    torch.cuda.synchronize()
    st = N.stats()
    ms = st.cmp_kernel_ms_sum / max(st.cmp_kernel_timed, 1)
    emit({"what": "cmp", "ctas_per_sm": ctas, "kernel_ms": ms, "gbs": 2 * int(cl.sum()) / ms / 1e6,
          "all_equal": bool(deq.cpu().numpy().all())})
N.set_option("cmp_ctas_per_sm", 0)

# pinned H2D bandwidth (the end-to-end bound)
for mb in (64, 256, 1024):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    dd = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    dd.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dd.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    emit({"what": "h2d", "mb": mb, "gbs": (mb << 20) / dt / 1e9})
out.close()
