#!/usr/bin/env python
"""Topology + concurrent pinned H2D bandwidth on a multi-GPU box, with and without NUMA binding.
usage: topo_probe.py            (parent: prints topology, then spawns one child per GPU, twice)"""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    idx, bind, start_at = int(sys.argv[2]), sys.argv[3] == "1", float(sys.argv[4])
    from snappy_b200 import numa
    info = numa.bind_to_gpu(idx) if bind else {"bound": False}
    import torch
    torch.cuda.set_device(idx)
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device=f"cuda:{idx}")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < start_at:
        time.sleep(0.001)
    t0 = time.perf_counter()
    for _ in range(8):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"gpu {idx} bind={bind} numa={info.get('numa_node')} bound={info.get('bound')}: {8 * n / dt / 1e9:.1f} GB/s", flush=True)
    sys.exit(0)

import torch  # noqa: E402
ngpu = torch.cuda.device_count()
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
print("allowed cpus:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...")
from snappy_b200 import numa  # noqa: E402
for i in range(ngpu):
    print("gpu", i, "numa node", numa.gpu_numa_node(i))
for bind in ("0", "1"):
    for group in ([0], list(range(ngpu))):
        start = time.time() + 25
        procs = [subprocess.Popen([sys.executable, __file__, "child", str(i), bind, str(start)], stdout=subprocess.PIPE, text=True)
                 for i in group]
        outs = [p.communicate()[0].strip() for p in procs]
        print(f"--- {len(group)} concurrent copies, bind={bind}")
        for o in outs:
            print(o)
