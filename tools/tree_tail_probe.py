#!/usr/bin/env python
"""A/B of the tree hasher's tail: a worker that is about to go idle writes the chunk it hands over back from
its core's cache first (the default), or does not (SNAPGPU_NO_WRITE_BACK=1); alternating calls on the config 2
tree.  The environment variable is read per call.  JSON lines on stdout."""
import ctypes
import json
import os
import shutil
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys_argv, sys.argv = sys.argv, ["bench"]
import bench                                  # noqa: E402
sys.argv = sys_argv
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import build, synth          # noqa: E402

N.init([0])
lengths = synth.lognormal_sizes(100_000)
data, off, ln = synth.make_host_batch(lengths)
root = Path("/dev/shm/snapgpu_tail_tree")
shutil.rmtree(root, ignore_errors=True)
bench.materialise_tree(root / "t", data, off, ln)
(root / "tar").write_bytes(b"")
libc = ctypes.CDLL(None)
for _ in range(3):
    build.hashes_yaml(str(root / "t"), str(root / "tar"))
res = {0: [], 1: []}
for it in range(int(sys_argv[1]) if len(sys_argv) > 1 else 24):
    for sleep in (1, 0):
        if sleep:
            libc.setenv(b"SNAPGPU_NO_WRITE_BACK", b"1", 1)
        else:
            libc.unsetenv(b"SNAPGPU_NO_WRITE_BACK")
        t0 = time.perf_counter()
        build.hashes_yaml(str(root / "t"), str(root / "tar"))
        ms = (time.perf_counter() - t0) * 1e3
        st = N.tree_stats()
        res[sleep].append((ms, st["pack_ms"], st["gpu_tail_ms"]))
for sleep in (1, 0):
    r = res[sleep]
    print(json.dumps({"write_back_before_idle": not sleep, "calls": len(r),
                      "total_ms": {"best": min(x[0] for x in r), "median": statistics.median(x[0] for x in r)},
                      "pack_ms_median": statistics.median(x[1] for x in r),
                      "gpu_tail_ms_median": statistics.median(x[2] for x in r)}), flush=True)
shutil.rmtree(root, ignore_errors=True)
