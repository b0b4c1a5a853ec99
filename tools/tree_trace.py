#!/usr/bin/env python
"""writeHashes on the config-2 tree (tmpfs) with the library's trace on: batch timeline of the
tree hasher and the kernels' start/end gaps.  usage: tree_trace.py [reps]"""
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sys.argv = ["bench"]
import bench                                   # noqa: E402
from snappy_b200 import _native as N           # noqa: E402
from snappy_b200 import build, synth           # noqa: E402

N.init([0])
lengths = synth.lognormal_sizes(100_000)
offs, total = synth.layout(lengths)
data = np.random.default_rng(1).integers(0, 256, total, dtype=np.uint8)
root = Path("/dev/shm/snapgpu_trace_tree")
shutil.rmtree(root, ignore_errors=True)
bench.materialise_tree(root / "t", data, offs, lengths)
tar = root / "tar"
tar.write_bytes(b"")
for i in range(reps):
    print(f"==== run {i}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    build.hashes_yaml(str(root / "t"), str(tar))
    print(f"==== run {i}: {1e3 * (time.perf_counter() - t0):.2f} ms (with the Python copy)", file=sys.stderr, flush=True)
    print(N.tree_stats(), file=sys.stderr, flush=True)
N.stats()
shutil.rmtree(root, ignore_errors=True)
