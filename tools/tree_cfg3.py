#!/usr/bin/env python
"""Config 3's shape through writeHashes on a real tree (tmpfs), scaled so that it fits /dev/shm and a
minute: N small files (config 2's sizes) + 4 long files of M MiB each + an archive of A MiB.  The long
files and the archive are single SHA-512 chains; the streamer advances them side by side beside the
tree's batches, so the call takes about ONE chain (M MiB / ~70 MB/s), not five.  Document checked
against the oracle.   usage: tree_cfg3.py [small_files=20000] [long_mib=64] [archive_mib=16]"""
import json
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
nsmall = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
long_mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
arch_mib = int(sys.argv[3]) if len(sys.argv) > 3 else 16
sys.argv = ["bench"]
import bench                                   # noqa: E402
from oracle import oracle as O                 # noqa: E402
from snappy_b200 import _native as N           # noqa: E402
from snappy_b200 import build, synth           # noqa: E402

N.init([0])
rng = np.random.default_rng(3)
lengths = synth.lognormal_sizes(100_000)[:nsmall]
offs, total = synth.layout(lengths)
data = rng.integers(0, 256, total, dtype=np.uint8)
root = Path("/dev/shm/snapgpu_cfg3_tree")
shutil.rmtree(root, ignore_errors=True)
bench.materialise_tree(root / "t", data, offs, lengths)
(root / "t" / "big").mkdir()
for i in range(4):
    (root / "t" / "big" / f"long{i}.bin").write_bytes(rng.integers(0, 256, (long_mib << 20) + 17 * i, dtype=np.uint8).tobytes())
tar = root / "data.tar.gz"
tar.write_bytes(rng.integers(0, 256, arch_mib << 20, dtype=np.uint8).tobytes())
build.hashes_yaml(str(root / "t"), str(tar))                      # warm-up
t0 = time.perf_counter()
doc = build.hashes_yaml(str(root / "t"), str(tar))
dt = time.perf_counter() - t0
st = N.tree_stats()
paths = bench.tree_paths(root / "t", nsmall) + [str(root / "t" / "big" / f"long{i}.bin") for i in range(4)]
sizes = list(lengths) + [(long_mib << 20) + 17 * i for i in range(4)]
cores = os.cpu_count() or 1
t0 = time.perf_counter()
dg = O.sha512sum_files(paths + [str(tar)], sizes + [arch_mib << 20], cores, True)
cpu_all = time.perf_counter() - t0
hexes = {p: d.tobytes().hex() for p, d in zip(paths + [str(tar)], dg)}
want = O.write_hashes(str(root / "t"), str(tar), hasher=lambda p: hexes[os.fsdecode(p)])
one_chain_s = (long_mib << 20) / 68.6e6
print(json.dumps({"workload": f"{nsmall} small files + 4 x {long_mib} MiB + a {arch_mib} MiB archive, tree on tmpfs",
                  "gpu_ms": dt * 1e3, "phases_ms": {k: st[k] for k in ("pack_ms", "gpu_tail_ms", "chain_tail_ms", "yaml_ms")},
                  "one_chain_of_the_longest_file_ms": one_chain_s * 1e3,
                  "five_chains_one_after_the_other_ms": (4 * (long_mib << 20) + (arch_mib << 20)) / 68.6e6 * 1e3,
                  "cpu_allcores_ms": cpu_all * 1e3, "cores": cores, "yaml_identical_to_oracle": doc == want}))
assert doc == want
shutil.rmtree(root, ignore_errors=True)
