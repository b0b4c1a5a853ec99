#!/usr/bin/env python
"""What the FIRST writeHashes of a process costs (a `snappy build` makes exactly one): the config 2 tree hashed
once in a fresh process, cold, and after snapgpu_warm(); each in its own process, argv[1] times.  SNAPGPU_PIN=hostalloc
in the environment gives round 2's earlier pinning path (cudaHostAlloc) for comparison.  JSON lines on stdout."""
import json
import os
import shutil
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
root = Path("/dev/shm/snapgpu_cold_tree")

if len(sys.argv) > 1 and sys.argv[1] == "child":          # child MODE [TREE ARCHIVE]
    if len(sys.argv) > 4:
        tree_dir, archive = sys.argv[3], sys.argv[4]
    else:
        tree_dir, archive = str(root / "t"), str(root / "tar")
    from snappy_b200 import _native as N
    from snappy_b200 import build
    t0 = time.perf_counter()
    N.init([0])
    t_init = time.perf_counter() - t0
    t_warm = 0.0
    if sys.argv[2] == "warm":
        t0 = time.perf_counter()
        build.warm()
        t_warm = time.perf_counter() - t0
    t0 = time.perf_counter()
    doc = build.hashes_yaml(tree_dir, archive)
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    build.hashes_yaml(tree_dir, archive)
    t_second = time.perf_counter() - t0
    print(json.dumps({"mode": sys.argv[2], "pin": os.environ.get("SNAPGPU_PIN", "huge pages + cudaHostRegister"),
                      "snapgpu_init_ms": t_init * 1e3, "snapgpu_warm_ms": t_warm * 1e3,
                      "first_write_hashes_ms": t_first * 1e3, "second_write_hashes_ms": t_second * 1e3,
                      "yaml_bytes": len(doc)}), flush=True)
    sys.exit(0)

sys_argv, sys.argv = sys.argv, ["bench"]
import bench                                  # noqa: E402
sys.argv = sys_argv
from snappy_b200 import synth                 # noqa: E402
shutil.rmtree(root, ignore_errors=True)
lengths = synth.lognormal_sizes(100_000)
data, off, ln = synth.make_host_batch(lengths)
bench.materialise_tree(root / "t", data, off, ln)
(root / "tar").write_bytes(b"")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    for mode in ("cold", "warm"):
        for pin in (None, "hostalloc"):
            env = dict(os.environ)
            if pin:
                env["SNAPGPU_PIN"] = pin
            else:
                env.pop("SNAPGPU_PIN", None)
            subprocess.run([sys.executable, __file__, "child", mode], env=env, check=True)
shutil.rmtree(root, ignore_errors=True)
