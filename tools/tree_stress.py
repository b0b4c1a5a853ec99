#!/usr/bin/env python
"""Randomised parity stress of the tree entry points on the GPU: random trees (nesting, wide and empty
directories, symlinks, names yaml.v2 quotes, file sizes across every packer class, archives of 0..6 MiB)
-> snapgpu_hashes_yaml vs the oracle's write_hashes; every third case also copyToBuildDir (copy forced)
+ writeHashes from the digest cache, and verification of the written document.  Two caller threads run
cases side by side (sessions, chains and the shared pools under contention).
usage: tree_stress.py [seconds=60] [seed=1]"""
import json
import os
import shutil
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O                 # noqa: E402  (checker)
from snappy_b200 import _native as N           # noqa: E402
from snappy_b200 import build                  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N.init([0])
base = Path("/dev/shm") / f"snapgpu_tree_stress_{os.getpid()}"
shutil.rmtree(base, ignore_errors=True)
base.mkdir()
NAMES = ["true", "123", "1e3", "~", "a b", "x: y", "-", "# c", "Grüße", "null", "0x1f", "1:30", "it's", "[a]", "DEBIANx"]
lock = threading.Lock()
stats = {"cases": 0, "files": 0, "bytes": 0, "copies": 0, "mismatches": 0}
errors = []


def make_tree(root: Path, rng):
    root.mkdir(parents=True)
    dirs = [root]
    for i in range(int(rng.integers(0, 25))):
        d = dirs[int(rng.integers(len(dirs)))] / f"d{i:02d}"
        d.mkdir()
        dirs.append(d)
    shape = int(rng.integers(0, 5))
    n = int(rng.integers(0, 1500))
    nbytes = 0
    for i in range(n):
        if shape == 0:
            size = int(rng.integers(0, 3000))
        elif shape == 1:
            size = int(np.clip(np.rint(np.exp(rng.normal(np.log(8192.0), 1.0))), 0, 70_000))
        elif shape == 2:
            size = int(rng.choice([0, 1, 111, 112, 127, 128, 129, 4096, 65536]))
        elif shape == 3:
            size = int(rng.integers(0, 400_000)) if rng.random() < 0.05 else int(rng.integers(0, 20_000))
        else:
            size = int(rng.integers(200 << 10, 3 << 20)) if rng.random() < 0.01 else int(rng.integers(0, 9000))
        d = dirs[int(rng.integers(len(dirs)))] if rng.random() < 0.7 else dirs[-1]
        p = d / (f"f{i:05d}" if rng.random() < 0.97 else f"{NAMES[int(rng.integers(len(NAMES)))]}{i}")
        p.write_bytes(rng.integers(0, 256, size, dtype=np.uint8).tobytes())
        os.chmod(p, int(rng.choice([0o644, 0o755, 0o600, 0o444, 0o4711])))
        nbytes += size
    if rng.random() < 0.3:
        big = dirs[0] / "long.bin"                       # a chain of its own beside the archive's
        size = (16 << 20) + int(rng.integers(1, 1 << 20))
        big.write_bytes(rng.integers(0, 256, size, dtype=np.uint8).tobytes())
        nbytes += size
        n += 1
    if rng.random() < 0.5:
        os.symlink("nowhere", root / "dangling")
    if rng.random() < 0.3:
        (root / "DEBIAN").mkdir()
        (root / "DEBIAN" / "control").write_bytes(b"skipped")
    return n, nbytes


def worker(wid):
    rng = np.random.default_rng(seed * 1000 + wid)
    t_end = time.time() + budget
    k = 0
    while time.time() < t_end and not errors:
        k += 1
        root = base / f"w{wid}_{k}"
        try:
            n, nbytes = make_tree(root / "t", rng)
            tar = root / "data.tar.gz"
            tar.write_bytes(rng.integers(0, 256, int(rng.choice([0, 1, 4096, 300_000, 2 << 20, (2 << 20) + 1, 6_000_000])), dtype=np.uint8).tobytes())
            want = O.write_hashes(str(root / "t"), str(tar))
            got = build.hashes_yaml(str(root / "t"), str(tar))
            bad = got != want
            copies = 0
            if not bad and k % 3 == 0 and not (root / "t" / "dangling").is_symlink():
                stage = root / "stage"
                build.copyToBuildDir(str(root / "t"), str(stage), no_link=True)
                want2 = O.write_hashes(str(stage), str(tar))
                bad = build.hashes_yaml(str(stage), str(tar)) != want2
                build.writeHashes(str(stage), str(tar))
                report = build.verifyHashes(str(stage), str(stage / "DEBIAN" / "hashes.yaml"), str(tar))
                bad = bad or report != []
                copies = 1
            with lock:
                stats["cases"] += 1
                stats["files"] += n
                stats["bytes"] += nbytes
                stats["copies"] += copies
                if bad:
                    stats["mismatches"] += 1
                    errors.append(f"worker {wid} case {k}: document differs (tree kept at {root})")
            if not bad:
                shutil.rmtree(root, ignore_errors=True)
        except Exception as exc:                        # noqa: BLE001 -- reported in the summary
            with lock:
                errors.append(f"worker {wid} case {k}: {type(exc).__name__}: {exc}")


threads = [threading.Thread(target=worker, args=(w,)) for w in range(2)]
for t in threads:
    t.start()
for t in threads:
    t.join()
print(json.dumps({"what": "randomised tree stress: hashes_yaml / copyToBuildDir / verifyHashes vs the oracle, two caller threads",
                  "seed": seed, "seconds": budget, **stats, "errors": errors[:5]}))
if not errors:
    shutil.rmtree(base, ignore_errors=True)
sys.exit(1 if errors else 0)
