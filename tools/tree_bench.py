#!/usr/bin/env python
"""writeHashes on real file trees (tmpfs): GPU drop-in vs the reference's serial CPU loop.

    tree_bench.py out.jsonl [cfg1] [cfg2]

For each config a tree `d%04d/f%07d.bin` is materialised under /dev/shm with the synthetic
content of SURVEY.md 8(d), plus a real data.tar.gz stand-in.  Timed, best of 3, page cache hot:
  gpu   snappy_b200.build.hashes_yaml  (walk + parallel packer + batched SHA-512 + YAML, C ABI)
  cpu   the oracle's write_hashes      (Python walk/YAML + C Sha512sum per file, 1 thread --
        the shape of the reference's single goroutine)
and the two documents are compared byte for byte."""
import json
import os
import shutil
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O                # noqa: E402  (checker + CPU baseline)
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import build, synth          # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
configs = sys.argv[2:] or ["cfg1", "cfg2"]
N.init([0])
base = Path("/dev/shm/snapgpu_tree")


def make_tree(root: Path, lengths):
    if root.exists():
        shutil.rmtree(root)
    root.mkdir(parents=True)
    data, off, ln = synth.make_host_batch(lengths)
    names = synth.tree_names(len(lengths))
    made = set()
    for i, name in enumerate(names):
        d = root / name.split("/")[0]
        if d not in made:
            d.mkdir()
            made.add(d)
        with open(root / name, "wb") as f:
            f.write(data[int(off[i]): int(off[i]) + int(ln[i])].tobytes())
    tar = root.parent / (root.name + "_data.tar.gz")
    tar.write_bytes(data[: 1 << 20].tobytes())       # stands in for the finished data.tar.gz (one more file to hash)
    return str(tar)


for cfg in configs:
    lengths = np.full(1000, 4096, dtype=np.uint64) if cfg == "cfg1" else synth.lognormal_sizes(100_000)
    root = base / cfg
    tar = make_tree(root, lengths)
    nbytes = int(lengths.sum()) + (1 << 20)
    build.hashes_yaml(str(root), tar)                # warm-up: pinned ring, page cache
    gpu_best, doc = 1e9, None
    for _ in range(3):
        t0 = time.perf_counter()
        doc = build.hashes_yaml(str(root), tar)
        gpu_best = min(gpu_best, time.perf_counter() - t0)
    t0 = time.perf_counter()
    want = O.write_hashes(str(root), tar)
    cpu_s = time.perf_counter() - t0
    # the C part alone (what Go's loop spends in Sha512sum), without the Python walk
    t0 = time.perf_counter()
    for name in synth.tree_names(len(lengths)):
        O.sha512sum(str(root / name))
    cpu_hash_s = time.perf_counter() - t0
    row = {"workload": f"{cfg}: {len(lengths)} files, {nbytes} bytes, tree on tmpfs", "files": len(lengths), "bytes": nbytes,
           "gpu_write_hashes_ms": gpu_best * 1e3, "gpu_files_per_s": len(lengths) / gpu_best, "gpu_gb_per_s": nbytes / gpu_best / 1e9,
           "cpu_1thread_write_hashes_ms": cpu_s * 1e3, "cpu_1thread_sha512sum_loop_ms": cpu_hash_s * 1e3,
           "cpu_files_per_s": len(lengths) / cpu_hash_s, "cpu_gb_per_s": nbytes / cpu_hash_s / 1e9,
           "yaml_bytes": len(doc), "yaml_identical_to_oracle": doc == want,
           "packer_threads": int(os.environ.get("SNAPGPU_PACK_THREADS", "0")) or min(16, os.cpu_count() or 1)}
    assert doc == want, "hashes.yaml differs from the oracle"
    # verification: the document just made, written beside the tree, checked against the tree
    yaml_path = base / (cfg + "_hashes.yaml")
    yaml_path.write_bytes(doc)
    build.verifyHashes(str(root), str(yaml_path), tar)
    verify_best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        report = build.verifyHashes(str(root), str(yaml_path), tar)
        verify_best = min(verify_best, time.perf_counter() - t0)
    assert report == [], report[:3]
    row["gpu_verify_hashes_ms"] = verify_best * 1e3
    yaml_path.unlink()
    # build staging: copyToBuildDir with the copy path forced, then writeHashes
    stage = base / (cfg + "_stage")
    fused_best = 1e9
    for _ in range(2):
        if stage.exists():
            shutil.rmtree(stage)
        N.lib().snapgpu_digest_cache_clear()
        t0 = time.perf_counter()
        build.copyToBuildDir(str(root), str(stage), no_link=True)
        t_copy = time.perf_counter() - t0
        doc2 = build.hashes_yaml(str(stage), tar)
        fused_best = min(fused_best, time.perf_counter() - t0)
    hits = build.digest_cache_stats()[1]
    # the staged tree has no DEBIAN/ yet at copy time, so its document is the source tree's
    assert doc2 == want, "hashes.yaml of the staged tree differs from the oracle"
    shutil.rmtree(stage)
    N.lib().snapgpu_digest_cache_clear()
    t0 = time.perf_counter()
    O.copy_to_build_dir(str(root), str(stage), no_link=True)
    cpu_copy_s = time.perf_counter() - t0
    shutil.rmtree(stage)
    row.update({"gpu_copy_then_write_hashes_ms": fused_best * 1e3, "gpu_copy_part_ms": t_copy * 1e3,
                "digest_cache_hits": int(hits), "cpu_oracle_copy_ms": cpu_copy_s * 1e3,
                "cpu_copy_plus_sha512sum_loop_ms": (cpu_copy_s + cpu_hash_s) * 1e3})
    print(json.dumps(row), flush=True)
    if out:
        out.write(json.dumps(row) + "\n")
        out.flush()
    shutil.rmtree(root)
    os.unlink(tar)
