#!/usr/bin/env python
"""In-process multi-device sharding (snapgpu_init over every visible GPU): parity against the
oracle and end-to-end timing of the host-buffer call at 1..N devices.  JSON lines on stdout."""
import ctypes
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O                # noqa: E402  (checker only)
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import helpers, synth        # noqa: E402

ngpu = torch.cuda.device_count()
tree_root = None
if "--tree" in sys.argv:
    # writeHashes on the config-2 tree with every device bound: the tree hasher's batches go to whichever
    # device has a free slot (no exchange between devices); the document must not depend on the device count
    import os
    import shutil
    sys_argv, sys.argv = sys.argv, ["bench"]
    import bench
    sys.argv = sys_argv
    from snappy_b200 import build
    tree_root = Path("/dev/shm/snapgpu_multi_tree")
    shutil.rmtree(tree_root, ignore_errors=True)
    tl = synth.lognormal_sizes(100_000)
    td, toff, tln = synth.make_host_batch(tl)
    bench.materialise_tree(tree_root / "t", td, toff, tln)
    (tree_root / "tar").write_bytes(td[: 3 << 20].tobytes())
    tree_want = None
lengths = synth.lognormal_sizes(100_000)
data, off, ln = synth.make_host_batch(lengths)
want = O.sha512_batch(data, off, ln, 16, bool(O.lib().oracle_have_openssl()))
nbytes = int(ln.sum())
for nd in [n for n in (1, 2, 4, 8) if n <= ngpu]:
    N.init(list(range(nd)))
    p = N.lib().snapgpu_alloc_pinned(len(data))
    host = np.frombuffer((ctypes.c_uint8 * len(data)).from_address(p), dtype=np.uint8)
    host[:] = data
    got = helpers.sha512_batch(host, off, ln)
    assert np.array_equal(got, want), f"{nd} devices: digests differ from the oracle"
    b = host.copy()
    flip = [5, 77_777, 99_999]
    for i in flip:
        b[int(off[i]) + int(ln[i]) - 1] ^= 1
    eq = helpers.cmp_batch(host, b, off, ln)
    assert np.nonzero(eq == 0)[0].tolist() == flip, f"{nd} devices: cmp flags wrong"
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        helpers.sha512_batch(host, off, ln)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"what": "in-process sharding, host buffers (config 2 batch)", "devices": nd, "ms": best * 1e3,
                      "gb_per_s": nbytes / best / 1e9, "bit_exact_with_oracle": True}), flush=True)
    N.lib().snapgpu_free_pinned(p)
    if tree_root is not None:
        build.hashes_yaml(str(tree_root / "t"), str(tree_root / "tar"))
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            doc = build.hashes_yaml(str(tree_root / "t"), str(tree_root / "tar"))
            best = min(best, time.perf_counter() - t0)
        if tree_want is None:
            tree_want = O.write_hashes(str(tree_root / "t"), str(tree_root / "tar"))
        assert doc == tree_want, f"{nd} devices: hashes.yaml differs from the oracle's"
        print(json.dumps({"what": "writeHashes on the config-2 tree (tmpfs, 3 MiB archive), in-process", "devices": nd,
                          "ms": best * 1e3, "phases": N.tree_stats(), "yaml_identical_to_oracle": True}), flush=True)
    if nd > 1 and "--weak" in sys.argv:
        # weak scaling: one config 2 batch PER DEVICE in one pinned buffer, one call
        big_ln = np.tile(ln, nd)
        big_off = np.concatenate([off + np.uint64(k * len(data)) for k in range(nd)])
        p = N.lib().snapgpu_alloc_pinned(len(data) * nd)
        host = np.frombuffer((ctypes.c_uint8 * (len(data) * nd)).from_address(p), dtype=np.uint8)
        for k in range(nd):
            host[k * len(data):(k + 1) * len(data)] = data
        got = helpers.sha512_batch(host, big_off, big_ln)
        assert np.array_equal(got, np.tile(want, (nd, 1))), f"{nd} devices, weak: digests differ from the oracle"
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            helpers.sha512_batch(host, big_off, big_ln, out=got)
            best = min(best, time.perf_counter() - t0)
        print(json.dumps({"what": "in-process sharding, host buffers, one config 2 batch per device (weak)", "devices": nd,
                          "files": int(len(big_ln)), "ms": best * 1e3, "gb_per_s": nbytes * nd / best / 1e9,
                          "bit_exact_with_oracle": True}), flush=True)
        del host
        N.lib().snapgpu_free_pinned(p)
    N.lib().snapgpu_shutdown()
if tree_root is not None:
    import shutil
    shutil.rmtree(tree_root, ignore_errors=True)
