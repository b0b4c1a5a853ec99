#!/usr/bin/env python
"""A/B of the long-file bin's lower bound (option long_min_blocks; 1024 = round 1's rule, 0 = the default of
the lane-pair form, 256): the config 2 batch end to end from pinned host memory, small batches kernel-only,
and writeHashes on the config 2 tree.  JSON lines on stdout."""
import ctypes
import json
import shutil
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys_argv, sys.argv = sys.argv, ["bench"]
import bench                                  # noqa: E402
sys.argv = sys_argv
from oracle import oracle as O                # noqa: E402  (checker only)
from snappy_b200 import _native as N          # noqa: E402
from snappy_b200 import build, device, helpers, synth  # noqa: E402

N.init([0])
lengths = synth.lognormal_sizes(100_000)
data, off, ln = synth.make_host_batch(lengths)
p = N.lib().snapgpu_alloc_pinned(len(data))
host = np.frombuffer((ctypes.c_uint8 * len(data)).from_address(p), dtype=np.uint8)
host[:] = data
want = O.sha512_batch(data, off, ln, 16, bool(O.lib().oracle_have_openssl()))
nbytes = int(ln.sum())
root = Path("/dev/shm/snapgpu_longbin_tree")
shutil.rmtree(root, ignore_errors=True)
bench.materialise_tree(root / "t", data, off, ln)
(root / "tar").write_bytes(data[: 3 << 20].tobytes())
tree_want = O.write_hashes(str(root / "t"), str(root / "tar"))
for rounds in range(2):
    for min_blocks in (1024, 0):
        N.set_option("long_min_blocks", min_blocks)
        got = helpers.sha512_batch(host, off, ln)
        assert np.array_equal(got, want)
        ts = []
        for _ in range(8):
            t0 = time.perf_counter()
            helpers.sha512_batch(host, off, ln, out=got)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        print(json.dumps({"what": "config 2 batch, pinned host buffer, one call", "long_min_blocks": min_blocks,
                          "ms_best": ts[0] * 1e3, "ms_median": ts[4] * 1e3, "gb_per_s_best": nbytes / ts[0] / 1e9}), flush=True)
        # small batches, device resident: the first k files of config 2 (kernel time from the library's events)
        for k in (1000, 3000, 10000, 21000, 70000, 100000):
            o2, total = synth.layout(lengths[:k])
            d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
            device.synth_fill_device(d, o2, lengths[:k])
            dg = torch.empty((k, 64), dtype=torch.uint8, device="cuda:0")
            device.sha512_batch_device(d, o2, lengths[:k], dg)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                t0 = time.perf_counter()
                device.sha512_batch_device(d, o2, lengths[:k], dg)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            assert np.array_equal(dg.cpu().numpy(), want[:k])
            print(json.dumps({"what": f"first {k} files of config 2, device resident, one call", "long_min_blocks": min_blocks,
                              "mib": total / 2**20, "ms_best": best * 1e3}), flush=True)
        for taper in (0, 1):
            N.set_option("taper", taper)
            ts = []
            for _ in range(8):
                t0 = time.perf_counter()
                helpers.sha512_batch(host, off, ln, out=got)
                ts.append(time.perf_counter() - t0)
            ts.sort()
            print(json.dumps({"what": f"config 2 batch, pinned host buffer, taper {taper}", "long_min_blocks": min_blocks,
                              "ms_best": ts[0] * 1e3, "ms_median": ts[4] * 1e3, "gb_per_s_best": nbytes / ts[0] / 1e9}), flush=True)
        build.hashes_yaml(str(root / "t"), str(root / "tar"))
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            doc = build.hashes_yaml(str(root / "t"), str(root / "tar"))
            ts.append(time.perf_counter() - t0)
        assert doc == tree_want
        ts.sort()
        print(json.dumps({"what": "writeHashes, config 2 tree (tmpfs)", "long_min_blocks": min_blocks,
                          "ms_best": ts[0] * 1e3, "ms_median": ts[3] * 1e3}), flush=True)
shutil.rmtree(root, ignore_errors=True)
