import ctypes, sys, time, os
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from snappy_b200 import _native as N
from snappy_b200 import device, helpers, synth
N.init([0])
lengths = synth.lognormal_sizes(100_000)
data, off, ln = synth.make_host_batch(lengths)
p = N.lib().snapgpu_alloc_pinned(len(data))
host = np.frombuffer((ctypes.c_uint8 * len(data)).from_address(p), dtype=np.uint8)
host[:] = data
for i in range(6):
    print(f"=== call {i}", file=sys.stderr, flush=True)
    t0=time.perf_counter(); helpers.sha512_batch(host, off, ln); print((time.perf_counter()-t0)*1e3)
