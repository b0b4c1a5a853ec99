#!/usr/bin/env python
"""Host-to-device copy rate of a 1 GiB pinned buffer moved in pieces of a given size, on one copy stream
and alternating between two, idle and while N host threads stream through DRAM (what the tree hasher's
packers do): explains the copy rate the batch session reaches with its 2 MiB spans."""
import json
import sys
import threading
import time

import numpy as np
import torch

total = 1 << 30
host = torch.empty(total, dtype=torch.uint8).pin_memory()
host.random_(0, 255)
dev = torch.empty(total, dtype=torch.uint8, device="cuda:0")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(piece, two_streams):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(s1)
    s2.wait_event(e0)
    for k, off in enumerate(range(0, total, piece)):
        st = s2 if (two_streams and k & 1) else s1
        with torch.cuda.stream(st):
            dev[off:off + piece].copy_(host[off:off + piece], non_blocking=True)
    s1.wait_stream(s2)
    e1.record(s1)
    torch.cuda.synchronize()
    return total / (e0.elapsed_time(e1) * 1e-3) / 1e9


stop = False


def hammer(src, dst):
    while not stop:
        np.copyto(dst, src)


for nthreads in (0, 8, 16):
    stop = False
    bufs = [(np.ones(64 << 20, dtype=np.uint8), np.empty(64 << 20, dtype=np.uint8)) for _ in range(nthreads)]
    th = [threading.Thread(target=hammer, args=b) for b in bufs]
    for t in th:
        t.start()
    time.sleep(0.2)
    for piece in (64 << 20, 8 << 20, 4 << 20, 2 << 20, 1 << 20):
        for two in (False, True):
            best = max(run(piece, two) for _ in range(3))
            print(json.dumps({"host_threads_streaming": nthreads, "piece_mib": piece >> 20, "copy_streams": 2 if two else 1, "gb_per_s": round(best, 2)}), flush=True)
    stop = True
    for t in th:
        t.join()
