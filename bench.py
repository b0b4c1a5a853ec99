#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: hashes.yaml SHA-512 GB/s + files/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2]

A "step" is one pass of the hot path over one batch of synthetic input.  At N=1 the batch is
BASELINE config 2 (100 000 files, log-normal 1-64 KiB, ~1.29 GB).  With N>1 (torchrun, one
process per GPU) every rank hashes its own batch of the same shape -- the file list is
sharded, there is no collective on the data path -- so scaling is "weak".

One JSON line on stdout (rank 0):
  value      whole-job GB/s of file bytes with the inputs resident in HBM (plan upload +
             kernel), CUDA events on the launching stream, max over ranks
  e2e        the same metric through the host-buffer C-ABI call (pinned host memory ->
             H2D -> kernel -> D2H digests), every step; h2d_peak_gbs = the raw pinned copy rate
             of the same ranks copying at the same time, frac = value / that ceiling
  e2e_tree   the reference's real entry point: writeHashes(buildDir, dataTar)
             (snappy/build.go:216-270) on trees materialised on tmpfs -- config 2 (100k files)
             and config 1 (1,000 x 4 KiB, with an empty and with a real tar|gzip archive) --
             through snapgpu_hashes_yaml, next to helpers.Sha512sum looped over the same files
             on 1 core (what the reference does) and on all cores; documents byte-identical
  cfg5       BASELINE config 5, strong scaling: 2 M files x 64 KiB split N ways, resident, digests
             copied to the host and gathered on rank 0 BY INDEX (gloo, host memory), spot-checked
  inprocess  (N > 1) what a cgo caller gets: ONE process bound to all N devices, one pinned
             buffer, one snapgpu_sha512_batch call
  roofline   the SHA-512 kernel against the integer-ALU issue peak (it is ALU-bound, not
             HBM-bound; DESIGN.md "Roofline"), plus the HBM figures for context
  cpu_baseline  the oracle (C restatement, OpenSSL block function) on this box's cores
Every rank checks its own digests against the oracle.  Other workloads (cfg1, cfg3) are
parity-test cases; `--workload` runs them for the tables in profiles/ and DESIGN.md.
"""
from __future__ import annotations

import argparse
import ctypes
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ALGO_INSTR_PER_BLOCK = 3568          # int32 instructions per 128-byte block
INT32_LANES_PER_SM = 64              # ALU pipe: 16 lanes x 4 sub-partitions


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "_source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed regions.

    nvidia-smi runs for the whole benchmark (it needs ~100 ms to produce its first row); the
    rows are then filtered by their timestamps to the windows marked with begin()/end()."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None
        self.windows = []
        self._t0 = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def begin(self):
        self._t0 = datetime.datetime.now()

    def end(self):
        if self._t0 is not None:
            self.windows.append((self._t0, datetime.datetime.now()))
            self._t0 = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows, inside = [], []
        for line in Path(self.path).read_text().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) != 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f")
                float(parts[1])
            except ValueError:
                continue
            rows.append(parts)
            pad = datetime.timedelta(milliseconds=25)      # a row describes the ~20 ms before its timestamp
            if any(a - pad <= ts <= b + pad for a, b in self.windows):
                inside.append(parts)
        os.unlink(self.path)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        use = inside or rows
        sm = sorted(float(r[1]) for r in use)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in use)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(use[0][2]),
                "samples": len(use), "samples_in_timed_regions": len(inside), "samples_total": len(rows),
                "power_w_max": max(num(r[3]) for r in use), "reasons": reasons}


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the newest committed ncu --set full capture of this
    bench command (profiles/r*_traffic.json, written by tools/ncu_traffic.py), or None."""
    best = None
    for p in sorted((ROOT / "profiles").glob("r*_traffic.json")):
        try:
            d = json.loads(p.read_text())
        except ValueError:
            continue
        if kernel in d:
            best = dict(d[kernel], source=f"replayed from profiles/{p.name} (an ncu --set full capture of this command; "
                                           f"not measured in this run)")
    return best


def workload(name: str, rank: int):
    """(lengths, first_index, description) of a benchmark config for one rank."""
    from snappy_b200 import synth
    if name == "cfg1":
        n, lengths = 1000, np.full(1000, 4096, dtype=np.uint64)
        desc = "config 1: 1,000 files x 4 KiB"
    elif name == "cfg2":
        n, lengths = 100_000, synth.lognormal_sizes(100_000)
        desc = "config 2: 100k files, log-normal 1-64 KiB (seed 20150423)"
    elif name == "cfg3":
        lengths = np.concatenate([synth.lognormal_sizes(100_000)[:50_000], np.full(4, 1 << 30, dtype=np.uint64)])
        n, desc = len(lengths), "config 3: 50k small files + 4 x 1 GiB"
    elif name == "cfg5":
        per = int(os.environ.get("SNAPGPU_CFG5_FILES", "250000"))      # 2M files / 8 GPUs = 250k x 64 KiB per GPU
        n, lengths = per, np.full(per, 65536, dtype=np.uint64)
        desc = f"config 5 shard: {per} files x 64 KiB per GPU (2M files / 8 GPUs = 250k)"
    else:
        raise SystemExit(f"unknown workload {name}")
    return lengths, rank * n, desc


def config_for(name: str, world: int) -> dict:
    """The `config` object of the JSON line -- the same in the repo arm and the reference arm."""
    from snappy_b200 import synth
    lengths, _, desc = workload(name, 0)
    file_bytes = int(lengths.sum())
    return {"workload": desc, "files_per_gpu": int(len(lengths)), "bytes_per_gpu": file_bytes,
            "blocks_per_gpu": int(synth.blocks(lengths).sum()),
            "sharding": f"file list sharded over {world} GPU(s), no collective",
            "cache": f"inputs ({file_bytes / 1e9:.2f} GB per GPU) larger than the 126 MB L2; no flush needed"}


def tree_paths(root: Path, n: int):
    from snappy_b200 import synth
    return [str(root / name) for name in synth.tree_names(n)]


def materialise_tree(root: Path, data: np.ndarray, offsets, lengths) -> None:
    """d%04d/f%07d.bin under `root` with the bytes of a packed batch."""
    from snappy_b200 import synth
    root.mkdir(parents=True)
    names = synth.tree_names(len(lengths))
    for d in sorted({nm.split("/")[0] for nm in names}):
        (root / d).mkdir()
    mv = memoryview(data)
    base = str(root) + "/"
    for i, name in enumerate(names):
        fd = os.open(base + name, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        os.write(fd, mv[int(offsets[i]): int(offsets[i]) + int(lengths[i])])
        os.close(fd)


def cpu_tree(O, paths, sizes, cores: int, use_ossl: bool) -> dict:
    """helpers.Sha512sum looped over the tree's files (snappy/build.go:240): 1 core = what the
    reference does, all cores = the file list statically sharded.  The walk and the
    YAML marshalling are NOT included (that favours the CPU figure)."""
    nbytes = int(np.sum(sizes))
    t0 = time.perf_counter()
    d1 = O.sha512sum_files(paths, sizes, 1, use_ossl)
    one = time.perf_counter() - t0
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        dn = O.sha512sum_files(paths, sizes, cores, use_ossl)
        best = min(best, time.perf_counter() - t0)
    assert np.array_equal(d1, dn)
    return {"digests": d1, "cpu_1core_ms": one * 1e3, "cpu_1core_gbs": nbytes / one / 1e9, "cpu_1core_files_per_s": len(paths) / one,
            "cpu_allcores_ms": best * 1e3, "cpu_allcores_gbs": nbytes / best / 1e9, "cpu_allcores_files_per_s": len(paths) / best,
            "cores": cores}


def bench_root() -> Path:
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    return base / f"snapgpu_bench_{os.getpid()}"


def pinned_array(nbytes: int):
    from snappy_b200 import _native as N
    p = N.lib().snapgpu_alloc_pinned(nbytes)
    if not p:
        raise RuntimeError("snapgpu_alloc_pinned failed: " + N.last_error())
    buf = (ctypes.c_uint8 * nbytes).from_address(p)
    return np.frombuffer(buf, dtype=np.uint8), p


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU path (oracle port; the reference itself is Go and
    cannot be built in this image) on the host cores, same config/metric/unit.  Under torchrun only
    rank 0 works; it starts once the other ranks' interpreters have gone (they compete for the
    cores while they start up, which made the N=8 figure of round 1 read a third low)."""
    flag_dir = Path(tempfile.gettempdir()) / f"snapgpu_ref_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
    if rank != 0:
        try:
            flag_dir.mkdir(exist_ok=True)
            (flag_dir / f"rank{rank}").write_text("gone")
        except OSError:
            pass
        return
    from oracle import oracle as O
    from snappy_b200 import synth
    O.build()
    lengths, first, desc = workload(args.workload, 0)
    if args.workload == "cfg3":
        lengths = lengths[:50_000]          # bounded sample: the small-file phase
    data, offsets, lengths = synth.make_host_batch(lengths, first_index=first)
    waited = 0.0
    while world > 1 and waited < 20.0 and len(list(flag_dir.glob("rank*"))) < world - 1:
        time.sleep(0.1)
        waited += 0.1
    if world > 1:
        time.sleep(0.5)                     # let them exit
        import shutil
        shutil.rmtree(flag_dir, ignore_errors=True)
    cores = len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    use_ossl = bool(O.lib().oracle_have_openssl())
    total = int(lengths.sum())
    for _ in range(max(args.warmup, 1)):
        O.sha512_batch(data, offsets, lengths, cores, use_ossl)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dg = O.sha512_batch(data, offsets, lengths, cores, use_ossl)
    dt = (time.perf_counter() - t0) / args.steps
    t1 = time.perf_counter()
    O.sha512_batch(data, offsets, lengths, 1, use_ossl)
    dt1 = time.perf_counter() - t1
    gbs = total / dt / 1e9
    kind = "port"
    sample = f"{desc}: the whole batch ({len(lengths)} files, {total} bytes) per step, in memory, {cores} threads"
    tree = None
    if args.workload == "cfg2" and not args.no_tree:
        # the same loop on a real tree (tmpfs): open / 32 KiB reads / close per file, as Sha512sum does
        import shutil
        root = bench_root()
        try:
            materialise_tree(root / "cfg2", data, offsets, lengths)
            res = cpu_tree(O, tree_paths(root / "cfg2", len(lengths)), lengths, cores, use_ossl)
            assert np.array_equal(res.pop("digests"), dg)
            tree = dict(res, workload="config 2 tree on tmpfs: helpers.Sha512sum looped over the files, walk and YAML not included")
        finally:
            shutil.rmtree(root, ignore_errors=True)
    line = {
        "impl": "reference", "metric": "hashes.yaml SHA-512 throughput", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "files_per_s": len(lengths) / dt,
        "config": config_for(args.workload, args.gpus),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample,
                         "value_1core": total / dt1 / 1e9,
                         "note": ("C restatement of helpers.Sha512sum's loop with OpenSSL's SHA-512 block function "
                                  "(not Go: no Go toolchain in this image); a fixed CPU job whatever --gpus says, so "
                                  "against N GPUs of weak-scaled work the ratio grows with N by construction" if use_ossl else
                                  "C restatement of helpers.Sha512sum with its own scalar block function"),
                         "tree": tree,
                         "digest_check": dg[0].tobytes().hex()[:16]},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg5"])
    ap.add_argument("--no-cmp", action="store_true", help="skip the config 4 compare kernel figures")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tail", action="store_true", help="skip the single-long-file latency sample")
    ap.add_argument("--no-tree", action="store_true", help="skip the writeHashes-on-a-tmpfs-tree leg")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the config 5 strong-scaling leg")
    ap.add_argument("--no-inprocess", action="store_true", help="skip the one-process-all-devices leg (N > 1)")
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--warps", type=int, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from snappy_b200 import _native as N
    from snappy_b200 import device, helpers, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsnapgpu has no CPU fallback")
    from snappy_b200 import numa
    placement = numa.bind_to_gpu(local_rank)      # before any pinned allocation (first touch)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # a second, host-side group: digests are gathered in host memory (north_star: "digests
        # gathered on the host"), and ranks that are done wait on sockets, not in a spinning kernel
        host_group = dist.new_group(backend="gloo")
    N.init([local_rank])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if args.variant is not None:
        N.set_option("sha_variant", args.variant)
    if args.warps is not None:
        N.set_option("sha_warps_per_sm", args.warps)
    peaks = measured_peaks()

    def barrier():
        if world > 1:
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs, generated in place in HBM ---------------------------------------------------
    lengths, first_index, desc = workload(args.workload, rank)
    offsets, total_alloc = synth.layout(lengths)
    file_bytes = int(lengths.sum())
    nblocks = int(synth.blocks(lengths).sum())
    d_data = torch.empty(total_alloc, dtype=torch.uint8, device=dev)
    device.synth_fill_device(d_data, offsets, lengths, first_index=first_index)
    d_digests = torch.empty((len(lengths), 64), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    # ---- kernel-only: data resident ----------------------------------------------------------
    for _ in range(args.warmup):
        device.sha512_batch_device(d_data, offsets, lengths, d_digests)
    torch.cuda.synchronize()
    N.reset_stats()
    barrier()
    sampler.begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device.sha512_batch_device(d_data, offsets, lengths, d_digests)
    e1.record()
    torch.cuda.synchronize()
    sampler.end()
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    st = N.stats()
    launches = int(st.kernel_launches)
    kernel_ms = st.sha512_kernel_ms_sum / max(st.sha512_kernel_timed, 1)
    kernel_ms = max_over_ranks(kernel_ms)
    digest_host = d_digests.cpu().numpy()

    # ---- end to end: pinned host memory -> C ABI -> digests on the host ----------------------
    e2e = None
    host_view = None
    if not args.no_e2e:
        host_view, host_ptr = pinned_array(total_alloc)
        host_t = torch.from_numpy(host_view)
        host_t.copy_(d_data.cpu())                      # same synthetic bytes as the resident copy
        for _ in range(2):
            dg = helpers.sha512_batch(host_view, offsets, lengths)
        assert np.array_equal(dg, digest_host), "end-to-end digests differ from the device-resident run"
        e2e_steps = max(3, min(args.steps, 10))
        N.reset_stats()
        barrier()
        sampler.begin()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            dg = helpers.sha512_batch(host_view, offsets, lengths, out=dg)     # digests land in host memory
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        sampler.end()
        st2 = N.stats()
        # the ceiling: the same ranks copying the same pinned buffers at the same time, nothing else
        probe_bytes = min(total_alloc, 1 << 30) & ~4095
        secs = ctypes.c_double()
        barrier()
        N.check(N.lib().snapgpu_h2d_probe(host_ptr, probe_bytes, 4, ctypes.byref(secs)))
        probe_s = max_over_ranks(secs.value)
        h2d_peak = world * probe_bytes * 4 / probe_s / 1e9
        barrier()
        e2e_gbs = world * file_bytes / dt / 1e9
        e2e = {"value": e2e_gbs, "unit": "GB/s", "ms_per_step": dt * 1e3,
               "files_per_s": world * len(lengths) / dt,
               "h2d_bytes_per_step": int(st2.h2d_bytes // e2e_steps), "d2h_bytes_per_step": int(st2.d2h_bytes // e2e_steps),
               "bound": "PCIe host-to-device copy (pinned host memory; chunks of 64 MiB, 256 MiB, then up to 1 GiB, ending on 128 MiB and 64 MiB, double buffered)",
               "timing": "host wall clock around the synchronous C-ABI call (digests are in host memory on return), max over ranks",
               "h2d_gbs_per_gpu": st2.h2d_bytes / e2e_steps / dt / 1e9,
               "h2d_peak_gbs": h2d_peak, "h2d_peak_gbs_per_gpu": h2d_peak / world, "frac": e2e_gbs / h2d_peak,
               "h2d_peak_how": f"snapgpu_h2d_probe: {world} rank(s) at once, each 4 x {probe_bytes >> 20} MiB cudaMemcpyAsync from its "
                               f"pinned buffer on the pipeline's copy stream, CUDA events, slowest rank"}

    # ---- every rank checks ITS digests against the oracle (not only rank 0) --------------------
    parity_checked = 0
    if host_view is not None and not args.no_cpu:
        from oracle import oracle as O
        O.build()
        ncores = len(os.sched_getaffinity(0)) or os.cpu_count() or 1
        threads = max(1, ncores // world)
        sl = lengths if args.workload != "cfg3" else lengths[:50_000]
        ref = O.sha512_batch(host_view, offsets[: len(sl)], sl, threads, bool(O.lib().oracle_have_openssl()))
        assert np.array_equal(ref, digest_host[: len(sl)]), f"rank {rank}: GPU digests differ from the oracle"
        parity_checked = len(sl)
    if world > 1:
        t = torch.tensor([float(parity_checked)], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        parity_total = int(t.item())
    else:
        parity_total = parity_checked

    # ---- config 5, strong scaling: 2 M x 64 KiB split over the ranks, digests gathered by index --
    cfg5 = None
    cfg5_fits = False
    if not args.no_cfg5 and args.workload == "cfg2":
        del d_data
        torch.cuda.empty_cache()
        total_files = int(os.environ.get("SNAPGPU_CFG5_TOTAL", "2000000"))
        from snappy_b200 import sharding
        shards5 = sharding.contiguous_shards(np.full(total_files - total_files % world, 65536, dtype=np.uint64), world)
        per = shards5[rank][1] - shards5[rank][0]                # uniform sizes: equal contiguous index ranges
        need5 = per * (65536 + 64 + 36 + 8) + (4 << 30)
        fits = torch.tensor([1.0 if torch.cuda.mem_get_info(dev)[0] > need5 else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(fits, op=dist.ReduceOp.MIN)          # every rank takes the same decision
        cfg5_fits = bool(fits.item() > 0)
        if not cfg5_fits and rank == 0:
            cfg5 = {"skipped": f"a rank has less than {need5 / 1e9:.0f} GB of free HBM for its {per} files x 64 KiB"}
    if cfg5_fits:
        lo = shards5[rank][0]
        l5 = np.full(per, 65536, dtype=np.uint64)
        o5, t5 = synth.layout(l5)
        d5 = torch.empty(t5, dtype=torch.uint8, device=dev)
        device.synth_fill_device(d5, o5, l5, first_index=lo)
        g5 = torch.empty((per, 64), dtype=torch.uint8, device=dev)
        h5 = torch.empty((per, 64), dtype=torch.uint8).pin_memory()
        device.sha512_batch_device(d5, o5, l5, g5)              # warm-up (plan slots of this size)
        torch.cuda.synchronize()
        N.reset_stats()
        steps5 = 3
        barrier()
        sampler.begin()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(steps5):
            device.sha512_batch_device(d5, o5, l5, g5)
            h5.copy_(g5, non_blocking=True)                    # the shard's digests land in host memory
        c1.record()
        torch.cuda.synchronize()
        sampler.end()
        ms5 = max_over_ranks(c0.elapsed_time(c1) / steps5)
        st5 = N.stats()
        k5 = max_over_ranks(st5.sha512_kernel_ms_sum / max(st5.sha512_kernel_timed, 1))
        launches += int(st5.kernel_launches)
        # gather on the host, by index: shards are contiguous by count, so rank order is index order
        t0 = time.perf_counter()
        all5 = sharding.gather_digests(h5.numpy(), shards5, rank, world, group=host_group)    # the code tests/test_dist_cpu.py runs under gloo
        gather_ms = (time.perf_counter() - t0) * 1e3
        if rank == 0:
            import hashlib
            rng5 = np.random.default_rng(5)
            picks = sorted(set([0, per - 1, per % total_files, world * per - 1] + rng5.integers(0, world * per, 28).tolist()))
            for i in picks:
                assert all5[i].tobytes() == hashlib.sha512(synth.file_bytes(int(i), 65536)).digest(), f"config 5: digest {i} wrong"
            b5 = world * per * 65536
            cfg5 = {"workload": f"config 5: {world * per} files x 64 KiB = {b5 / 1e9:.1f} GB split over {world} GPU(s) "
                                f"({per} files each, contiguous by count), resident in HBM",
                    "scaling": "strong", "files_total": world * per, "files_per_gpu": per, "ms_per_step": ms5, "kernel_ms": k5,
                    "gbs": b5 / (ms5 * 1e-3) / 1e9, "files_per_s": world * per / (ms5 * 1e-3),
                    "roofline_frac": ALGO_INSTR_PER_BLOCK * per * 513 / (k5 * 1e-3) / 1e12 /
                                     (torch.cuda.get_device_properties(dev).multi_processor_count * INT32_LANES_PER_SM * peaks["sm_max_mhz"] * 1e-6),
                    "timed": "kernel + copy of the shard's digests to pinned host memory, CUDA events, max over ranks",
                    "digest_gather_ms": gather_ms, "digest_gather": "snappy_b200.sharding.gather_digests over a gloo (host) group: rank 0 places every rank's digests at its shard's index range" if world > 1 else "single rank",
                    "spot_checked_vs_hashlib": len(picks),
                    "speedup_note": "strong scaling: compare ms_per_step with the --gpus 1 line of the same build"}
        del d5, g5, h5
        torch.cuda.empty_cache()

    if rank != 0:
        # idle on a socket (not in an NCCL kernel) while rank 0 runs its single-process legs --
        # among them one process driving every GPU of the box, this one's included
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        dist.destroy_process_group()
        return
    clocks = sampler.stop()

    # ---- writeHashes on real trees (tmpfs): the reference's entry point -------------------------
    e2e_tree = None
    if not args.no_tree and host_view is not None and args.workload == "cfg2":
        import shutil
        from oracle import oracle as O
        from snappy_b200 import build as B
        O.build()
        ncores = len(os.sched_getaffinity(0)) or os.cpu_count() or 1
        use_ossl = bool(O.lib().oracle_have_openssl())
        root = bench_root()
        tree_error = None
        e2e_tree = {"where": str(root.parent), "call": "snapgpu_hashes_yaml(buildDir, dataTar): walk + read + H2D + SHA-512 + "
                    "D2H + hashes.yaml in memory, best of 3 after one warm-up call, page cache hot",
                    "cpu": "helpers.Sha512sum (open / 32 KiB reads / close, OpenSSL block function) looped over the same files; "
                           "walk and YAML marshalling not included in the CPU figures"}
        try:
            def one_tree(tag, tree_dir, tar, paths, sizes, nbytes_tree):
                B.hashes_yaml(str(tree_dir), str(tar))                      # warm-up
                best, doc, phases = 1e30, None, None
                N.reset_stats()
                for _ in range(3):
                    # the C-ABI call itself is timed; copying the malloc'd document into a Python
                    # bytes object afterwards is the harness's business
                    ptr, ln = ctypes.c_void_p(), ctypes.c_size_t()
                    t0 = time.perf_counter()
                    rc = N.lib().snapgpu_hashes_yaml(N.fs(str(tree_dir)), N.fs(str(tar)), ctypes.byref(ptr), ctypes.byref(ln))
                    dt_ = time.perf_counter() - t0
                    N.check(rc)
                    doc = N.take_string(ptr, ln.value)
                    if dt_ < best:
                        best, phases = dt_, N.tree_stats()
                stt = N.stats()
                cpu = cpu_tree(O, paths, sizes, ncores, use_ossl)
                hexes = {p: d.tobytes().hex() for p, d in zip(paths, cpu.pop("digests"))}
                t0 = time.perf_counter()
                tar_hex = O.sha512sum(str(tar))
                tar_ms = (time.perf_counter() - t0) * 1e3
                hexes[str(tar)] = tar_hex
                want = O.write_hashes(str(tree_dir), str(tar), hasher=lambda p: hexes[os.fsdecode(p)])
                assert doc == want, f"{tag}: hashes.yaml differs from the oracle's"
                tar_bytes = os.path.getsize(tar)
                nb = nbytes_tree + tar_bytes
                res = {"files": len(paths), "bytes": nb, "archive_bytes": tar_bytes,
                       "gpu_ms": best * 1e3, "gpu_gbs": nb / best / 1e9, "gpu_files_per_s": len(paths) / best,
                       "phases_ms": {k: phases[k] for k in ("total_ms", "pack_ms", "gpu_tail_ms", "chain_tail_ms", "yaml_ms")},
                       "pack_threads": phases["pack_threads"], "batches": phases["batches"], "yaml_bytes": len(doc),
                       "yaml_identical_to_oracle": True, "gpu_launches": int(stt.kernel_launches),
                       "h2d_bytes": int(stt.h2d_bytes // 3), "cpu_archive_1core_ms": tar_ms}
                res.update(cpu)
                res["cpu_1core_ms"] += tar_ms
                res["cpu_allcores_ms"] += tar_ms                                # the archive is one chain on one core
                res["speedup_vs_1core"] = res["cpu_1core_ms"] / res["gpu_ms"]
                res["speedup_vs_allcores"] = res["cpu_allcores_ms"] / res["gpu_ms"]
                return res

            # config 2: 100k files, the bytes of this rank's batch; the archive stand-in is empty, so
            # the figure is the tree's (a real archive is one more, serial, file: see config 1)
            t2 = root / "cfg2"
            materialise_tree(t2, host_view, offsets, lengths)
            tar2 = root / "cfg2_data.tar.gz"
            tar2.write_bytes(b"")
            e2e_tree["cfg2"] = one_tree("cfg2", t2, tar2, tree_paths(t2, len(lengths)), lengths, file_bytes)
            # What a `snappy build` pays is the FIRST call of a process: the same tree in a fresh process, cold and
            # after snapgpu_warm() (tools/cold_start_probe.py; the child binds device 0 of this rank's view).
            try:
                first = {}
                for mode in ("cold", "warm"):
                    out = subprocess.run([sys.executable, str(ROOT / "tools" / "cold_start_probe.py"), "child", mode, str(t2),
                                          str(tar2)], capture_output=True, text=True, timeout=300, check=True).stdout
                    r = json.loads(out.strip().splitlines()[-1])
                    first[mode] = {k: round(r[k], 2) for k in ("snapgpu_warm_ms", "first_write_hashes_ms", "second_write_hashes_ms")}
                first["what"] = ("config 2 tree, one fresh process per row: its first and second snapgpu_hashes_yaml; 'warm' "
                                 "calls snapgpu_warm() first (what a build does from a goroutine while it copies and compresses)")
                e2e_tree["cfg2_first_call_of_a_process"] = first
            except Exception as exc:                                 # a measurement beside the line, never its failure
                e2e_tree["cfg2_first_call_of_a_process"] = {"error": repr(exc)[:200]}
            shutil.rmtree(t2)
            # config 1: 1,000 x 4 KiB, with an empty archive and with a real tar | gzip of the tree
            l1 = np.full(1000, 4096, dtype=np.uint64)
            d1, o1, _ = synth.make_host_batch(l1)
            t1_ = root / "cfg1"
            materialise_tree(t1_, d1, o1, l1)
            tar1 = root / "cfg1_empty.tar.gz"
            tar1.write_bytes(b"")
            e2e_tree["cfg1_empty_archive"] = one_tree("cfg1", t1_, tar1, tree_paths(t1_, 1000), l1, int(l1.sum()))
            tar1r = root / "cfg1_data.tar.gz"
            subprocess.check_call(["tar", "-C", str(t1_), "-czf", str(tar1r), "."])
            shutil.rmtree(t1_ / "DEBIAN", ignore_errors=True)
            r1 = one_tree("cfg1", t1_, tar1r, tree_paths(t1_, 1000), l1, int(l1.sum()))
            # the flow of INTEGRATION.md 3b: the archive goes through a snapgpu_hasher WHILE tar | gzip -9
            # writes it, writeHashes then gets the digest (nothing is read again)
            from snappy_b200 import helpers as H
            shutil.rmtree(t1_ / "DEBIAN", ignore_errors=True)
            tz = time.perf_counter()
            subprocess.check_call(["tar", "-C", str(t1_), "-czf", str(root / "plain.tar.gz"), "."])
            gzip_only_ms = (time.perf_counter() - tz) * 1e3
            hs = H.Sha512Stream()
            tz = time.perf_counter()
            pz = subprocess.Popen(["tar", "-C", str(t1_), "-cz", "."], stdout=subprocess.PIPE)
            with open(root / "streamed.tar.gz", "wb") as fz:
                while True:
                    piece = pz.stdout.read(1 << 18)
                    if not piece:
                        break
                    fz.write(piece)
                    hs.Write(piece)
            pz.wait()
            written_ms = (time.perf_counter() - tz) * 1e3
            dz = hs.Sum()
            sum_ms = (time.perf_counter() - tz) * 1e3 - written_ms
            assert dz.hex() == O.sha512sum(str(root / "streamed.tar.gz"))
            best_d = 1e30
            for _ in range(3):
                tz = time.perf_counter()
                doc_d = B.hashes_yaml_digest(str(t1_), dz)
                best_d = min(best_d, time.perf_counter() - tz)
            hx = {p: O.sha512sum(p) for p in tree_paths(t1_, 1000)}
            hx[str(root / "streamed.tar.gz")] = dz.hex()
            assert doc_d == O.write_hashes(str(t1_), str(root / "streamed.tar.gz"), hasher=lambda p: hx[os.fsdecode(p)])
            e2e_tree["cfg1_archive_hashed_while_written"] = {
                "flow": "tar | gzip -9 -> io.MultiWriter(file, snapgpu_hasher) -> snapgpu_hashes_yaml_digest (INTEGRATION.md 3b)",
                "tar_gzip_alone_ms": gzip_only_ms, "tar_gzip_with_hasher_ms": written_ms, "hasher_sum_after_last_write_ms": sum_ms,
                "write_hashes_with_digest_ms": best_d * 1e3, "archive_bytes": os.path.getsize(root / "streamed.tar.gz"),
                "yaml_identical_to_oracle": True,
                "note": "the chain runs under the compressor: what writeHashes itself adds to the build is the last line"}
            r1["note"] = ("the archive is ONE SHA-512 chain: ~70 MB/s on a GPU lane pair against ~0.8 GB/s on a CPU core, so a "
                          "package's data.tar.gz, not its tree, sets writeHashes' time on the GPU -- which is why INTEGRATION.md "
                          "hashes it while gzip writes it (snapgpu_hasher) instead of re-reading the finished file")
            e2e_tree["cfg1_real_archive"] = r1
        except (OSError, subprocess.CalledProcessError, MemoryError) as exc:      # no room on tmpfs, no tar, ...: the
            tree_error = f"{type(exc).__name__}: {exc}"                         # other legs of the line still stand
            e2e_tree["error"] = tree_error
            log("e2e_tree leg failed:", tree_error)
        finally:
            shutil.rmtree(root, ignore_errors=True)

    # ---- roofline of the dominant kernel (rank 0's launch; every rank runs the same shape) ---
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    nominal_peak = sms * INT32_LANES_PER_SM * peaks["sm_max_mhz"] * 1e6 / 1e12           # T int32-instr/s
    achieved = ALGO_INSTR_PER_BLOCK * nblocks * 32 / 32 / (kernel_ms * 1e-3) / 1e12     # T int32-instr/s
    alu = device.pipe_microbench(2, 16)          # SHF: the instruction the kernel issues most
    mix = device.pipe_microbench(7, 16)
    measured_alu_peak = alu["warp_inst_per_clk_per_sm"] * 32 * sms * alu["sm_clock_mhz"] * 1e6 / 1e12
    tr_sha = ncu_traffic("sha512") if args.workload == "cfg2" else None
    roofline = {
        "bound": "int_alu", "kernel": "sha512_segments_kernel_v2" if not args.variant or args.variant == 1 else "sha512_segments_kernel",
        # ptxas moves some adds to the FMA pipe (IMAD.X), so the honest utilisation figure is
        # ncu's ALU-pipe busy share of the same launch, which the paper-count `frac` below overstates a little
        "alu_pipe_util_ncu": (tr_sha or {}).get("alu_pipe_util", 0.923 if args.workload == "cfg2" else None),
        "alu_pipe_util_ncu_source": (tr_sha or {}).get("source", "profiles/r01_bench_sha512_ncu_full.txt (replayed)") if args.workload == "cfg2" else None,
        "achieved": achieved, "peak": nominal_peak, "unit": "T int32-instr/s", "frac": achieved / nominal_peak,
        "peak_source": f"{sms} SMs x {INT32_LANES_PER_SM} int32 lanes/clk x sm_max_mhz {peaks['sm_max_mhz']} ({peaks['_source']})",
        "algorithmic_instr_per_block": ALGO_INSTR_PER_BLOCK, "blocks_per_launch": nblocks,
        "kernel_ms": kernel_ms, "padded_gbs": nblocks * 128 / (kernel_ms * 1e-3) / 1e9,
        "measured_alu_pipe": {"warp_inst_per_clk_per_sm": alu["warp_inst_per_clk_per_sm"], "sm_clock_mhz": alu["sm_clock_mhz"],
                              "peak_T_instr_s": measured_alu_peak, "frac": achieved / measured_alu_peak if measured_alu_peak else None},
        "sha_mix_probe_warp_inst_per_clk_per_sm": mix["warp_inst_per_clk_per_sm"],
        "traffic": (tr_sha or {}).get("dram_bytes"),
        "traffic_detail": tr_sha,
        "algorithmic_bytes": file_bytes + 64 * len(lengths) + 36 * len(lengths),
        "hbm": {"achieved_gbs": file_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peaks["hbm_gbs"],
                "frac": file_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
    }

    # ---- compare kernel, config 4 (HBM-bound) -------------------------------------------------
    cmp_res = None
    if not args.no_cmp:
        npairs = 10_000
        cl = np.full(npairs, 1 << 20, dtype=np.uint64)
        co, ctotal = synth.layout(cl)
        da = torch.empty(ctotal, dtype=torch.uint8, device=dev)
        device.synth_fill_device(da, co, cl)
        db = da.clone()
        rng = np.random.default_rng(synth.SEED)
        differ = rng.choice(npairs, 100, replace=False)
        pos = rng.integers(0, 1 << 20, 100)
        idx = torch.from_numpy((co[differ].astype(np.int64) + pos)).to(dev)
        db[idx] ^= 0x01
        algo_bytes = int((npairs - 100) * 2 * (1 << 20) + sum(2 * 16384 * (int(p) // 16384 + 1) for p in pos))
        d_eq = torch.empty(npairs, dtype=torch.uint8, device=dev)
        for _ in range(3):
            device.cmp_batch_device(da, db, co, cl, d_eq)
        torch.cuda.synchronize()
        N.reset_stats()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        csteps = 10
        for _ in range(csteps):
            device.cmp_batch_device(da, db, co, cl, d_eq)
        c1.record()
        torch.cuda.synchronize()
        sc = N.stats()
        cms = sc.cmp_kernel_ms_sum / max(sc.cmp_kernel_timed, 1)
        eq = d_eq.cpu().numpy()
        assert sorted(np.nonzero(eq == 0)[0].tolist()) == sorted(differ.tolist()), "cmp flags wrong"
        launches += int(sc.kernel_launches)
        cmp_res = {"workload": "config 4: 10,000 pairs x 1 MiB, 1% differ by one byte", "bound": "hbm",
                   "achieved": algo_bytes / (cms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": algo_bytes / (cms * 1e-3) / 1e9 / peaks["hbm_gbs"], "kernel_ms": cms,
                   "step_ms": c0.elapsed_time(c1) / csteps, "algorithmic_bytes": algo_bytes,
                   "traffic": (ncu_traffic("cmp") or {}).get("dram_bytes"), "traffic_detail": ncu_traffic("cmp"),
                   "peak_source": peaks["_source"]}
        # end to end: the first 1,000 pairs from pinned host memory through snapgpu_cmp_batch
        if not args.no_e2e:
            ne = 1000
            nbytes_e = int(co[ne - 1] + cl[ne - 1])
            ha, pa = pinned_array(nbytes_e + 64)
            hb, pb = pinned_array(nbytes_e + 64)
            torch.from_numpy(ha)[:nbytes_e].copy_(da[:nbytes_e].cpu())
            torch.from_numpy(hb)[:nbytes_e].copy_(db[:nbytes_e].cpu())
            want_eq = eq[:ne]
            got_eq = helpers.cmp_batch(ha, hb, co[:ne], cl[:ne])
            assert np.array_equal(got_eq, want_eq), "end-to-end cmp flags differ from the device-resident run"
            N.reset_stats()
            t0 = time.perf_counter()
            for _ in range(3):
                helpers.cmp_batch(ha, hb, co[:ne], cl[:ne])
            dte = (time.perf_counter() - t0) / 3
            se = N.stats()
            cmp_res["e2e"] = {"value": 2 * int(cl[:ne].sum()) / dte / 1e9, "unit": "GB/s of compared bytes (both streams)",
                              "pairs": ne, "ms": dte * 1e3, "h2d_bytes_per_step": int(se.h2d_bytes // 3),
                              "d2h_bytes_per_step": int(se.d2h_bytes // 3),
                              "bound": "PCIe host-to-device copy of both streams"}
            launches += int(se.kernel_launches)
            N.lib().snapgpu_free_pinned(pa)
            N.lib().snapgpu_free_pinned(pb)
            del ha, hb
        del da, db

    # ---- tail latency: the serial chain of ONE long file (north_star: reported separately) ----
    tail = None
    if not args.no_tail:
        tl = np.array([16 << 20], dtype=np.uint64)
        to, ttotal = synth.layout(tl)
        dt_ = torch.empty(ttotal, dtype=torch.uint8, device=dev)
        device.synth_fill_device(dt_, to, tl)
        dgt = torch.empty((1, 64), dtype=torch.uint8, device=dev)
        # through the default long-file bin (2: a lane pair per chain), its one-lane form (1) and,
        # with the bin off (0), the batched kernel -- the two comparison points
        by_mode = {}
        for mode in (2, 1, 0):
            N.set_option("long_kernel", mode)
            device.sha512_batch_device(dt_, to, tl, dgt)
            torch.cuda.synchronize()
            N.reset_stats()
            device.sha512_batch_device(dt_, to, tl, dgt)
            torch.cuda.synchronize()
            by_mode[mode] = N.stats().sha512_kernel_ms_sum
            launches += 2
        N.set_option("long_kernel", 2)
        tms = by_mode[2]
        tblocks = int(synth.blocks(tl)[0])
        tail = {"what": "one 16 MiB file alone on the GPU: a single SHA-512 chain cannot be split, so this is the "
                        "latency floor of the longest file of a batch",
                "kernel": "sha512_pair_kernel (one chain on a lane pair)",
                "file_bytes": int(tl[0]), "blocks": tblocks, "kernel_ms": tms, "us_per_block": tms * 1e3 / tblocks,
                "mb_per_s_per_stream": int(tl[0]) / (tms * 1e-3) / 1e6,
                "extrapolated_s_per_GiB": tms * 1e-3 * (1 << 30) / int(tl[0]),
                "one_lane_kernel_us_per_block": by_mode[1] * 1e3 / tblocks,
                "batched_kernel_us_per_block": by_mode[0] * 1e3 / tblocks,
                "sha512_hex_prefix": dgt.cpu().numpy().tobytes().hex()[:16]}
        del dt_

    # ---- CPU baseline: the oracle on this box's cores, bounded sample ------------------------
    cpu = None
    if not args.no_cpu and host_view is not None:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        use_ossl = bool(O.lib().oracle_have_openssl())
        sl = lengths if args.workload != "cfg3" else lengths[:50_000]
        so = offsets[: len(sl)]
        sbytes = int(sl.sum())
        t0 = time.perf_counter()
        ref1 = O.sha512_batch(host_view, so, sl, 1, use_ossl)
        dt1 = time.perf_counter() - t0
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            refn = O.sha512_batch(host_view, so, sl, cores, use_ossl)
            best = min(best, time.perf_counter() - t0)
        assert np.array_equal(ref1, digest_host[: len(sl)]) and np.array_equal(refn, ref1), \
            "GPU digests differ from the oracle"
        cpu = {"value": sbytes / best / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
               "value_1core": sbytes / dt1 / 1e9, "files_per_s": len(sl) / best,
               "sample": f"the whole batch ({len(sl)} files, {sbytes} bytes), in memory, best of 3 on {cores} threads; "
                         f"1-core figure = what the reference's single goroutine does",
               "impl": ("C restatement of helpers.Sha512sum's loop, OpenSSL SHA-512 block function (not Go)"
                        if use_ossl else "C restatement, scalar block function"),
               "bit_exact_with_gpu": True}

    if cpu is not None and e2e_tree is not None and "cfg2" in e2e_tree:
        c2 = e2e_tree["cfg2"]
        cpu.update({"tree_1core_gbs": c2["cpu_1core_gbs"], "tree_allcores_gbs": c2["cpu_allcores_gbs"],
                    "tree_1core_files_per_s": c2["cpu_1core_files_per_s"], "tree_allcores_files_per_s": c2["cpu_allcores_files_per_s"],
                    "tree_note": "config 2 as a tree on tmpfs: Sha512sum looped over the files (see e2e_tree.cfg2)"})

    # ---- one process, every GPU of the box: what a cgo caller gets (N > 1 only) ---------------
    inproc = None
    if world > 1 and not args.no_inprocess and host_view is not None and e2e is not None:
        def inprocess_leg():
            N.init(list(range(world)))                      # rank 0 re-binds: devices 0..N-1, the other ranks idle
            stride = (total_alloc + 4095) & ~4095
            big, big_ptr = pinned_array(world * stride)
            lens_all = np.tile(lengths, world)
            offs_all = np.concatenate([offsets + np.uint64(r * stride) for r in range(world)])
            for r in range(world):                          # N copies of rank 0's batch: every copy hashes to rank 0's digests
                big[r * stride: r * stride + total_alloc] = host_view
            N.lib().snapgpu_free_pinned(host_ptr)
            out_all = np.empty((len(lens_all), 64), dtype=np.uint8)
            for _ in range(2):
                helpers.sha512_batch(big, offs_all, lens_all, out=out_all)
            assert np.array_equal(out_all.reshape(world, len(lengths), 64), np.broadcast_to(digest_host, (world, len(lengths), 64))), \
                "in-process digests differ from the per-rank run"
            N.reset_stats()
            isteps = 5
            t0 = time.perf_counter()
            for _ in range(isteps):
                helpers.sha512_batch(big, offs_all, lens_all, out=out_all)
            dti = (time.perf_counter() - t0) / isteps
            sti = N.stats()
            secs = ctypes.c_double()
            pb = min(stride, 1 << 30) & ~4095
            N.check(N.lib().snapgpu_h2d_probe(big_ptr, pb, 4, ctypes.byref(secs)))
            ip_peak = world * pb * 4 / secs.value / 1e9
            ip_gbs = world * file_bytes / dti / 1e9
            return int(sti.kernel_launches), {"what": f"ONE process bound to {world} devices (snapgpu_init), one pinned buffer holding {world} config-2 batches, one "
                              f"snapgpu_sha512_batch call per step; the other ranks idle",
                      "gbs": ip_gbs, "files_per_s": world * len(lengths) / dti, "ms_per_step": dti * 1e3,
                      "frac_of_torchrun": ip_gbs / e2e["value"], "h2d_peak_gbs": ip_peak, "frac_of_h2d_peak": ip_gbs / ip_peak,
                      "h2d_bytes_per_step": int(sti.h2d_bytes // isteps), "bit_exact_with_per_rank_run": True}
        try:
            more, inproc = inprocess_leg()
            launches += more
        except (RuntimeError, MemoryError, OSError) as exc:      # resources; a digest mismatch still raises
            inproc = {"error": f"{type(exc).__name__}: {exc}"}
            log("inprocess leg failed:", inproc["error"])

    total_bytes = world * file_bytes
    line = {
        "metric": "hashes.yaml SHA-512 throughput", "value": total_bytes / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "files_per_s": world * len(lengths) / (ms_step * 1e-3),
        "config": config_for(args.workload, world),
        "run": {"sha_variant": int(args.variant or 0), "host_placement": placement,
                "parity": f"every rank checked its digests against the oracle: {parity_total} files over {world} rank(s)"},
        "clocks": clocks, "e2e": e2e, "e2e_tree": e2e_tree, "cfg5": cfg5, "inprocess": inproc, "gpu_launches": launches,
        "roofline": roofline, "roofline_cmp": cmp_res, "tail_latency": tail, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=host_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
