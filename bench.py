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
             H2D -> kernel -> D2H digests), every step
  roofline   the SHA-512 kernel against the integer-ALU issue peak (it is ALU-bound, not
             HBM-bound; DESIGN.md "Roofline"), plus the HBM figures for context
  cpu_baseline  the oracle (C restatement, OpenSSL block function) on this box's cores
Other workloads (cfg1, cfg3, cfg4, cfg5) are parity-test cases; `--workload` runs them for
the tables in profiles/ and DESIGN.md.
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
            best = dict(d[kernel], source=p.name)
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


def pinned_array(nbytes: int):
    from snappy_b200 import _native as N
    p = N.lib().snapgpu_alloc_pinned(nbytes)
    if not p:
        raise RuntimeError("snapgpu_alloc_pinned failed: " + N.last_error())
    buf = (ctypes.c_uint8 * nbytes).from_address(p)
    return np.frombuffer(buf, dtype=np.uint8), p


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU path (oracle port; the reference itself is Go and
    cannot be built in this image) on the host cores, same config/metric/unit."""
    if rank != 0:
        return
    from oracle import oracle as O
    from snappy_b200 import synth
    O.build()
    lengths, first, desc = workload(args.workload, 0)
    if args.workload == "cfg3":
        lengths = lengths[:50_000]          # bounded sample: the small-file phase
    data, offsets, lengths = synth.make_host_batch(lengths, first_index=first)
    cores = os.cpu_count() or 1
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
    line = {
        "impl": "reference", "metric": "hashes.yaml SHA-512 throughput", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "files_per_s": len(lengths) / dt,
        "config": {"workload": desc, "files": int(len(lengths)), "bytes": total},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample,
                         "value_1core": total / dt1 / 1e9,
                         "note": ("C restatement of helpers.Sha512sum's loop with OpenSSL's SHA-512 block function "
                                  "(not Go: no Go toolchain in this image)" if use_ossl else
                                  "C restatement of helpers.Sha512sum with its own scalar block function"),
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
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
        e2e = {"value": world * file_bytes / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3,
               "files_per_s": world * len(lengths) / dt,
               "h2d_bytes_per_step": int(st2.h2d_bytes // e2e_steps), "d2h_bytes_per_step": int(st2.d2h_bytes // e2e_steps),
               "bound": "PCIe host-to-device copy (pinned host memory; chunks of 64 MiB, 256 MiB, then 1 GiB, double buffered)",
               "timing": "host wall clock around the synchronous C-ABI call (digests are in host memory on return), max over ranks",
               "h2d_gbs_per_gpu": st2.h2d_bytes / e2e_steps / dt / 1e9}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (rank 0's launch; every rank runs the same shape) ---
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    nominal_peak = sms * INT32_LANES_PER_SM * peaks["sm_max_mhz"] * 1e6 / 1e12           # T int32-instr/s
    achieved = ALGO_INSTR_PER_BLOCK * nblocks * 32 / 32 / (kernel_ms * 1e-3) / 1e12     # T int32-instr/s
    alu = device.pipe_microbench(2, 16)          # SHF: the instruction the kernel issues most
    mix = device.pipe_microbench(7, 16)
    measured_alu_peak = alu["warp_inst_per_clk_per_sm"] * 32 * sms * alu["sm_clock_mhz"] * 1e6 / 1e12
    roofline = {
        "bound": "int_alu", "kernel": "sha512_segments_kernel_v2" if not args.variant or args.variant == 1 else "sha512_segments_kernel",
        "achieved": achieved, "peak": nominal_peak, "unit": "T int32-instr/s", "frac": achieved / nominal_peak,
        "peak_source": f"{sms} SMs x {INT32_LANES_PER_SM} int32 lanes/clk x sm_max_mhz {peaks['sm_max_mhz']} ({peaks['_source']})",
        "algorithmic_instr_per_block": ALGO_INSTR_PER_BLOCK, "blocks_per_launch": nblocks,
        "kernel_ms": kernel_ms, "padded_gbs": nblocks * 128 / (kernel_ms * 1e-3) / 1e9,
        "measured_alu_pipe": {"warp_inst_per_clk_per_sm": alu["warp_inst_per_clk_per_sm"], "sm_clock_mhz": alu["sm_clock_mhz"],
                              "peak_T_instr_s": measured_alu_peak, "frac": achieved / measured_alu_peak if measured_alu_peak else None},
        "sha_mix_probe_warp_inst_per_clk_per_sm": mix["warp_inst_per_clk_per_sm"],
        "traffic": (ncu_traffic("sha512") or {}).get("dram_bytes") if args.workload == "cfg2" else None,
        "traffic_detail": ncu_traffic("sha512") if args.workload == "cfg2" else None,
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

    total_bytes = world * file_bytes
    line = {
        "metric": "hashes.yaml SHA-512 throughput", "value": total_bytes / (ms_step * 1e-3) / 1e9, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "files_per_s": world * len(lengths) / (ms_step * 1e-3),
        "config": {"workload": desc, "files_per_gpu": int(len(lengths)), "bytes_per_gpu": file_bytes,
                   "blocks_per_gpu": nblocks, "sharding": f"file list sharded over {world} GPU(s), no collective",
                   "cache": f"inputs ({file_bytes / 1e9:.2f} GB per GPU) larger than the 126 MB L2; no flush needed",
                   "sha_variant": int(args.variant or 0),
                   "host_placement": placement},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "roofline_cmp": cmp_res, "tail_latency": tail, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
