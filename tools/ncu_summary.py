#!/usr/bin/env python
"""Text summary of one ncu report (.ncu-rep), for profiles/.

usage: ncu_summary.py <report.ncu-rep> [--top N] > profiles/<name>.txt

Prints, per profiled launch: duration, launch shape, registers, pipe utilisation (ALU / FMA /
LSU / issue slots), DRAM bytes (the `traffic` figure of bench.py's roofline object), warp
stall mix, and the N SASS instructions with the most stall samples (source page; needs the
library to be built with -lineinfo).
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12

RAW = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active, % of 64/SM"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("sm__cycles_elapsed.max.per_second", "SM clock during capture"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy, %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe, % of peak (active cycles)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "ALU pipe, % of peak (elapsed cycles)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe, % of peak"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe, % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe, % of peak"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tensor pipe, % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (busiest unit), %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate, %"),
    ("l1tex__t_bytes_pipe_lsu_mem_global_op_ldgsts_cache_access.sum", "LDGSTS (cp.async) bytes through L1"),
]
STALLS = ["selected", "wait", "no_instruction", "long_scoreboard", "short_scoreboard", "branch_resolving",
          "math_pipe_throttle", "mio_throttle", "lg_throttle", "not_selected", "dispatch_stall", "barrier"]


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return out


raw = list(csv.reader(io.StringIO(page("raw"))))
hdr, units = raw[0], raw[1]
col = {h: i for i, h in enumerate(hdr)}
print(f"# ncu summary of {rep.split('/')[-1]}  (ncu --set full --clock-control none; cold caches, replayed passes)")
for row in raw[2:]:
    print(f"\n## launch {row[col['ID']]}: {row[col['Kernel Name']][:110]}")
    for key, label in RAW:
        if key in col and row[col[key]] != "":
            print(f"  {label:48s} {row[col[key]]:>18s} {units[col[key]]}")
    print("  warp stall reasons (warps per issue-active cycle):")
    for s in STALLS:
        key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if key in col and row[col[key]] not in ("", "0"):
            print(f"    {s:28s} {float(row[col[key]]):8.4f}")

src = page("source")
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = {n: i for i, n in enumerate(rows[0])}
body = [r for r in rows[1:] if len(r) > h["# Samples"]]
tot = sum(int(r[h["# Samples"]] or 0) for r in body)
exe = sum(int(r[h["Instructions Executed"]] or 0) for r in body)
print(f"\n## source page: {len(body)} SASS instructions, {tot} stall samples, {exe} warp instructions executed")
ops = {}
for r in body:
    t = r[h["Source"]].split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    op = op.rstrip(";")
    ops[op] = ops.get(op, 0) + int(r[h["Instructions Executed"]] or 0)
print("  executed warp instructions by opcode (top 14):")
for op, n in sorted(ops.items(), key=lambda x: -x[1])[:14]:
    print(f"    {op:24s} {n:14d}  {100.0 * n / max(exe, 1):5.1f} %")
print(f"  top {top} instructions by stall samples:")
stall_cols = [n for n in rows[0] if n.startswith("stall_") and "Not Issued" not in n]
for r in sorted(body, key=lambda r: -int(r[h["# Samples"]] or 0))[:top]:
    why = sorted(((int(r[h[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    why = ", ".join(f"{c}={n}" for n, c in why if n)
    print(f"    {int(r[h['# Samples']]):6d}  {r[h['Source']].strip()[:70]:70s} {why}")
