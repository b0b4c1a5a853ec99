#!/usr/bin/env python
"""DRAM traffic per launch from ncu --set full captures -> profiles/rNN_traffic.json.

usage: ncu_traffic.py out.json name=report.ncu-rep [name=report.ncu-rep ...]
bench.py reads the newest profiles/r*_traffic.json for the `traffic` key of its roofline
objects (names: sha512, cmp).  The captures must come from the bench command itself, so that
a launch is the launch bench.py times."""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        def val(k):
            return float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        alu = "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"
        launches.append({"kernel": r[col["Kernel Name"]].split("(")[0], "dram_bytes_read": rd, "dram_bytes_write": wr,
                         "alu_pipe_pct_elapsed": float(r[col[alu]].replace(",", "")) if alu in col else None,
                         "duration_under_ncu_ms": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                         {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[units[col["gpu__time_duration.sum"]]]})
    n = len(launches)
    out[name] = {"dram_bytes": sum(l["dram_bytes_read"] + l["dram_bytes_write"] for l in launches) / n,
                 "dram_bytes_read": sum(l["dram_bytes_read"] for l in launches) / n,
                 "dram_bytes_write": sum(l["dram_bytes_write"] for l in launches) / n,
                 "launches_captured": n, "kernel": launches[0]["kernel"], "report": rep.split("/")[-1],
                 "alu_pipe_util": (sum(l["alu_pipe_pct_elapsed"] for l in launches) / n / 100.0
                                   if all(l["alu_pipe_pct_elapsed"] is not None for l in launches) else None),
                 "duration_under_ncu_ms": sum(l["duration_under_ncu_ms"] for l in launches) / n}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
