#!/usr/bin/env python
"""Pretty-print a sweep .jsonl written by tools/gpu_sweep.py."""
import json
import sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    if d["what"] == "pipe":
        print(f"{d['name']:24s} w={d['warps_per_sm']:2d} warp-inst/clk/SM={d['warp_inst_per_clk_per_sm']:.3f} "
              f"ms={d['elapsed_ms']:.3f} MHz={d['sm_clock_mhz']:.0f}")
    elif d["what"] == "sha":
        print(f"{d['workload']:15s} v={d['variant']} R={d['ctas_per_sm']} step={d['step_ms']:.3f} k={d['kernel_ms']:.3f} "
              f"GB/s={d['file_gbs']:.1f} frac={d['frac_nominal']:.3f} ok={d['same_digests']}")
    else:
        print(d)
