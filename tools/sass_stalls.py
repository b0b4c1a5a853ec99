#!/usr/bin/env python
"""Single-warp issue model of a kernel's hot loop from its SASS control words.

usage: sass_stalls.py <lib.so|.cubin> <function-substring> [--loop]

Every sm_100 instruction carries its scheduling control in the upper word: bits [105:109) of
the 128-bit encoding are the stall count (cycles before the next instruction of the same warp
may issue).  Summing them over the biggest backward-branch loop gives the cycles ONE warp
needs per iteration when nothing else runs on its SM sub-partition (fixed-latency code; memory
scoreboard waits excluded) -- the figure that bounds the serial chain of a long file."""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
funcs = re.findall(r"Function : (\S+)", text)
fn = [f for f in funcs if pat in f]
if not fn:
    sys.exit(f"no function matching {pat}")
fn = fn[0]
body = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True, check=True).stdout
lines = body.splitlines()
insts = []   # (addr, opcode, text, stall, yield)
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            stall = (hi >> 41) & 0xF
            yld = (hi >> 45) & 1
            t = m.group(2).strip()
            op = t.split()[1] if t.startswith("@") else t.split()[0]
            insts.append((int(m.group(1), 16), op, t, stall, yld))
            i += 2
            continue
    i += 1
print(f"function {fn}: {len(insts)} instructions")
# every backward branch is a loop; report all, detail the biggest
loops = []
for k, (addr, op, t, st, y) in enumerate(insts):
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
_alu = re.compile(r"^(IADD3|LOP3|SHF|PRMT|SEL|ISETP|IMNMX|VIADD|LEA|MOV|VIMNMX|BMSK|SGXT|FLO|POPC|IABS|PLOP3|VOTE|R2P|P2R)")
for lo_, hi_ in sorted(loops):
    body = [x for x in insts if lo_ <= x[0] <= hi_]
    print(f"  loop 0x{lo_:x}..0x{hi_:x}: {len(body)} instr, ALU {sum(1 for x in body if _alu.match(x[1]))}, "
          f"IMAD* {sum(1 for x in body if x[1].startswith('IMAD'))}, stall sum {sum(x[3] for x in body)}")
best = None
for k, (addr, op, t, st, y) in enumerate(insts):
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
if best is None:
    sys.exit("no loop found")
loop = [x for x in insts if best[0] <= x[0] <= best[1]]
alu = re.compile(r"^(IADD3|LOP3|SHF|PRMT|SEL|ISETP|IMNMX|VIADD|LEA|MOV|VIMNMX|BMSK|SGXT|FLO|POPC|IABS|PLOP3|VOTE|R2P|P2R)")
fma = re.compile(r"^(IMAD|FFMA|FMUL|FADD)")
n_alu = sum(1 for x in loop if alu.match(x[1]))
n_fma = sum(1 for x in loop if fma.match(x[1]) and "WIDE" not in x[1])
n_wide = sum(1 for x in loop if "WIDE" in x[1])
stalls = sum(x[3] for x in loop)
hist = Counter(x[3] for x in loop)
print(f"loop 0x{best[0]:x}..0x{best[1]:x}: {len(loop)} instructions, ALU {n_alu}, FMA {n_fma}, WIDE {n_wide}, "
      f"other {len(loop) - n_alu - n_fma - n_wide}")
print(f"sum of stall counts = {stalls} cycles/iteration (single warp); ALU-pipe bound = {2 * n_alu}")
print("stall histogram:", dict(sorted(hist.items())))
if "--range" in sys.argv:              # --range 0xb70 0x2520: the control words of an address range
    k = sys.argv.index("--range")
    lo_, hi_ = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
    for x in insts:
        if lo_ <= x[0] <= hi_:
            print(f"{x[0]:06x} s{x[3]:2d} {'Y' if x[4] else ' '} {x[2]}")
elif "--dump" in sys.argv:
    for x in loop:
        print(f"{x[0]:06x} s{x[3]:2d} {'Y' if x[4] else ' '} {x[2]}")
