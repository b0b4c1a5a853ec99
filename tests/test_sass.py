"""Static checks of the machine code (cuobjdump, no GPU): the instruction mix the design
relies on is what ptxas actually emitted."""
import re
import shutil
import subprocess
from collections import Counter

import pytest

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def sass(native):
    text = subprocess.run(["cuobjdump", "-sass", str(native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    return funcs


def find(funcs, *needles):
    hits = [f for f in funcs if all(n in f for n in needles)]
    assert len(hits) == 1, (needles, hits)
    return funcs[hits[0]]


def count(c, prefix):
    return sum(v for k, v in c.items() if k.startswith(prefix))


def test_only_sm100a_code(native):
    out = subprocess.run(["cuobjdump", "-lelf", str(native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_[89]\d", out)


def test_default_kernel_is_compact_and_staged(sass):
    """Variant 0: 16-round loop (fits the instruction cache), cp.async staging, no spills."""
    c = find(sass, "sha512_segments_kernel_v2", "ILi0ELi3ELb1E")
    total = sum(c.values())
    assert total < 2600                                     # ~33 KB of code, was ~65 KB fully unrolled
    assert count(c, "LDGSTS") >= 16                         # cp.async: 8 full + 8 tail copies per block
    assert count(c, "LDS") >= 8 and count(c, "LDG.E.128") <= 8          # LDG only for descriptors / state
    assert count(c, "SHF.R.W") >= 16 * 12 * 2 + 16 * 10     # two round groups + one schedule group
    assert count(c, "LDL") == 0 and count(c, "STL") == 0
    assert count(c, "IMAD.WIDE") < 16 and count(c, "IMAD.HI") == 0  # both measured at half rate (address and
                                                                    # balance arithmetic only, none in the round loop)


def test_any_alignment_kernel_is_staged_too(sass):
    """Unaligned input: nine zero-filling cp.async per block, 33 word loads from the file's own phase, and
    the same 32 byte-permutes as the aligned form (realign + byte swap in one PRMT); no spills."""
    c = find(sass, "sha512_segments_kernel_v2", "ILi0ELi3ELb0E")
    assert count(c, "LDGSTS") >= 9 and count(c, "LDS") >= 33
    assert 32 <= count(c, "PRMT") <= 80
    assert count(c, "LDL") == 0 and count(c, "STL") == 0
    assert sum(c.values()) < 2700


def test_fma_add_variant_mix(sass):
    """Variant 3: rotates are funnel shifts, logic is LOP3, 64-bit adds sit on the FMA pipe."""
    c = find(sass, "sha512_segments_kernel", "Li127ELi7ELb1")
    assert count(c, "SHF.R.W") == 1600                      # 80*12 + 64*10 funnel shifts
    assert 896 <= count(c, "LOP3") <= 1000                  # 80*8 + 64*4 (+ padding logic)
    assert count(c, "IMAD.WIDE") >= 750                     # 7*80 + 3*64 + 8 = 760 wide accumulates
    assert count(c, "IADD3") < 60                           # the adds really left the ALU pipe
    assert count(c, "LDL") == 0 and count(c, "STL") == 0    # no spills
    assert sum(v for k, v in c.items() if k.startswith("LDG") and ".128" in k) >= 8   # 128-bit message loads


def test_alu_add_variant_mix(sass):
    c = find(sass, "sha512_segments_kernel", "Li0ELi0ELb1")
    assert count(c, "SHF.R.W") == 1600
    assert count(c, "IADD3") >= 700 and count(c, "IMAD.WIDE") < 10
    assert count(c, "LDL") == 0 and count(c, "STL") == 0


def test_cmp_kernel_uses_128_bit_loads(sass):
    c = find(sass, "cmp_pairs_kernel", "Lb1")
    wide = sum(v for k, v in c.items() if k.startswith("LDG") and ".128" in k)
    assert wide >= 8                                        # 4 rows x 2 streams, 128 bits per lane
    assert all(".NA." in k for k in c if k.startswith("LDG") and ".128" in k)   # streamed past L1
    assert count(c, "VOTE") >= 1


@pytest.mark.parametrize("kind,opcode,least", [(0, "IADD3", 128), (1, "LOP3", 128), (2, "SHF.R.W", 128),
                                               (3, "IMAD", 128), (4, "IMAD.WIDE", 128)])
def test_probe_kernels_issue_what_they_claim(sass, kind, opcode, least):
    c = find(sass, "pipe_probe_kernel", f"ILi{kind}E")
    assert count(c, opcode) >= least, c.most_common(6)


def test_pair_kernel_round_loop(native):
    """sha512_pair_kernel, mailbox form (the default): the 16-round loop of the consumer has ~17 ALU
    instructions per round (28 in the one-lane consumer) and talks through shared memory only.  The
    exchange is safe because the loop body is ONE convergence check (the BRA.DIV of __syncwarp) followed
    by branch-free code in which the volatile accesses keep their order -- load, load, store per round, so
    every mailbox store sits ahead of the partner's load in the next round.  No spills."""
    text = subprocess.run(["cuobjdump", "-sass", str(native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    body = [p for p in text.split("Function : ") if p.startswith("_ZN7snapgpu18sha512_pair_kernelILb1ELb0E")]
    assert len(body) == 1
    insts = []
    for line in body[0].splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            insts.append((int(m.group(1), 16), m.group(2), m.group(3)))
    assert not any(op.startswith(("LDL", "STL")) for _, op, _ in insts)
    # loops = backward branches; the round loop is the one whose body holds 16 STS.64 and 32 LDS.64
    loops = []
    for addr, op, rest in insts:
        if op.startswith("BRA") and not op.startswith("BRA.DIV"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
    alu = re.compile(r"^(IADD3|LOP3|SHF|PRMT|SEL|ISETP|VIADD|LEA|MOV|IMNMX|VIMNMX)")
    found = False
    for lo, hi in loops:
        ops = [op for a, op, _ in insts if lo <= a <= hi]
        if (sum(op.startswith("STS.64") for op in ops) == 16 and sum(op.startswith("LDS.64") for op in ops) == 32
                and len(ops) < 500):
            found = True
            assert sum(op.startswith("SHF.R.W") for op in ops) == 16 * 6
            assert not any(op.startswith(("SHFL", "WARPSYNC", "BAR", "BSSY", "BSYNC", "CALL", "RET", "EXIT")) for op in ops)
            branches = [op for op in ops if op.startswith(("BRA", "BRX", "JMP"))]
            assert sorted(branches) == ["BRA", "BRA.DIV"], branches          # the back edge and one convergence check
            assert ops.index("BRA.DIV") < next(i for i, op in enumerate(ops) if op.startswith(("LDS", "STS")))
            mem = ["L" if op.startswith("LDS") else "S" for op in ops if op.startswith(("LDS.64", "STS.64"))]
            assert "".join(mem) == "LLS" * 16, "".join(mem)
            n_alu = sum(bool(alu.match(op)) for op in ops)
            assert n_alu <= 16 * 17.5, n_alu
    assert found


def test_pair_kernel_shuffle_form(native):
    """The cross-check form (pair_form 1): the lanes exchange by SHFL.BFLY, two per round, and the round
    loop stores nothing to shared memory."""
    text = subprocess.run(["cuobjdump", "-sass", str(native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    body = [p for p in text.split("Function : ") if p.startswith("_ZN7snapgpu18sha512_pair_kernelILb1ELb1E")]
    assert len(body) == 1
    ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body[0])
    assert ops.count("SHFL.BFLY") >= 2 * 18 and not any(op.startswith(("LDL", "STL")) for op in ops)
