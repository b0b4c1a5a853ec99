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


def _pair_stretches(native, mangled_prefix):
    """(instructions, branch-free stretches) of one pair kernel: a stretch is a maximal run without a control
    instruction (BSSY only declares a reconvergence point and does not end one), kept with the control
    instruction in front of it and the one that ends it."""
    text = subprocess.run(["cuobjdump", "-sass", str(native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    body = [p for p in text.split("Function : ") if p.startswith(mangled_prefix)]
    assert len(body) == 1
    insts = []
    for line in body[0].splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            insts.append((int(m.group(1), 16), m.group(2), m.group(3)))
    assert not any(op.startswith(("LDL", "STL")) for _, op, _ in insts)          # no spills
    control = ("BRA", "BRX", "JMP", "WARPSYNC", "BAR", "BSYNC", "CALL", "RET", "EXIT", "SHFL", "NANOSLEEP")
    stretches, before, run = [], None, []
    for inst in insts:
        if inst[1].startswith(control):
            stretches.append((before, run, inst))
            before, run = inst, []
        else:
            run.append(inst)
    return insts, stretches


_PAIR_ALU = re.compile(r"^(IADD3|LOP3|SHF|PRMT|SEL|ISETP|VIADD|LEA|MOV|IMNMX|VIMNMX)")


def _mem_string(run):
    return "".join("L" if op.startswith("LDS") else "S" for _, op, _ in run if op.startswith(("LDS.64", "STS.64")))


def test_pair_kernel_round_loop(native):
    """sha512_pair_kernel, mailbox form, two regions per block (3+ files per CTA): the consumer's loop body
    is ONE convergence check (the BRA.DIV of __syncwarp) followed by 41 rounds of branch-free code in which
    the volatile accesses keep their order -- load, load, store per round, so every mailbox store sits
    ahead of the partner's load in the next round -- and the back edge.  ~17 ALU instructions per round
    (28 in the one-lane consumer), no shuffles, no spills."""
    _, stretches = _pair_stretches(native, "_ZN7snapgpu18sha512_pair_kernelILb1ELb0ELi2E")
    found = 0
    for before, run, after in stretches:
        if _mem_string(run) != "LLS" * 41:
            continue
        found += 1
        assert before is not None and before[1] == "BRA.DIV", before         # entered through the convergence point
        assert after[1] == "BRA" and int(re.search(r"0x([0-9a-f]+)", after[2]).group(1), 16) <= before[0], after
        ops = [op for _, op, _ in run]
        assert sum(op.startswith("SHF.R.W") for op in ops) == 41 * 6
        assert sum(bool(_PAIR_ALU.match(op)) for op in ops) <= 41 * 17.5
    assert found == 1


def test_pair_kernel_single_region(native):
    """One or two files per CTA: prologue and all 82 rounds of a block are ONE branch-free stretch behind the
    prologue's convergence point."""
    _, stretches = _pair_stretches(native, "_ZN7snapgpu18sha512_pair_kernelILb1ELb0ELi1E")
    found = 0
    for before, run, _ in stretches:
        mem = _mem_string(run)
        if not mem.endswith("LLS" * 82):
            continue
        found += 1
        assert before is not None and before[1] == "BRA.DIV", before
        assert len(mem) == 3 * 82 + 5, mem[:12]                              # the prologue: kin[0], two seeds, din[0], a seed
        ops = [op for _, op, _ in run]
        assert sum(op.startswith("SHF.R.W") for op in ops) >= 82 * 6
        assert sum(bool(_PAIR_ALU.match(op)) for op in ops) <= 82 * 17.5 + 60
    assert found == 1


def test_pair_kernel_shuffle_form(native):
    """The cross-check form (pair_form 1): the lanes exchange by SHFL.BFLY, two per round, and the round
    loop stores nothing to shared memory."""
    text = subprocess.run(["cuobjdump", "-sass", str(native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    body = [p for p in text.split("Function : ") if p.startswith("_ZN7snapgpu18sha512_pair_kernelILb1ELb1ELi2E")]
    assert len(body) == 1
    ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body[0])
    assert ops.count("SHFL.BFLY") >= 2 * 18 and not any(op.startswith(("LDL", "STL")) for op in ops)


def test_pair_cta_keeps_its_sm_to_itself(native):
    """A lane-pair CTA asks for 200 KB of dynamic shared memory (csrc/sha512_pair.cuh: kPairSmemBytes) so that no
    batched-kernel CTA fits beside it on the SM -- a second warp on the consumer's sub-partition would halve the
    chain's share of the ALU pipe.  Checked against the shared memory the built kernels really use."""
    text = subprocess.run(["cuobjdump", "-res-usage", str(native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    usage = {}
    name = None
    for line in text.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"SHARED:(\d+)", line)
        if m and name:
            usage[name] = int(m.group(1))
    # the shipped batched kernel (variants 0, 1, 5: cp.async staging through shared memory); the register-prefetch
    # variants 2-4 are comparison builds and use no shared memory
    batched = [v for k, v in usage.items() if "sha512_segments_kernel_v2" in k]
    pair_static = [v for k, v in usage.items() if "sha512_pair_kernel" in k]
    assert batched and pair_static
    sm_bytes, reserved = 228 * 1024, 1024                     # per SM; reserved per resident CTA
    pair_cta = 200 * 1024 + max(pair_static) + reserved
    assert pair_cta <= sm_bytes                               # the pair CTA itself fits
    assert pair_cta + min(batched) + reserved > sm_bytes      # ... and leaves no room for the smallest batched CTA
