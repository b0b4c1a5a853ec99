"""The lane-pair SHA-512 schedule of snappy_b200/csrc/sha512_pair.cuh, restated in Python with the same
structure (per-lane parameters, mailboxes, two-round skew, seeds, 82 iterations per block) and checked
against hashlib.  No GPU: this pins the index algebra the kernel relies on; the kernel itself is checked
against the oracle in test_gpu_parity.py::test_long_file_kernel."""
import hashlib
import struct

import numpy as np

M = (1 << 64) - 1
K = [
    0x428a2f98d728ae22, 0x7137449123ef65cd, 0xb5c0fbcfec4d3b2f, 0xe9b5dba58189dbbc, 0x3956c25bf348b538, 0x59f111f1b605d019,
    0x923f82a4af194f9b, 0xab1c5ed5da6d8118, 0xd807aa98a3030242, 0x12835b0145706fbe, 0x243185be4ee4b28c, 0x550c7dc3d5ffb4e2,
    0x72be5d74f27b896f, 0x80deb1fe3b1696b1, 0x9bdc06a725c71235, 0xc19bf174cf692694, 0xe49b69c19ef14ad2, 0xefbe4786384f25e3,
    0x0fc19dc68b8cd5b5, 0x240ca1cc77ac9c65, 0x2de92c6f592b0275, 0x4a7484aa6ea6e483, 0x5cb0a9dcbd41fbd4, 0x76f988da831153b5,
    0x983e5152ee66dfab, 0xa831c66d2db43210, 0xb00327c898fb213f, 0xbf597fc7beef0ee4, 0xc6e00bf33da88fc2, 0xd5a79147930aa725,
    0x06ca6351e003826f, 0x142929670a0e6e70, 0x27b70a8546d22ffc, 0x2e1b21385c26c926, 0x4d2c6dfc5ac42aed, 0x53380d139d95b3df,
    0x650a73548baf63de, 0x766a0abb3c77b2a8, 0x81c2c92e47edaee6, 0x92722c851482353b, 0xa2bfe8a14cf10364, 0xa81a664bbc423001,
    0xc24b8b70d0f89791, 0xc76c51a30654be30, 0xd192e819d6ef5218, 0xd69906245565a910, 0xf40e35855771202a, 0x106aa07032bbd1b8,
    0x19a4c116b8d2d0c8, 0x1e376c085141ab53, 0x2748774cdf8eeb99, 0x34b0bcb5e19b48a8, 0x391c0cb3c5c95a63, 0x4ed8aa4ae3418acb,
    0x5b9cca4f7763e373, 0x682e6ff3d6b2b8a3, 0x748f82ee5defb2fc, 0x78a5636f43172f60, 0x84c87814a1f0ab72, 0x8cc702081a6439ec,
    0x90befffa23631e28, 0xa4506cebde82bde9, 0xbef9a3f7b2c67915, 0xc67178f2e372532b, 0xca273eceea26619c, 0xd186b8c721c0c207,
    0xeada7dd6cde0eb1e, 0xf57d4f7fee6ed178, 0x06f067aa72176fba, 0x0a637dc5a2c898a6, 0x113f9804bef90dae, 0x1b710b35131c471b,
    0x28db77f523047d84, 0x32caab7b40c72493, 0x3c9ebe0a15c9bebc, 0x431d67c49c100d4c, 0x4cc5d4becb3e42b6, 0x597f299cfc657e2a,
    0x5fcb6fab3ad6faec, 0x6c44198c4a475817]
IV = [0x6a09e667f3bcc908, 0xbb67ae8584caa73b, 0x3c6ef372fe94f82b, 0xa54ff53a5f1d36f1,
      0x510e527fade682d1, 0x9b05688c2b3e6c1f, 0x1f83d9abfb41bd6b, 0x5be0cd19137e2179]


def rotr(x, r):
    return ((x >> r) | (x << (64 - r))) & M if r else x


def producer(block):
    """What the producer warp puts into the ring for one block: W[t] + K[t], t = 0..79."""
    w = list(struct.unpack(">16Q", block))
    for t in range(16, 80):
        s0 = rotr(w[t - 15], 1) ^ rotr(w[t - 15], 8) ^ (w[t - 15] >> 7)
        s1 = rotr(w[t - 2], 19) ^ rotr(w[t - 2], 61) ^ (w[t - 2] >> 6)
        w.append((w[t - 16] + s0 + w[t - 7] + s1) & M)
    return [(w[t] + K[t]) & M for t in range(80)]


# per-lane parameters of the one instruction stream: lane 0 = (e,f,g,h), lane 1 = (a,b,c,d)
ROT = [(4, 27, 14), (6, 11, 28)]        # Sigma1 / Sigma0 as rotr(x ^ rotr(x,p) ^ rotr(x,q), s), all below 32
MASK = [0, M]
MUL = [1, 0]


def pair_sigma(x, lane):
    p, q, s = ROT[lane]
    assert max(p, q, s) < 32
    return rotr(x ^ rotr(x, p) ^ rotr(x, q), s)


def pair_f(x0, x1, x2, lane):
    t = (~(x0 ^ x1) & M) if MASK[lane] else x0          # one LOP3 with the lane mask
    return (t & x1) | (~t & M & x2)                     # Ch(t, x1, x2): Ch on lane 0, Maj on lane 1


def compress_pair(state, ring):
    """One block: 82 iterations of the shared body; lane 1 runs two rounds behind lane 0."""
    st = [state[4:8], state[0:4]]
    T, A = {}, {}                                       # mailboxes: T[i] = T1 of round i, A[i] = a after round i-1
    a, b, c, d = st[1]
    A[-2], A[-1] = d, c                                 # seed_a
    T[-1] = (a - pair_sigma(b, 1) - pair_f(b, c, d, 1)) & M            # seed_t[1]: makes iteration 1 produce a
    seed0 = (b - pair_sigma(c, 1) - pair_f(c, d, 0, 1)) & M            # kept in a register: iteration 0 produces b
    win = [list(st[0]), [c, d, 0, 0]]
    kin = lambda lane, i: (ring[i] if i < 80 else 0xdeadbeef) if lane == 0 else T[i - 2]
    din = lambda lane, i: A[i - 2] if lane == 0 else 0
    D = [din(0, 0), din(1, 0)]
    PD = [(win[0][3] + kin(0, 0) + D[0]) & M, seed0]
    final0 = None
    for i in range(82):
        nxt = [(kin(l, i + 1), din(l, i + 1)) if i < 81 else (0, 0) for l in (0, 1)]   # loads come first
        out = []
        for l in (0, 1):
            s0, s1, s2, s3 = win[l]
            e = (pair_sigma(s0, l) + pair_f(s0, s1, s2, l) + PD[l]) & M
            out.append((e - D[l]) & M)
            PD[l] = (s2 * MUL[l] + nxt[l][0] + nxt[l][1]) & M
            D[l] = nxt[l][1]
            win[l] = [e, s0, s1, s2]
        T[i], A[i] = out
        if i == 79:
            final0 = list(win[0])
    efgh = [(x + y) & M for x, y in zip(st[0], final0)]
    abcd = [(x + y) & M for x, y in zip(st[1], win[1])]
    return abcd + efgh


def sha512_pair(msg):
    n = len(msg)
    msg = msg + b"\x80" + b"\0" * ((111 - n) % 128) + (8 * n).to_bytes(16, "big")
    s = list(IV)
    for o in range(0, len(msg), 128):
        s = compress_pair(s, producer(msg[o:o + 128]))
    return b"".join(x.to_bytes(8, "big") for x in s)


def test_factored_sigmas_and_maj_identity():
    rng = np.random.default_rng(1)
    for x, y, z in rng.integers(0, 1 << 63, (200, 3)).tolist():
        assert pair_sigma(x, 0) == rotr(x, 14) ^ rotr(x, 18) ^ rotr(x, 41)       # Sigma1
        assert pair_sigma(x, 1) == rotr(x, 28) ^ rotr(x, 34) ^ rotr(x, 39)       # Sigma0
        assert pair_f(x, y, z, 0) == ((x & y) ^ (~x & M & z))                    # Ch
        assert pair_f(x, y, z, 1) == ((x & y) ^ (x & z) ^ (y & z))               # Maj


def test_pair_schedule_matches_hashlib():
    rng = np.random.default_rng(2)
    for n in (0, 1, 3, 111, 112, 127, 128, 129, 1000, 4096, 5001):
        m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert sha512_pair(m) == hashlib.sha512(m).digest(), n


# ---- the shuffle form (sha512_pair.cuh, kShuffle): the lanes exchange through a warp shuffle -------------------

def compress_pair_shuffle(state, ring):
    """Same two-lane round, but what a lane needs from its partner -- always the value the partner
    produced in the PREVIOUS iteration -- arrives by an exchange at the end of every iteration
    (__shfl_xor_sync with lane mask 1) instead of through shared-memory mailboxes.  Lane 1 loads a
    zero where lane 0 loads W+K, so that both form PD = S2*mul + kw + r."""
    st = [state[4:8], state[0:4]]
    a, b, c, d = st[1]
    seed0 = (b - pair_sigma(c, 1) - pair_f(c, d, 0, 1)) & M            # iteration 0 of lane 1 produces b
    seed1 = (a - pair_sigma(b, 1) - pair_f(b, c, d, 1)) & M            # iteration 1 of lane 1 produces a
    win = [list(st[0]), [c, d, 0, 0]]
    # prologue exchange: lane 0 gets d and c from lane 1 (two shuffles per block)
    D = [d * MUL[0], 0 * MUL[1]]                       # D = (what the partner sent two iterations ago) * mul
    r = [c, seed1]                                     # what the partner "sent" before iteration 0
    PD = [(win[0][3] + ring[0] + D[0]) & M, seed0]
    final0 = None
    for i in range(82):
        kw = [ring[i + 1] if i + 1 < 80 else 0xdeadbeef, 0]            # lane 1 reads the zero array
        out = []
        for l in (0, 1):
            s0, s1, s2, s3 = win[l]
            e = (pair_sigma(s0, l) + pair_f(s0, s1, s2, l) + PD[l]) & M
            out.append((e - D[l]) & M)
            PD[l] = (s2 * MUL[l] + kw[l] + r[l]) & M
            D[l] = (r[l] * MUL[l]) & M
            win[l] = [e, s0, s1, s2]
        r = [out[1], out[0]]                           # the exchange
        if i == 79:
            final0 = list(win[0])
    efgh = [(x + y) & M for x, y in zip(st[0], final0)]
    abcd = [(x + y) & M for x, y in zip(st[1], win[1])]
    return abcd + efgh


def sha512_pair_shuffle(msg):
    n = len(msg)
    msg = msg + b"\x80" + b"\0" * ((111 - n) % 128) + (8 * n).to_bytes(16, "big")
    s = list(IV)
    for o in range(0, len(msg), 128):
        s = compress_pair_shuffle(s, producer(msg[o:o + 128]))
    return b"".join(x.to_bytes(8, "big") for x in s)


def test_shuffle_form_matches_hashlib():
    rng = np.random.default_rng(4)
    for n in (0, 1, 3, 111, 112, 127, 128, 129, 1000, 4096, 5001):
        m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert sha512_pair_shuffle(m) == hashlib.sha512(m).digest(), n


# ---- design check for the next step: no pipeline refill between blocks ----------------

def sha512_pair_continuous(msg):
    """Lane 1 stays two rounds behind lane 0 ACROSS block boundaries: 80 iterations per block instead of 82,
    no seeds after the first block.  Each lane feeds forward at its own time (lane 0 before iteration 80 n,
    lane 1 before 80 n + 2).  Lane 1 always publishes the plain a; what lane 0 needs as d in rounds 0..3 of a
    block is a + H of lane 1's previous block, so lane 0 adds Hd, Hc, Hb, Ha (handed over by lane 1) to its
    PD of those four rounds."""
    n = len(msg)
    msg = msg + b"\x80" + b"\0" * ((111 - n) % 128) + (8 * n).to_bytes(16, "big")
    total = 80 * (len(msg) // 128)
    ring = []
    for o in range(0, len(msg), 128):
        ring += producer(msg[o:o + 128])
    H = [list(IV[4:8]), list(IV[0:4])]
    T, A = {}, {}
    a, b, c, d = H[1]
    A[-2], A[-1] = d, c
    T[-1] = (a - pair_sigma(b, 1) - pair_f(b, c, d, 1)) & M
    seed0 = (b - pair_sigma(c, 1) - pair_f(c, d, 0, 1)) & M          # once per message, not per block
    win = [list(H[0]), [c, d, 0, 0]]
    handed = [0, 0, 0, 0]                                            # lane 1's H (a,b,c,d) as lane 0 last read it
    published = list(H[1])                                           # what lane 1 wrote into the hand-over words

    def kin(lane, g):
        return (ring[g] if g < total else 0) if lane == 0 else T[g - 2]

    def din(lane, g):
        return A[g - 2] if lane == 0 else 0

    def fix(g):
        """H word lane 0 adds to the PD of round g (rounds 0..3 of every block but the first)."""
        t = g % 80
        return handed[3 - t] if g >= 80 and t < 4 and g < total else 0   # t = 0: Hd, 1: Hc, 2: Hb, 3: Ha

    D = [din(0, 0), din(1, 0)]
    PD = [(win[0][3] + kin(0, 0) + D[0]) & M, seed0]
    for g in range(total + 2):
        if g % 80 == 0 and g > 0:                                   # lane 0 crosses a block boundary
            handed = list(published)                                # lane 1 wrote them 78 iterations ago
            old_h = H[0][3]
            win[0] = [(x + y) & M for x, y in zip(win[0], H[0])]
            H[0] = list(win[0])
            PD[0] = (PD[0] + old_h + fix(g)) & M                    # PD had been built from the un-fed h and the plain d
            D[0] = (D[0] + fix(g)) & M                              # ... and so had D, which is subtracted to publish T1
        if (g - 2) % 80 == 0 and g > 2:                             # lane 1 does, two iterations later
            win[1] = [(x + y) & M for x, y in zip(win[1], H[1])]
            H[1] = list(win[1])
            published = list(H[1])
        nxt = [(kin(l, g + 1), din(l, g + 1)) if g + 1 < total + 2 else (0, 0) for l in (0, 1)]
        out = []
        for l in (0, 1):
            s0, s1, s2, s3 = win[l]
            e = (pair_sigma(s0, l) + pair_f(s0, s1, s2, l) + PD[l]) & M
            out.append((e - D[l]) & M)
            PD[l] = (s2 * MUL[l] + nxt[l][0] + nxt[l][1]) & M
            if l == 0 and (g + 1) % 80 in (1, 2, 3):
                PD[l] = (PD[l] + fix(g + 1)) & M                    # rounds 1..3: in the peeled iterations 0..2
            D[l] = nxt[l][1]
            if l == 0 and (g + 1) % 80 in (1, 2, 3):
                D[l] = (D[l] + fix(g + 1)) & M
            if not (l == 0 and g >= total):                         # lane 0 is done after its last block
                win[l] = [e, s0, s1, s2]
        T[g], A[g] = out
    win[1] = [(x + y) & M for x, y in zip(win[1], H[1])]
    return b"".join(x.to_bytes(8, "big") for x in win[1] + H[0])


def test_continuous_pair_schedule_matches_hashlib():
    rng = np.random.default_rng(3)
    for n in (0, 1, 111, 112, 128, 129, 1000, 4096, 5001, 20000):
        m = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert sha512_pair_continuous(m) == hashlib.sha512(m).digest(), n
