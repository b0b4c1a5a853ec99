// sha512_pair.cuh -- the long-file bin, second form: one SHA-512 chain on TWO lanes.
//
// sha512_long_kernel (sha512_long.cuh) takes the message schedule off the serial chain; what is
// left on the consumer's lane is the 80 rounds, ~2250 ALU instructions per block, and a lone
// warp issues one ALU instruction per 2 clocks however many of its lanes are active.  This
// kernel halves the round itself across a lane pair, in SIMT: the SHA-2 round is two coupled
// recurrences,
//
//     e' = d + T1,  T1 = h + Sigma1(e) + Ch(e,f,g) + (W+K)        (lane 0 keeps e,f,g,h)
//     a' = T1 + Sigma0(a) + Maj(a,b,c)                            (lane 1 keeps a,b,c,d)
//
// and both have the same shape: new = Sigma(x0) + F(x0,x1,x2) + P (+ D).  One instruction
// stream serves both lanes with per-lane parameters:
//   * Sigma in factored form rotr(x ^ rotr(x,p) ^ rotr(x,q), s) with (p,q,s) = (4,27,14) for
//     Sigma1 and (6,11,28) for Sigma0 -- all below 32, so the same six funnel shifts with the
//     amounts in registers;
//   * Maj(a,b,c) = Ch(~(a^b), b, c): t = LOP3(x0, x1, lane-mask) picks x0 or ~(x0^x1), then one
//     Ch LOP3;
//   * P = h + (W+K) on lane 0 and T1 on lane 1; D = d on lane 0 and 0 on lane 1.
// 18 ALU instructions per round for the pair instead of 28 for one lane.  The lanes talk through
// shared-memory mailboxes read by the very loads that fetch W+K (lane 0's "next W+K" load is
// lane 1's "T1 of two rounds ago" load; lane 0's "d" load reads what lane 1 stored two iterations
// earlier), so no select or shuffle sits in the loop.  Lane 1 runs two rounds behind lane 0,
// which puts a whole iteration between every store and the load that needs it; a block is 82
// iterations, and lane 1's two idle ones at the start are used to shift its window in: its
// mailbox is seeded so that they reproduce b and a (the round is solved for T1).
//
// Producer warp, ring protocol and descriptors are those of sha512_long_kernel; the ring is
// laid out [step][file][81 words] so that a round's W+K is at base + 8*t whatever the number of
// files.
#pragma once
#include "sha512_long.cuh"

namespace snapgpu {

constexpr int kPairFilesPerCta = 16;                 // one consumer lane PAIR per file
constexpr int kPairMailWords = 89;                   // T[-2..86] / A[-2..86] of one file.  Odd, so that a file's T and A
                                                     // and the T's of neighbouring files (stride 178 = 2 mod 16 words)
                                                     // fall into different banks for the 16 lanes of a half-warp
constexpr int kPairRingPad = 8;                      // the last slot's W+K prefetch runs 2 words over
constexpr size_t kPairRingBytes = (size_t)(kLongRingWords + kPairRingPad) * 8;
constexpr size_t kPairMailBytes = (size_t)kPairFilesPerCta * 2 * kPairMailWords * 8;
constexpr size_t kPairZeroBytes = (size_t)kPairMailWords * 8;
constexpr size_t kPairUsedBytes = kPairRingBytes + kPairMailBytes + 2 * kPairZeroBytes + 16;
// A CTA of this kernel keeps its SM to itself: a batched-kernel CTA on the same SM would put a second warp on the
// consumer's sub-partition and halve the chain's share of the ALU pipe (measured: 1.40 -> 1.58 ms for a 10 000-file
// batch once the two were allowed to share).  The request is rounded up so that what is left of the SM's 228 KB
// cannot hold the smallest batched-kernel CTA (33 KB + 1 KB reserved).
constexpr size_t kPairSmemBytes = kPairUsedBytes > (200u << 10) ? kPairUsedBytes : (200u << 10);
static_assert(kPairSmemBytes + 1024 <= 227u * 1024, "dynamic + static shared memory of one CTA must fit an SM");

// rotr64 by a per-lane amount below 32
__device__ __forceinline__ u64 rotr64_var(u64 x, u32 r) {
    u32 lo, hi;
    unpack64(x, lo, hi);
    return pack64(__funnelshift_r(lo, hi, r), __funnelshift_r(hi, lo, r));
}
// Sigma0 / Sigma1 by lane: rotr(x ^ rotr(x,p) ^ rotr(x,q), s)
__device__ __forceinline__ u64 pair_sigma(u64 x, u32 p, u32 q, u32 s) {
    u32 lo, hi, alo, ahi, blo, bhi;
    unpack64(x, lo, hi);
    alo = __funnelshift_r(lo, hi, p);
    ahi = __funnelshift_r(hi, lo, p);
    blo = __funnelshift_r(lo, hi, q);
    bhi = __funnelshift_r(hi, lo, q);
    const u32 ylo = lop3_xor3(lo, alo, blo), yhi = lop3_xor3(hi, ahi, bhi);
    return pack64(__funnelshift_r(ylo, yhi, s), __funnelshift_r(yhi, ylo, s));
}
// lane 0 (mask 0): Ch(x0,x1,x2); lane 1 (mask ~0): Maj(x0,x1,x2) = Ch(~(x0^x1), x1, x2)
__device__ __forceinline__ u32 pair_f32(u32 x0, u32 x1, u32 x2, u32 mask) {
    u32 t;
    asm("lop3.b32 %0, %1, %2, %3, 0xD2;" : "=r"(t) : "r"(x0), "r"(x1), "r"(mask));   // m ? ~(a^b) : a
    return lop3_ch(t, x1, x2);
}
__device__ __forceinline__ u64 pair_f(u64 x0, u64 x1, u64 x2, u32 mask) {
    u32 a0, a1, b0, b1, c0, c1;
    unpack64(x0, a0, a1);
    unpack64(x1, b0, b1);
    unpack64(x2, c0, c1);
    return pack64(pair_f32(a0, b0, c0, mask), pair_f32(a1, b1, c1, mask));
}
// x * mul + y with mul in {0,1}: lane 0 adds its h to W+K, lane 1 takes T1 as it is.  The two
// multiplies run on the FMA pipe.
__device__ __forceinline__ u64 pair_muladd(u64 x, u32 mul, u64 y) {
    u32 lo, hi;
    unpack64(x, lo, hi);
    return pack64(lo * mul, hi * mul) + y;
}

__device__ __forceinline__ u64 pair_exchange(u64 x) {      // with the other lane of the pair
    return __shfl_xor_sync(0xffffffffu, x, 1);
}

// How the two lanes of a pair exchange their round results -- each needs exactly what its partner
// produced in the previous iteration:
//
// kShuffle = false (option pair_form 0, the default): through shared-memory mailboxes read by the
//   loads that fetch W+K (see the file header).  A store of iteration i is loaded by the partner
//   in iteration i+1, so the pair must execute the straight-line round code together.  That is
//   arranged, not assumed: the prologue and every region of rounds start with __syncwarp() -- a
//   convergence point; the slow path of its BRA.DIV re-converges a diverged warp -- and contain
//   no branch up to the next one, and converged lanes do not part in branch-free code: ptxas
//   relies on the same fact when it checks convergence ONCE for the 32 shfl.sync of a group of
//   the shuffle form below.  tests/test_sass.py pins the shape (no branch inside a region, every
//   mailbox store ahead of the load of the next round in program order).
//   A convergence point costs ~25 clocks of the chain, so regions are as long as the instruction
//   cache allows (profiles/r02_pair_regions.jsonl):
//     kRegions = 2: two regions of 41 rounds (15 KB loop body, fits the L0 instruction cache):
//                   1.90 us per block whatever the producer does;
//     kRegions = 1: prologue and all 82 rounds are one region (30 KB): 1.83 us per block while
//                   the producer warp is mostly idle (1 file per CTA; 1.88 with 2), but it streams
//                   from the shared L1 instruction cache and loses to a busy producer (2.6 us with
//                   12+ files per CTA).  The host therefore gives a chain its own CTA while a quarter
//                   of the SMs last, two per CTA up to twice that, and picks kRegions = 1 for those.
// kShuffle = true (pair_form 1): a warp shuffle (__shfl_xor_sync) carries the exchange, so the
//   synchronisation is in the instruction itself.  Role 1 loads a zero where role 0 loads W+K,
//   and both form PD = S2*mul + kw + r (tests/test_pair_schedule.py::compress_pair_shuffle).
//   2.04 us per block -- two SHFL and two more IMAD per round cost more than the LDS/STS they
//   replace (profiles/r02_pair_forms.jsonl) -- so it is the cross-check, not the default.
template <bool kAligned16, bool kShuffle, int kRegions>
__global__ void __launch_bounds__(kLongThreads, 1)
sha512_pair_kernel(const uint8_t *__restrict__ data, const SegDesc *__restrict__ descs, u32 nsegs,
                   uint8_t *__restrict__ digests, u32 files_per_cta) {
    extern __shared__ __align__(16) uint8_t pair_smem[];
    u64 *ring = reinterpret_cast<u64 *>(pair_smem);
    u64 *mail = reinterpret_cast<u64 *>(pair_smem + kPairRingBytes);
    u64 *zero = reinterpret_cast<u64 *>(pair_smem + kPairRingBytes + kPairMailBytes);
    u64 *junk = zero + kPairMailWords;
    u32 *produced = reinterpret_cast<u32 *>(pair_smem + kPairRingBytes + kPairMailBytes + 2 * kPairZeroBytes);
    u32 *consumed = produced + 1;

    const u32 lane = threadIdx.x & 31;
    const u32 warp = threadIdx.x >> 5;
    const u32 first = blockIdx.x * files_per_cta;         // files_per_cta <= kPairFilesPerCta, chosen by the host
    const u32 count = min(files_per_cta, nsegs - first);

    if (threadIdx.x == 0) {
        st_volatile_shared(produced, 0);
        st_volatile_shared(consumed, 0);
    }
    for (u32 i = threadIdx.x; i < (u32)kPairMailWords; i += kLongThreads) zero[i] = 0;
    __syncthreads();

    const u32 nf = count;                                  // files of this CTA (1..16)
    const u32 ring_steps = 256 / nf;                       // block steps the ring holds

    if (warp == 0) {
        // ---------------- consumer: lane pair (2f, 2f+1) runs the rounds of file f ----------------
        const u32 file = lane >> 1, role = lane & 1;       // role 0: e,f,g,h   role 1: a,b,c,d
        const bool have = file < count;
        SegDesc sd;
        sd.off = 0; sd.len = 0; sd.prefix = 0; sd.out_idx = 0; sd.flags = kSegNoFinal;
        if (have) sd = descs[first + file];
        const u32 my_blocks = have ? (u32)seg_blocks(sd.len, sd.flags) : 0u;
        const u32 steps = __reduce_max_sync(0xffffffffu, my_blocks);
        uint8_t *out_digest = digests + (size_t)sd.out_idx * 64 + (role ? 0 : 32);
        u64 st[4];                                         // role 1: H0..H3, role 0: H4..H7
        if (have && (sd.flags & kSegContinue)) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(out_digest);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint4 v = s4[i];
                st[2 * i] = be64_from_le_words(v.x, v.y);
                st[2 * i + 1] = be64_from_le_words(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) st[i] = kIV512[(role ? 0 : 4) + i];
        }
        // per-lane parameters of the shared instruction stream
        const u32 rp = role ? 6u : 4u, rq = role ? 11u : 27u, rs = role ? 28u : 14u;
        const u32 mask = role ? 0xffffffffu : 0u, mul = role ? 0u : 1u;
        volatile u64 *Tm = mail + (size_t)(2 * min(file, (u32)kPairFilesPerCta - 1)) * kPairMailWords;   // T[k] = Tm[k+2]
        volatile u64 *Am = Tm + kPairMailWords;                                                     // A[k] = Am[k+2]
        volatile u64 *const out = (role ? Am : Tm) + 2;    // iteration i stores T1_i / a_(i-1)
        volatile u64 *const din = role ? zero : Am;        // iteration i adds A[i-2] = a_(i-3) / 0
        volatile u64 *const seed_a = role ? Am : junk;     // prologue: A[-2], A[-1]
        volatile u64 *const seed_t = role ? Tm : junk + 4; // prologue: T[-2], T[-1]

        u32 ready = 0, slot_index = 0;
        for (u32 b = 0; b < steps; b++) {
            if (b >= ready) {
                do ready = ld_volatile_shared(produced); while (b >= ready);
                __threadfence_block();
            }
            volatile u64 *const kin = role ? (kShuffle ? zero : Tm)
                                           : ring + ((size_t)slot_index * nf + min(file, nf - 1)) * kLongSlotWords;
            slot_index = slot_index + 1 == ring_steps ? 0 : slot_index + 1;
            if (kShuffle) {
                // role 0 needs role 1's d and c (H3, H2) for its first two rounds; role 1 starts from
                // the two seeds that make its idle iterations produce b and a
                const u64 kw0 = kin[0];
                const u64 x3 = pair_exchange(st[3]), x2 = pair_exchange(st[2]);
                const u64 seed0 = st[1] - pair_sigma(st[2], rp, rq, rs) - pair_f(st[2], st[3], 0, mask);
                const u64 seed1 = st[0] - pair_sigma(st[1], rp, rq, rs) - pair_f(st[1], st[2], st[3], mask);
                u64 D = role ? 0 : x3;                     // what is subtracted from this round's result: d | 0
                u64 r = role ? seed1 : x2;                 // what the partner sent last: T1 | a
                u64 S0 = role ? st[2] : st[0], S1 = role ? st[3] : st[1], S2 = role ? 0 : st[2], S3 = role ? 0 : st[3];
                u64 PD = role ? seed0 : st[3] + kw0 + D;
#define SNAPGPU_PAIR_ITER_X(KIN, I)                                                           \
    {                                                                                            \
        const u64 kwn = (KIN)[(I) + 1];                                                          \
        const u64 e = pair_sigma(S0, rp, rq, rs) + pair_f(S0, S1, S2, mask) + PD;                \
        const u64 out = e - D;                                                                   \
        PD = pair_muladd(S2, mul, kwn) + r;                                                      \
        D = pair_muladd(r, mul, 0);                                                              \
        r = pair_exchange(out);                                                                  \
        S3 = S2; S2 = S1; S1 = S0; S0 = e;                                                       \
    }
#pragma unroll 1
                for (int grp = 0; grp < 5; grp++) {
                    volatile u64 *const k = kin + 16 * grp;
                    SNAPGPU_PAIR_ITER_X(k, 0)  SNAPGPU_PAIR_ITER_X(k, 1)  SNAPGPU_PAIR_ITER_X(k, 2)  SNAPGPU_PAIR_ITER_X(k, 3)
                    SNAPGPU_PAIR_ITER_X(k, 4)  SNAPGPU_PAIR_ITER_X(k, 5)  SNAPGPU_PAIR_ITER_X(k, 6)  SNAPGPU_PAIR_ITER_X(k, 7)
                    SNAPGPU_PAIR_ITER_X(k, 8)  SNAPGPU_PAIR_ITER_X(k, 9)  SNAPGPU_PAIR_ITER_X(k, 10) SNAPGPU_PAIR_ITER_X(k, 11)
                    SNAPGPU_PAIR_ITER_X(k, 12) SNAPGPU_PAIR_ITER_X(k, 13) SNAPGPU_PAIR_ITER_X(k, 14) SNAPGPU_PAIR_ITER_X(k, 15)
                }
                const u64 e0 = S0, e1 = S1, e2 = S2, e3 = S3;      // role 0 is done after iteration 79; role 1 needs two more
                {
                    volatile u64 *const k = kin + 80;
                    SNAPGPU_PAIR_ITER_X(k, 0)  SNAPGPU_PAIR_ITER_X(k, 1)
                }
#undef SNAPGPU_PAIR_ITER_X
                if (b < my_blocks) {
                    st[0] += role ? S0 : e0;
                    st[1] += role ? S1 : e1;
                    st[2] += role ? S2 : e2;
                    st[3] += role ? S3 : e3;
                }
                __syncwarp();                              // this step's slot may be overwritten once every lane has read it
                if (lane == 0) st_volatile_shared(consumed, b + 1);
                continue;
            }

            // prologue.  role 1 publishes d and c, and the seed that makes its second idle iteration
            // produce a (the first one's, which produces b, it keeps in a register); its window
            // starts as (c, d, 0).  Ordered so that every load sits well behind the store it needs.
            __syncwarp();                                  // convergence point (see the template's comment)
            const u64 kw0 = kin[0];
            seed_a[0] = st[3];
            seed_a[1] = st[2];
            const u64 seed0 = st[1] - pair_sigma(st[2], rp, rq, rs) - pair_f(st[2], st[3], 0, mask);
            u64 D = din[0];
            seed_t[1] = st[0] - pair_sigma(st[1], rp, rq, rs) - pair_f(st[1], st[2], st[3], mask);
            u64 S0 = role ? st[2] : st[0], S1 = role ? st[3] : st[1], S2 = role ? 0 : st[2], S3 = role ? 0 : st[3];
            u64 PD = role ? seed0 : st[3] + kw0 + D;       // T1 that yields b  |  h + (W+K) + d

#define SNAPGPU_PAIR_ITER(KIN, DIN, OUT, I)                                                   \
    {                                                                                            \
        const u64 kwn = (KIN)[(I) + 1];                                                          \
        const u64 dn = (DIN)[(I) + 1];                                                           \
        const u64 e = pair_sigma(S0, rp, rq, rs) + pair_f(S0, S1, S2, mask) + PD;                \
        (OUT)[(I)] = e - D;                                                                      \
        PD = pair_muladd(S2, mul, kwn) + dn;                                                     \
        D = dn;                                                                                  \
        S3 = S2; S2 = S1; S1 = S0; S0 = e;                                                       \
    }
            // 82 iterations in kRegions branch-free regions (1 or 2), each behind a convergence point; role 0
            // is done after iteration 79, role 1 needs two more
            u64 e0 = 0, e1 = 0, e2 = 0, e3 = 0;
#define SNAPGPU_PAIR_ITER4(k, d, o, I)                                                           \
    SNAPGPU_PAIR_ITER(k, d, o, I) SNAPGPU_PAIR_ITER(k, d, o, (I) + 1) SNAPGPU_PAIR_ITER(k, d, o, (I) + 2) SNAPGPU_PAIR_ITER(k, d, o, (I) + 3)
#define SNAPGPU_PAIR_ITER36(k, d, o, I)                                                          \
    SNAPGPU_PAIR_ITER4(k, d, o, I) SNAPGPU_PAIR_ITER4(k, d, o, (I) + 4) SNAPGPU_PAIR_ITER4(k, d, o, (I) + 8)                 \
    SNAPGPU_PAIR_ITER4(k, d, o, (I) + 12) SNAPGPU_PAIR_ITER4(k, d, o, (I) + 16) SNAPGPU_PAIR_ITER4(k, d, o, (I) + 20)        \
    SNAPGPU_PAIR_ITER4(k, d, o, (I) + 24) SNAPGPU_PAIR_ITER4(k, d, o, (I) + 28) SNAPGPU_PAIR_ITER4(k, d, o, (I) + 32)
            if (kRegions == 1) {
                // no convergence point of its own: nothing but straight-line code since the prologue's
                SNAPGPU_PAIR_ITER36(kin, din, out, 0) SNAPGPU_PAIR_ITER36(kin, din, out, 36)
                SNAPGPU_PAIR_ITER4(kin, din, out, 72) SNAPGPU_PAIR_ITER4(kin, din, out, 76)
                e0 = S0; e1 = S1; e2 = S2; e3 = S3;
                SNAPGPU_PAIR_ITER(kin, din, out, 80) SNAPGPU_PAIR_ITER(kin, din, out, 81)
            } else {
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    volatile u64 *const k = kin + 41 * half, *const d = din + 41 * half, *const o = out + 41 * half;
                    __syncwarp();                          // convergence point: the 41 rounds below are branch-free
                    SNAPGPU_PAIR_ITER36(k, d, o, 0)
                    SNAPGPU_PAIR_ITER(k, d, o, 36) SNAPGPU_PAIR_ITER(k, d, o, 37) SNAPGPU_PAIR_ITER(k, d, o, 38)
                    e0 = S0; e1 = S1; e2 = S2; e3 = S3;    // the second half's is iteration 79
                    SNAPGPU_PAIR_ITER(k, d, o, 39) SNAPGPU_PAIR_ITER(k, d, o, 40)
                }
            }
#undef SNAPGPU_PAIR_ITER36
#undef SNAPGPU_PAIR_ITER4
#undef SNAPGPU_PAIR_ITER
            if (b < my_blocks) {
                st[0] += role ? S0 : e0;
                st[1] += role ? S1 : e1;
                st[2] += role ? S2 : e2;
                st[3] += role ? S3 : e3;
            }
            // this step's slot may be overwritten once every lane has read it
            __syncwarp();
            if (lane == 0) st_volatile_shared(consumed, b + 1);
        }
        if (have) {
            uint4 *o4 = reinterpret_cast<uint4 *>(out_digest);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                u32 a_lo, a_hi, b_lo, b_hi;
                unpack64(st[2 * i], a_lo, a_hi);
                unpack64(st[2 * i + 1], b_lo, b_hi);
                o4[i] = make_uint4(bswap32(a_hi), bswap32(a_lo), bswap32(b_hi), bswap32(b_lo));
            }
        }
    } else {
        // ---------------- producer: W[t] + K[t] of 32 (file, block) pairs per round ----------------
        // every lane needs the step count: the longest file of the CTA
        u32 blocks_of = 0;
        if (lane < count) {
            const SegDesc *d = descs + first + lane;
            blocks_of = (u32)seg_blocks(d->len, d->flags);
        }
        const u32 steps = __reduce_max_sync(0xffffffffu, blocks_of);
        // lane -> (file = lane % nf, step offset = lane / nf); a round covers `per` consecutive steps
        const u32 per = 32 / nf;                           // steps per round (>= 2); ring_steps >= 8 * per
        const u32 pf = lane % nf, ps = lane / nf;
        const bool worker = ps < per;                      // 32 % nf leftover lanes idle
        const SegDesc sd = descs[first + pf];
        const u32 f_blocks = (u32)seg_blocks(sd.len, sd.flags);
        const bool final_seg = !(sd.flags & kSegNoFinal);
        const u64 total_len = sd.prefix + sd.len;
        u32 done = 0;                                      // steps published so far
        while (done < steps) {
            const u32 upto = min(steps, done + per);
            u32 freed;
            while (freed = ld_volatile_shared(consumed), upto > freed + ring_steps) __nanosleep(200);
            __threadfence_block();
            const u32 b = done + ps;
            if (worker && b < upto) {
                u64 w[16];
                if (b < f_blocks) {
                    const long long rem = (long long)sd.len - (long long)b * 128;
                    u32 raw[32];
                    load_block<kAligned16>(data + sd.off + (size_t)b * 128, rem, raw);
#pragma unroll
                    for (int j = 0; j < 16; j++) w[j] = be64_from_le_words(raw[2 * j], raw[2 * j + 1]);
                    if (rem < 128) pad_block(w, rem, final_seg && (b + 1 == f_blocks), total_len);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) w[j] = 0;       // past this file's end: result is discarded
                }
                u64 *slot = ring + ((size_t)(b % ring_steps) * nf + pf) * kLongSlotWords;
#pragma unroll 1
                for (int grp = 0; grp < 5; grp++) {
#pragma unroll
                    for (int i = 0; i < 16; i++) slot[grp * 16 + i] = w[i] + c_K512[grp * 16 + i];
                    if (grp < 4) {
#pragma unroll
                        for (int i = 0; i < 16; i++)
                            w[i] = small_sigma1(w[(i + 14) & 15]) + w[(i + 9) & 15] + small_sigma0(w[(i + 1) & 15]) + w[i];
                    }
                }
            }
            __threadfence_block();
            __syncwarp();
            done = upto;
            if (lane == 0) st_volatile_shared(produced, done);
        }
    }
}

}  // namespace snapgpu
