// sha512_core.cuh -- SHA-512 compression function for sm_100a, one message per lane.
//
// Replaces the block function of Go's crypto/sha512 that helpers.Sha512sum drives
// (reference: helpers/helpers.go:188-201).  FIPS 180-4 section 6.4.2.
//
// Lowering (checked with cuobjdump -sass, see DESIGN.md "SHA-512 kernel"):
//   * every 64-bit rotate is two funnel shifts (SHF.R.W.U32), shr is SHF + SHF.R.U32.HI;
//   * xor-of-three, Ch and Maj are one LOP3.LUT per 32-bit half;
//   * 64-bit adds selected by kRoundFma / kSchedFma run on the FMA pipe instead of the ALU
//     pipe: a + b = IMAD.WIDE.U32(a.lo, one, b) followed by IMAD(a.hi, one, acc.hi), where
//     `one` is a run-time 1 the compiler cannot fold.  On B200 the ALU pipe (IADD3 / LOP3 /
//     SHF / PRMT) issues one warp instruction per 2 clocks per SM sub-partition and is the
//     bound of this kernel; moving the adds to the otherwise idle FMA pipe takes the ALU
//     count per block from ~3420 to ~2670;
//   * the 80 rounds are fully unrolled, so the round constants are immediates and the
//     16-word rolling schedule lives in registers.
#pragma once
#include <cstdint>

namespace snapgpu {

typedef uint64_t u64;
typedef uint32_t u32;

__device__ __forceinline__ void unpack64(u64 x, u32 &lo, u32 &hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x));
}
__device__ __forceinline__ u64 pack64(u32 lo, u32 hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}

// 64-bit add on the FMA pipe.  ptxas fuses mul.wide.u32 + add.u64 into one
// IMAD.WIDE.U32 Rd, Ra, Rb, Rc(pair); the high halves are summed by a plain IMAD.
__device__ __forceinline__ u64 add64_fma(u64 a, u64 b, u32 one) {
    u32 alo, ahi;
    unpack64(a, alo, ahi);
    u64 m;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(m) : "r"(alo), "r"(one));
    u64 acc = m + b;
    u32 lo, hi;
    unpack64(acc, lo, hi);
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(ahi), "r"(one), "r"(hi));
    return pack64(lo, hi);
}

template <bool kFma>
__device__ __forceinline__ u64 add64(u64 a, u64 b, u32 one) {
    if constexpr (kFma) return add64_fma(a, b, one);
    else return a + b;
}

__device__ __forceinline__ u32 lop3_xor3(u32 a, u32 b, u32 c) {
    u32 r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ u32 lop3_ch(u32 e, u32 f, u32 g) {   // (e & f) ^ (~e & g)
    u32 r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(e), "r"(f), "r"(g));
    return r;
}
__device__ __forceinline__ u32 lop3_maj(u32 a, u32 b, u32 c) {  // majority
    u32 r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// One 32-bit half pair of rotr64(x, N): N < 32 takes (lo,hi), N >= 32 the swapped pair.
template <int N>
__device__ __forceinline__ void rotr64_halves(u32 lo, u32 hi, u32 &rlo, u32 &rhi) {
    if constexpr (N < 32) {
        rlo = __funnelshift_r(lo, hi, N);
        rhi = __funnelshift_r(hi, lo, N);
    } else {
        rlo = __funnelshift_r(hi, lo, N - 32);
        rhi = __funnelshift_r(lo, hi, N - 32);
    }
}

// rotr(x,A) ^ rotr(x,B) ^ (kShr ? x >> C : rotr(x,C))
template <int A, int B, int C, bool kShr>
__device__ __forceinline__ u64 sigma(u64 x) {
    u32 lo, hi, alo, ahi, blo, bhi, clo, chi;
    unpack64(x, lo, hi);
    rotr64_halves<A>(lo, hi, alo, ahi);
    rotr64_halves<B>(lo, hi, blo, bhi);
    if constexpr (kShr) {
        clo = __funnelshift_r(lo, hi, C);
        chi = hi >> C;
    } else {
        rotr64_halves<C>(lo, hi, clo, chi);
    }
    return pack64(lop3_xor3(alo, blo, clo), lop3_xor3(ahi, bhi, chi));
}

__device__ __forceinline__ u64 big_sigma0(u64 x) { return sigma<28, 34, 39, false>(x); }
__device__ __forceinline__ u64 big_sigma1(u64 x) { return sigma<14, 18, 41, false>(x); }
__device__ __forceinline__ u64 small_sigma0(u64 x) { return sigma<1, 8, 7, true>(x); }
__device__ __forceinline__ u64 small_sigma1(u64 x) { return sigma<19, 61, 6, true>(x); }

__device__ __forceinline__ u64 ch64(u64 e, u64 f, u64 g) {
    u32 el, eh, fl, fh, gl, gh;
    unpack64(e, el, eh);
    unpack64(f, fl, fh);
    unpack64(g, gl, gh);
    return pack64(lop3_ch(el, fl, gl), lop3_ch(eh, fh, gh));
}
__device__ __forceinline__ u64 maj64(u64 a, u64 b, u64 c) {
    u32 al, ah, bl, bh, cl, chh;
    unpack64(a, al, ah);
    unpack64(b, bl, bh);
    unpack64(c, cl, chh);
    return pack64(lop3_maj(al, bl, cl), lop3_maj(ah, bh, chh));
}

// FIPS 180-4 section 4.2.3
__device__ static constexpr u64 kK512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL,
    0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL,
    0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL,
    0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL,
    0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL,
    0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, 0x06ca6351e003826fULL, 0x142929670a0e6e70ULL,
    0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL,
    0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL,
    0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL,
    0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL,
    0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL,
    0xca273eceea26619cULL, 0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL,
    0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL,
    0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL,
};

// FIPS 180-4 section 5.3.5
__device__ static constexpr u64 kIV512[8] = {
    0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
    0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL,
};

// Bits of kRoundFma: which of the seven 64-bit adds of a round go to the FMA pipe
//   0: K+W   1: h+(K+W)   2: Sigma1+Ch   3: T1   4: T2=Sigma0+Maj   5: e'=d+T1   6: a'=T1+T2
// Bits of kSchedFma: the three adds of a schedule word
//   0: sigma0+W[t-16]   1: sigma1+W[t-7]   2: their sum
#define SNAPGPU_RBIT(i) (((kRoundFma) >> (i)) & 1)
#define SNAPGPU_SBIT(i) (((kSchedFma) >> (i)) & 1)

// One 128-byte block.  w[16] holds the big-endian message words on entry and is clobbered.
// `commit` (per lane) gates the feed-forward so that lanes past their last block keep
// their state while the warp stays converged.
template <int kRoundFma, int kSchedFma>
__device__ __forceinline__ void sha512_compress(u64 (&st)[8], u64 (&w)[16], bool commit, u32 one) {
    u64 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int t = 0; t < 80; t++) {
        if (t >= 16) {
            u64 x = add64<SNAPGPU_SBIT(0)>(small_sigma0(w[(t - 15) & 15]), w[t & 15], one);
            u64 y = add64<SNAPGPU_SBIT(1)>(small_sigma1(w[(t - 2) & 15]), w[(t - 7) & 15], one);
            w[t & 15] = add64<SNAPGPU_SBIT(2)>(x, y, one);
        }
        u64 kw = add64<SNAPGPU_RBIT(0)>(kK512[t], w[t & 15], one);
        u64 hk = add64<SNAPGPU_RBIT(1)>(h, kw, one);
        u64 sc = add64<SNAPGPU_RBIT(2)>(big_sigma1(e), ch64(e, f, g), one);
        u64 t1 = add64<SNAPGPU_RBIT(3)>(hk, sc, one);
        u64 t2 = add64<SNAPGPU_RBIT(4)>(big_sigma0(a), maj64(a, b, c), one);
        h = g;
        g = f;
        f = e;
        e = add64<SNAPGPU_RBIT(5)>(d, t1, one);
        d = c;
        c = b;
        b = a;
        a = add64<SNAPGPU_RBIT(6)>(t1, t2, one);
    }
    if (commit) {
        st[0] += a; st[1] += b; st[2] += c; st[3] += d;
        st[4] += e; st[5] += f; st[6] += g; st[7] += h;
    }
}


// ------------------------------------------------------------------------------------------
// Compact form: 4 x (16 rounds + 16 schedule words) in a loop, then 16 rounds.  ~20 KB of code
// instead of ~65 KB, so the loop stays in the SM's instruction cache (ncu: "no instruction"
// was the top stall reason of the fully unrolled form).  Round constants come from constant
// memory through the uniform datapath.
// ------------------------------------------------------------------------------------------

__constant__ u64 c_K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL,
    0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL,
    0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL,
    0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL,
    0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL,
    0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, 0x06ca6351e003826fULL, 0x142929670a0e6e70ULL,
    0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL,
    0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL,
    0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL,
    0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL,
    0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL,
    0xca273eceea26619cULL, 0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL,
    0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL,
    0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL,
};

// 64-bit add with the low half on the ALU pipe (IADD3, carry out) and the high half on the
// FMA pipe (IMAD.X: a.hi * one + b.hi + carry).
__device__ __forceinline__ u64 add64_split(u64 a, u64 b, u32 one) {
    u32 al, ah, bl, bh, lo, hi;
    unpack64(a, al, ah);
    unpack64(b, bl, bh);
    asm("{add.cc.u32 %0, %2, %3;\n\tmadc.lo.u32 %1, %4, %6, %5;}"
        : "=r"(lo), "=r"(hi) : "r"(al), "r"(bl), "r"(ah), "r"(bh), "r"(one));
    return pack64(lo, hi);
}

// kAddMode 0: plain 64-bit adds (ptxas pairs them into 3-input IADD3 / IADD3.X)
// kAddMode 1: every add split ALU(lo) / FMA(hi)
// kAddMode 0x1000 | sched << 8 | round: per-add pipe choice.  A set bit sends that 64-bit add
//   to the FMA pipe as IMAD.WIDE.U32 + IMAD (add64_fma); a clear bit leaves it on the ALU.
//   round bits  0: W+K   1: h+(W+K)   2: Sigma1+Ch   3: T1   4: T2=Sigma0+Maj   5: e'=d+T1   6: a'=T1+T2
//   sched bits  0: sigma0+W[t-16]   1: sigma1+W[t-7]   2: their sum
template <int kAddMode, int kBit>
__device__ __forceinline__ u64 addm(u64 a, u64 b, u32 one) {
    if constexpr (kAddMode == 1) return add64_split(a, b, one);
    else if constexpr ((kAddMode & 0x1000) != 0 && ((kAddMode >> kBit) & 1) != 0) return add64_fma(a, b, one);
    else return a + b;
}

template <int kAddMode>
__device__ __forceinline__ void sha512_round(u64 a, u64 b, u64 c, u64 &d, u64 e, u64 f, u64 g, u64 &h, u64 w, u64 k, u32 one) {
    u64 kw = addm<kAddMode, 0>(w, k, one);
    u64 t1 = addm<kAddMode, 3>(addm<kAddMode, 1>(h, kw, one), addm<kAddMode, 2>(big_sigma1(e), ch64(e, f, g), one), one);
    u64 t2 = addm<kAddMode, 4>(big_sigma0(a), maj64(a, b, c), one);
    d = addm<kAddMode, 5>(d, t1, one);
    h = addm<kAddMode, 6>(t1, t2, one);
}

#define SNAPGPU_R8(W, KB, I)                                                            \
    sha512_round<kAddMode>(a, b, c, d, e, f, g, h, W[I + 0], KB[I + 0], one);           \
    sha512_round<kAddMode>(h, a, b, c, d, e, f, g, W[I + 1], KB[I + 1], one);           \
    sha512_round<kAddMode>(g, h, a, b, c, d, e, f, W[I + 2], KB[I + 2], one);           \
    sha512_round<kAddMode>(f, g, h, a, b, c, d, e, W[I + 3], KB[I + 3], one);           \
    sha512_round<kAddMode>(e, f, g, h, a, b, c, d, W[I + 4], KB[I + 4], one);           \
    sha512_round<kAddMode>(d, e, f, g, h, a, b, c, W[I + 5], KB[I + 5], one);           \
    sha512_round<kAddMode>(c, d, e, f, g, h, a, b, W[I + 6], KB[I + 6], one);           \
    sha512_round<kAddMode>(b, c, d, e, f, g, h, a, W[I + 7], KB[I + 7], one);

template <int kAddMode>
__device__ __forceinline__ void sha512_compress_compact(u64 (&st)[8], u64 (&w)[16], bool commit, u32 one) {
    u64 a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int grp = 0; grp < 4; grp++) {
        const u64 *kb = c_K512 + 16 * grp;
        SNAPGPU_R8(w, kb, 0)
        SNAPGPU_R8(w, kb, 8)
        // schedule for the next 16 rounds, in place: W[t+16] = s1(W[t+14]) + W[t+9] + s0(W[t+1]) + W[t]
#pragma unroll
        for (int i = 0; i < 16; i++) {
            u64 x = addm<kAddMode, 8>(small_sigma0(w[(i + 1) & 15]), w[i], one);
            u64 y = addm<kAddMode, 9>(small_sigma1(w[(i + 14) & 15]), w[(i + 9) & 15], one);
            w[i] = addm<kAddMode, 10>(x, y, one);
        }
    }
    {
        const u64 *kb = c_K512 + 64;
        SNAPGPU_R8(w, kb, 0)
        SNAPGPU_R8(w, kb, 8)
    }
    if (commit) {
        st[0] += a; st[1] += b; st[2] += c; st[3] += d;
        st[4] += e; st[5] += f; st[6] += g; st[7] += h;
    }
}

__device__ __forceinline__ u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }

// Two little-endian 32-bit loads (memory order lo, hi) -> the big-endian 64-bit word.
__device__ __forceinline__ u64 be64_from_le_words(u32 first, u32 second) {
    return pack64(bswap32(second), bswap32(first));
}

}  // namespace snapgpu
