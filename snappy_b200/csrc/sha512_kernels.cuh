// sha512_kernels.cuh -- batched multi-message SHA-512 over a packed buffer of files.
//
// Device side of snapgpu_sha512_batch[_device]; replaces the per-file loop
//   snappy/build.go:240-247  ->  helpers.Sha512sum (helpers/helpers.go:188-201).
//
// Work layout
//   * the host sorts the indices of the segment descriptors by 128-byte block count, longest
//     first (length binning: the 32 lanes of a warp get neighbours of the sorted list, so
//     they run the same number of blocks and stay converged);
//   * a "unit" is 32 consecutive entries of that order = one warp's worth of files;
//   * the kernel is persistent: one CTA of 4 warps per (SM x kCtasPerSm), every warp pulls
//     the next unit from a global counter (longest-processing-time-first list scheduling,
//     which is what bounds the makespan when a few files are much longer than the rest);
//   * each lane walks its own file: 8 x 128-bit loads per block straight into registers,
//     issued one block ahead of the compression that consumes them.
//
// A descriptor is a *segment* of a message so that a file larger than the staging buffer
// can be hashed in pieces: kSegContinue takes the chaining value from the output slot,
// kSegNoFinal stores the chaining value instead of padding and finishing.
#pragma once
#include "sha512_core.cuh"

namespace snapgpu {

struct SegDesc {
    u64 off;      // byte offset of the segment in the packed buffer
    u64 len;      // bytes in this segment (multiple of 128 unless it is the final one)
    u64 prefix;   // message bytes that came before this segment
    u32 out_idx;  // digest slot
    u32 flags;
};
enum : u32 { kSegContinue = 1u, kSegNoFinal = 2u, kSegSkip = 4u /* hashed by the long-file kernel instead */ };

__host__ __device__ inline u64 seg_blocks(u64 len, u32 flags) {
    if (flags & kSegSkip) return 0;
    return (flags & kSegNoFinal) ? (len >> 7) : ((len + 144) >> 7);   // (L+144)/128
}

constexpr int kShaWarpsPerCta = 4;
constexpr int kShaThreads = kShaWarpsPerCta * 32;

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ldg_nc_u32(const void *p) {
    u32 v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Raw 128 bytes of one block as 32 little-endian words, loads predicated on the bytes that
// exist (rem = bytes of the segment left at this block; <= 0 means none).
template <bool kAligned16>
__device__ __forceinline__ void load_block(const uint8_t *p, long long rem, u32 (&raw)[32]) {
    if (kAligned16) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if ((long long)(16 * i) < rem) v = ldg_nc_v4(p + 16 * i);
            raw[4 * i + 0] = v.x; raw[4 * i + 1] = v.y; raw[4 * i + 2] = v.z; raw[4 * i + 3] = v.w;
        }
    } else {
        // any alignment: 33 aligned words, realigned with a funnel shift by the byte phase
        const uintptr_t addr = (uintptr_t)p;
        const int phase = (int)(addr & 3);
        const uint8_t *base = p - phase;
        u32 prev = 0;
        if ((long long)(-phase) < rem) prev = ldg_nc_u32(base);
#pragma unroll
        for (int k = 0; k < 32; k++) {
            u32 next = 0;
            if ((long long)(4 * (k + 1) - phase) < rem) next = ldg_nc_u32(base + 4 * (k + 1));
            raw[k] = __funnelshift_r(prev, next, 8 * phase);
            prev = next;
        }
    }
}

// Turn the raw words of a block into big-endian message words and apply the FIPS 180-4
// 5.1.2 padding when the block holds fewer than 128 message bytes.
__device__ __forceinline__ void pad_block(u64 (&w)[16], long long rem, bool last_block, u64 total_len) {
    if (rem >= 128) return;
    const int r = rem > 0 ? (int)rem : 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        int nb = r - 8 * j;                       // message bytes in word j
        nb = nb < 0 ? 0 : (nb > 8 ? 8 : nb);
        u64 mask = nb == 0 ? 0ULL : (~0ULL << (64 - 8 * nb));
        u64 v = w[j] & mask;
        if (rem >= 0 && (r >> 3) == j) v |= 0x80ULL << (56 - 8 * (r & 7));
        w[j] = v;
    }
    if (last_block) {
        w[14] = total_len >> 61;
        w[15] = total_len << 3;
    }
}

template <int kRoundFma, int kSchedFma, bool kAligned16, int kCtasPerSm>
__global__ void __launch_bounds__(kShaThreads, kCtasPerSm)
sha512_segments_kernel(const uint8_t *__restrict__ data, const SegDesc *__restrict__ descs,
                       const u32 *__restrict__ order, u32 nsegs,
                       uint8_t *__restrict__ digests, u32 *__restrict__ unit_counter, u32 one, u32 /*first_wave*/) {
    const u32 lane = threadIdx.x & 31;
    const u32 warp = threadIdx.x >> 5;
    const u32 nunits = (nsegs + 31) >> 5;
    const u32 total_warps = gridDim.x * kShaWarpsPerCta;
    // first unit: consecutive (long) units land on different CTAs, hence different SMs
    u32 unit = warp * gridDim.x + blockIdx.x;

    while (unit < nunits) {
        const u32 idx = unit * 32 + lane;
        bool have = idx < nsegs;
        SegDesc sd;
        sd.off = 0; sd.len = 0; sd.prefix = 0; sd.out_idx = 0; sd.flags = kSegNoFinal;
        if (have) {
            const uint4 *q = reinterpret_cast<const uint4 *>(descs + order[idx]);
            uint4 q0 = q[0], q1 = q[1];
            sd.off = pack64(q0.x, q0.y); sd.len = pack64(q0.z, q0.w);
            sd.prefix = pack64(q1.x, q1.y); sd.out_idx = q1.z; sd.flags = q1.w;
            if (sd.flags & kSegSkip) have = false;
        }
        const u32 nblk = have ? (u32)seg_blocks(sd.len, sd.flags) : 0u;
        const u32 nblk_max = __reduce_max_sync(0xffffffffu, nblk);
        const bool final_seg = !(sd.flags & kSegNoFinal);
        const u64 total_len = sd.prefix + sd.len;
        uint8_t *out = digests + (size_t)sd.out_idx * 64;

        u64 st[8];
        if (have && (sd.flags & kSegContinue)) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(out);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 v = s4[i];
                st[2 * i] = be64_from_le_words(v.x, v.y);
                st[2 * i + 1] = be64_from_le_words(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = kIV512[i];
        }

        const uint8_t *p = data + sd.off;
        long long rem = (long long)sd.len;
        u32 raw[32];
        load_block<kAligned16>(p, rem, raw);

        for (u32 blk = 0; blk < nblk_max; blk++) {
            u64 w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) w[j] = be64_from_le_words(raw[2 * j], raw[2 * j + 1]);
            const bool active = blk < nblk;
            const long long rem_now = rem;
            // next block's bytes are requested before this block's 80 rounds
            p += 128;
            rem -= 128;
            load_block<kAligned16>(p, rem, raw);
            if (__any_sync(0xffffffffu, active && rem_now < 128))
                pad_block(w, rem_now, final_seg && (blk + 1 == nblk), total_len);
            sha512_compress<kRoundFma, kSchedFma>(st, w, active, one);
        }

        if (have) {
            uint4 *o4 = reinterpret_cast<uint4 *>(out);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                u32 a_lo, a_hi, b_lo, b_hi;
                unpack64(st[2 * i], a_lo, a_hi);
                unpack64(st[2 * i + 1], b_lo, b_hi);
                o4[i] = make_uint4(bswap32(a_hi), bswap32(a_lo), bswap32(b_hi), bswap32(b_lo));
            }
        }

        u32 next = 0;
        if (lane == 0) next = atomicAdd(unit_counter, 1u) + total_warps;
        unit = __shfl_sync(0xffffffffu, next, 0);
    }
}


// ------------------------------------------------------------------------------------------
// v2: message bytes staged through shared memory with cp.async (LDGSTS), two stages per warp.
//
// Each lane copies the 128 bytes of its own next block as 8 x 16 B asynchronous copies, one
// block ahead of the compression that consumes them; the copies of a tail block carry a
// src-size so that bytes past the end of the file are zero-filled, never read.  The stage is
// laid out [chunk][lane] (16 B units), so both the LDGSTS writes and the LDS.128 reads of a
// warp touch 32 consecutive 16-byte words: no bank conflicts.  A lane only ever reads what it
// copied itself, so cp.async.wait_group is all the synchronisation there is.
// Compared with loading into registers this keeps 32 registers free and leaves ptxas no way
// to pull a load's first use forward (ncu showed exactly that: a 16 % long-scoreboard stall).
// ------------------------------------------------------------------------------------------

constexpr int kStageBytesPerWarp = 8 * 32 * 16;   // 4 KiB: one 128-byte block per lane

__device__ __forceinline__ void cp_async_16(u32 smem_addr, const void *gptr, u32 src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cp_async_16_full(u32 smem_addr, const void *gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}

// queue the copies of the block at p (rem = bytes of the segment left there)
__device__ __forceinline__ void stage_block(u32 stage_addr, u32 lane, const uint8_t *p, long long rem,
                                            const uint8_t *safe) {
    const u32 dst = stage_addr + lane * 16u;
    if (rem >= 128) {                       // whole block: no per-chunk arithmetic
#pragma unroll
        for (int j = 0; j < 8; j++) cp_async_16_full(dst + (u32)j * 512u, p + 16 * j);
    } else {                                // tail (or nothing): zero-fill past the end
#pragma unroll
        for (int j = 0; j < 8; j++) {
            long long left = rem - 16 * j;
            u32 nbytes = left <= 0 ? 0u : (left >= 16 ? 16u : (u32)left);
            const uint8_t *src = nbytes ? p + 16 * j : safe;
            cp_async_16(dst + (u32)j * 512u, src, nbytes);
        }
    }
    cp_async_commit();
}

// Any byte alignment: the copies start at the 16-byte boundary at or below the block (p16), so the
// staged line holds `phase` bytes of whatever precedes the file followed by the block, 144 bytes
// in nine chunks.  end = phase + (bytes of the file in this block); nothing past it is read.
__device__ __forceinline__ void stage_block_any(u32 stage_addr, u32 lane, const uint8_t *p16, long long end,
                                                const uint8_t *safe) {
    const u32 dst = stage_addr + lane * 16u;
#pragma unroll
    for (int j = 0; j < 9; j++) {
        long long left = end - 16 * j;
        u32 nbytes = left <= 0 ? 0u : (left >= 16 ? 16u : (u32)left);
        const uint8_t *src = nbytes ? p16 + 16 * j : safe;
        cp_async_16(dst + (u32)j * 512u, src, nbytes);
    }
    cp_async_commit();
}

constexpr int kStageBytesPerWarpAny = 9 * 32 * 16;   // 4.5 KiB: 144 bytes per lane

#ifdef SNAPGPU_TRACE_WARPS
// experimental build only (tools/warp_trace_probe.py): when every warp of a launch started and
// finished, and how much it did
__device__ unsigned long long g_warp_trace[8192][4];
#endif

// kAligned16 = false: files may start at any byte.  The staged line is read back as 33 words from
// the file's own phase (four per-lane address deltas, one per word position mod 4, so a load is
// still "register + immediate"), and ONE byte-permute per word both realigns by the sub-word
// phase and swaps to big-endian -- the same 32 PRMTs the aligned form spends on the swap alone.
template <int kAddMode, int kCtasPerSm, bool kAligned16 = true>
__global__ void __launch_bounds__(kShaThreads, kCtasPerSm)
sha512_segments_kernel_v2(const uint8_t *__restrict__ data, const SegDesc *__restrict__ descs,
                          const u32 *__restrict__ order, u32 nsegs,
                          uint8_t *__restrict__ digests, u32 *__restrict__ unit_counter, u32 one, u32 first_wave) {
    constexpr int kStage = kAligned16 ? kStageBytesPerWarp : kStageBytesPerWarpAny;
    __shared__ __align__(16) uint8_t stages[kShaWarpsPerCta][2][kStage];
    const u32 lane = threadIdx.x & 31;
    const u32 warp = threadIdx.x >> 5;
    const u32 nunits = (nsegs + 31) >> 5;
    const u32 stage0 = (u32)__cvta_generic_to_shared(&stages[warp][0][0]);
    // The first unit of every warp is assigned statically, the others are claimed from a counter in
    // the order of the length-sorted plan (longest first): list scheduling.  (A variant that also
    // balanced the claims between SM sub-partitions with bounded spin-waits was measured in round 1
    // -- better on some mixed batches, worse on config 2, different from process to process -- and
    // has been removed: no wait loop lives in this kernel.)
    //
    // Two-ended claims (first_wave != 0, launches with several CTAs per SM).  An SM sub-partition
    // does not share its ALU pipe evenly: of two resident warps one gets ~83 % and the other ~16 %
    // (DESIGN.md section 10), so a long unit claimed by the slow one is a backlog that ends the
    // launch late.  Here a warp that is running slowly -- it times its own units: more than
    // kSlowClocksPerBlock per block -- claims its next unit from the SHORT end of the sorted plan,
    // everybody else from the long end; the two ends meet in the middle.  One 64-bit counter holds
    // both ends (low word: claims from the long end, high word: from the short end), so the order
    // of the atomic adds numbers the claims and every unit is claimed exactly once.  Nobody waits.
    const u32 total_warps = gridDim.x * kShaWarpsPerCta;   // units handed out before the counter starts
    // Which warp of a sub-partition is the favoured one is not known beforehand, so EVERY warp
    // starts with one of the shortest units -- a few blocks, enough to time itself -- and chooses
    // its end from then on.
    const bool two_ended = first_wave != 0;
    const u32 head0 = 0, tail0 = total_warps;
    u32 unit = warp * gridDim.x + blockIdx.x;
    if (two_ended) unit = nunits - 1 - unit;
    constexpr long long kSlowClocksPerBlock = 16000;       // a lone warp needs ~7600, the favoured one of two ~8400, the other ~43000
#ifdef SNAPGPU_TRACE_WARPS
    unsigned long long tr_start, tr_units = 0, tr_blocks = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_start));
#endif

    while (unit < nunits) {
        const long long t_unit = two_ended ? clock64() : 0;
        const u32 idx = unit * 32 + lane;
        bool have = idx < nsegs;
        SegDesc sd;
        sd.off = 0; sd.len = 0; sd.prefix = 0; sd.out_idx = 0; sd.flags = kSegNoFinal;
        if (have) {
            const uint4 *q = reinterpret_cast<const uint4 *>(descs + order[idx]);
            uint4 q0 = q[0], q1 = q[1];
            sd.off = pack64(q0.x, q0.y); sd.len = pack64(q0.z, q0.w);
            sd.prefix = pack64(q1.x, q1.y); sd.out_idx = q1.z; sd.flags = q1.w;
            if (sd.flags & kSegSkip) have = false;
        }
        const u32 nblk = have ? (u32)seg_blocks(sd.len, sd.flags) : 0u;
        const u32 nblk_max = __reduce_max_sync(0xffffffffu, nblk);
#ifdef SNAPGPU_TRACE_WARPS
        tr_units++;
        tr_blocks += nblk_max;
#endif
        const bool final_seg = !(sd.flags & kSegNoFinal);
        const u64 total_len = sd.prefix + sd.len;
        uint8_t *out = digests + (size_t)sd.out_idx * 64;

        u64 st[8];
        if (have && (sd.flags & kSegContinue)) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(out);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 v = s4[i];
                st[2 * i] = be64_from_le_words(v.x, v.y);
                st[2 * i + 1] = be64_from_le_words(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = kIV512[i];
        }

        const uint8_t *p = data + sd.off;
        long long rem = (long long)sd.len;
        // any alignment: phase of the file inside its 16-byte line, and what follows from it
        const u32 phase = kAligned16 ? 0u : (u32)((uintptr_t)p & 15);
        u32 delta[4] = {0, 0, 0, 0};      // smem byte offset of staged word (4m + r) minus 512 m
        u32 prmt_sel = 0x0123;            // big-endian swap, shifted by the sub-word phase
        if (!kAligned16) {
            p -= phase;                   // 16-byte aligned from here on
            const u32 ws = phase >> 2, bs = phase & 3;
#pragma unroll
            for (u32 r = 0; r < 4; r++) delta[r] = lane * 16u + (ws + r < 4 ? (ws + r) * 4u : 512u + (ws + r - 4) * 4u);
            prmt_sel = (bs + 3) | ((bs + 2) << 4) | ((bs + 1) << 8) | (bs << 12);
        }
        // source address of the copies that read nothing (src-size 0): cp.async wants it 16-byte
        // aligned all the same, and `data` need not be on the any-alignment path
        const uint8_t *const safe = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(data) & ~(uintptr_t)15);
        auto stage = [&](u32 addr, const uint8_t *q, long long left) {
            if (kAligned16) stage_block(addr, lane, q, left, safe);
            else stage_block_any(addr, lane, q, left > 0 ? (long long)phase + (left < 128 ? left : 128) : 0, safe);
        };
        stage(stage0, p, rem);

        for (u32 blk = 0; blk < nblk_max; blk++) {
            const u32 cur = stage0 + (blk & 1) * kStage;
            const u32 nxt = stage0 + ((blk + 1) & 1) * kStage;
            const bool active = blk < nblk;
            const long long rem_now = rem;
            p += 128;
            rem -= 128;
            stage(nxt, p, rem);                          // block blk+1 starts moving ...
            cp_async_wait<1>();                          // ... block blk has landed
            u64 w[16];
            if (kAligned16) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(cur + lane * 16u + (u32)j * 512u));
                    w[2 * j] = be64_from_le_words(v.x, v.y);
                    w[2 * j + 1] = be64_from_le_words(v.z, v.w);
                }
            } else {
                u32 x[33];
#pragma unroll
                for (int k = 0; k < 33; k++)
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[k]) : "r"(cur + delta[k & 3] + (u32)(k >> 2) * 512u));
#pragma unroll
                for (int j = 0; j < 16; j++)      // word j = bytes [8j, 8j+8) of the block, big-endian
                    w[j] = pack64(__byte_perm(x[2 * j + 1], x[2 * j + 2], prmt_sel), __byte_perm(x[2 * j], x[2 * j + 1], prmt_sel));
            }
            if (__any_sync(0xffffffffu, active && rem_now < 128))
                pad_block(w, rem_now, final_seg && (blk + 1 == nblk), total_len);
            sha512_compress_compact<kAddMode>(st, w, active, one);
        }
        cp_async_wait<0>();

        if (have) {
            uint4 *o4 = reinterpret_cast<uint4 *>(out);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                u32 a_lo, a_hi, b_lo, b_hi;
                unpack64(st[2 * i], a_lo, a_hi);
                unpack64(st[2 * i + 1], b_lo, b_hi);
                o4[i] = make_uint4(bswap32(a_hi), bswap32(a_lo), bswap32(b_hi), bswap32(b_lo));
            }
        }

        u32 next = 0;
        if (lane == 0) {
            if (!two_ended) {
                next = atomicAdd(unit_counter, 1u) + total_warps;
            } else {
                const bool slow = nblk_max != 0 && (clock64() - t_unit) > kSlowClocksPerBlock * (long long)nblk_max;
                const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(unit_counter),
                                                         slow ? (1ull << 32) : 1ull);
                const u32 from_head = (u32)old + head0, from_tail = (u32)(old >> 32) + tail0;
                next = from_head + from_tail >= nunits ? nunits : (slow ? nunits - 1 - from_tail : from_head);
            }
        }
        unit = __shfl_sync(0xffffffffu, next, 0);
    }
#ifdef SNAPGPU_TRACE_WARPS
    if (lane == 0) {
        const u32 wi = blockIdx.x * kShaWarpsPerCta + warp;
        if (wi < 8192) {
            unsigned long long tr_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_end));
            u32 smid, hw;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
            g_warp_trace[wi][0] = tr_start;
            g_warp_trace[wi][1] = tr_end;
            g_warp_trace[wi][2] = (tr_units << 32) | tr_blocks;
            g_warp_trace[wi][3] = ((unsigned long long)smid << 32) | hw;
        }
    }
#endif
}

}  // namespace snapgpu
