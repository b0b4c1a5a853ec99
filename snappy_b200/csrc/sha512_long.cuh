// sha512_long.cuh -- the long-file bin: SHA-512 of a few very long messages.
//
// One SHA-512 chain cannot be split across blocks (helpers.Sha512sum of one file,
// helpers/helpers.go:188-201, is a single Merkle-Damgard chain), and in the batched kernel one
// warp already saturates its SM sub-partition's ALU pipe: a lane hashes at ~3.9 us per block
// whatever else runs.  What CAN leave the chain is the message schedule: W[16..79] of a block
// depends on the message only, not on the chaining value -- 1280 of the ~3400 ALU instructions
// of a block.  For the handful of files that are much longer than everything else in a batch
// (the data.tar.gz of a package, config 3's 1 GiB files) this kernel splits the work:
//
//   warp 1 (producer)  32 lanes = 32 (file, block) pairs at once: loads the 128 message bytes,
//                      pads, byte-swaps, expands the schedule and stores W[t] + K[t], t = 0..79,
//                      into a ring in shared memory -- SIMT-efficient even for ONE file, because
//                      the 32 lanes work on 32 consecutive blocks of it;
//   warp 0 (consumer)  one lane per file: only the 80 rounds, reading W+K from the ring.
//
// The two warps sit on different SM sub-partitions, so the chain runs at the speed of the
// rounds alone (~2250 ALU instructions per block instead of ~3400).  Ring protocol: the
// producer publishes `produced` (block steps completely written), the consumer `consumed`;
// both are volatile words in shared memory, ordered with __threadfence_block().
#pragma once
#include "sha512_kernels.cuh"

namespace snapgpu {

constexpr int kLongThreads = 64;          // warp 0 consumer, warp 1 producer
constexpr int kLongFilesPerCta = 32;      // one consumer lane per file
// ring[step][t][file], packed for the number of files nf the CTA actually has: it holds
// 256 / nf block steps (8 for 32 files, 256 for a single file), i.e. 8 producer rounds.
constexpr int kLongSlotWords = 81;        // 80 words per (step, file) + 1 of padding: a single file's
                                          // producer lanes (32 consecutive steps) then spread over the banks
constexpr int kLongRingWords = 256 * kLongSlotWords;  // 162 KiB of dynamic shared memory
constexpr size_t kLongRingBytes = (size_t)kLongRingWords * 8;
constexpr size_t kLongSmemBytes = kLongRingBytes + 16;

__device__ __forceinline__ u32 ld_volatile_shared(const u32 *p) {
    u32 v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"((u32)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared(u32 *p, u32 v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// descs[first .. first+count) are the long segments of this CTA, count <= 32, similar lengths.
// kAligned16 = false: the producer reads 33 aligned words per block and realigns them by the
// file's byte phase (load_block<false>); the producer has time to spare, the chain is unaffected.
template <bool kAligned16>
__global__ void __launch_bounds__(kLongThreads, 1)
sha512_long_kernel(const uint8_t *__restrict__ data, const SegDesc *__restrict__ descs, u32 nsegs,
                   uint8_t *__restrict__ digests) {
    extern __shared__ __align__(16) uint8_t long_smem[];
    u64 *ring = reinterpret_cast<u64 *>(long_smem);
    u32 *produced = reinterpret_cast<u32 *>(long_smem + kLongRingBytes);
    u32 *consumed = produced + 1;

    const u32 lane = threadIdx.x & 31;
    const u32 warp = threadIdx.x >> 5;
    const u32 first = blockIdx.x * kLongFilesPerCta;
    const u32 count = min((u32)kLongFilesPerCta, nsegs - first);

    if (threadIdx.x == 0) {
        st_volatile_shared(produced, 0);
        st_volatile_shared(consumed, 0);
    }
    __syncthreads();

    // every thread needs the block counts; the longest file sets the number of steps
    u32 my_blocks = 0;
    if (lane < count) {
        const SegDesc *d = descs + first + lane;
        my_blocks = (u32)seg_blocks(d->len, d->flags);
    }
    const u32 steps = __reduce_max_sync(0xffffffffu, my_blocks);
    const u32 nf = count;                                  // files of this CTA (1..32)
    const u32 ring_steps = 256 / nf;                       // block steps the ring holds

    if (warp == 0) {
        // ---------------- consumer: the 80 rounds of file `lane` ----------------
        const bool have = lane < count;
        SegDesc sd;
        sd.off = 0; sd.len = 0; sd.prefix = 0; sd.out_idx = 0; sd.flags = kSegNoFinal;
        if (have) sd = descs[first + lane];
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
        u32 ready = 0, slot_index = 0;
        for (u32 b = 0; b < steps; b++) {
            if (b >= ready) {
                do ready = ld_volatile_shared(produced); while (b >= ready);
                __threadfence_block();
            }
            const u64 *slot = ring + (size_t)slot_index * kLongSlotWords * nf + min(lane, nf - 1);
            slot_index = slot_index + 1 == ring_steps ? 0 : slot_index + 1;
            u64 a = st[0], bb = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
            for (int grp = 0; grp < 5; grp++) {
                const u64 *kw = slot + (size_t)grp * 16 * nf;
#define SNAPGPU_LR(A, B, C, D, E, F, G, H, I)                                                    \
    {                                                                                            \
        const u64 t1 = H + big_sigma1(E) + ch64(E, F, G) + kw[(I) * nf];           \
        const u64 t2 = big_sigma0(A) + maj64(A, B, C);                                           \
        D += t1;                                                                                 \
        H = t1 + t2;                                                                             \
    }
                SNAPGPU_LR(a, bb, c, d, e, f, g, h, 0)  SNAPGPU_LR(h, a, bb, c, d, e, f, g, 1)
                SNAPGPU_LR(g, h, a, bb, c, d, e, f, 2)  SNAPGPU_LR(f, g, h, a, bb, c, d, e, 3)
                SNAPGPU_LR(e, f, g, h, a, bb, c, d, 4)  SNAPGPU_LR(d, e, f, g, h, a, bb, c, 5)
                SNAPGPU_LR(c, d, e, f, g, h, a, bb, 6)  SNAPGPU_LR(bb, c, d, e, f, g, h, a, 7)
                SNAPGPU_LR(a, bb, c, d, e, f, g, h, 8)  SNAPGPU_LR(h, a, bb, c, d, e, f, g, 9)
                SNAPGPU_LR(g, h, a, bb, c, d, e, f, 10) SNAPGPU_LR(f, g, h, a, bb, c, d, e, 11)
                SNAPGPU_LR(e, f, g, h, a, bb, c, d, 12) SNAPGPU_LR(d, e, f, g, h, a, bb, c, 13)
                SNAPGPU_LR(c, d, e, f, g, h, a, bb, 14) SNAPGPU_LR(bb, c, d, e, f, g, h, a, 15)
#undef SNAPGPU_LR
            }
            if (b < my_blocks) {
                st[0] += a; st[1] += bb; st[2] += c; st[3] += d;
                st[4] += e; st[5] += f; st[6] += g; st[7] += h;
            }
            // this step's slot may be overwritten once every lane has read it
            __syncwarp();
            if (lane == 0) st_volatile_shared(consumed, b + 1);
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
    } else {
        // ---------------- producer: W[t] + K[t] of 32 (file, block) pairs per round ----------------
        // lane -> (file = lane % nf, step offset = lane / nf); a round covers `per` consecutive steps
        const u32 per = 32 / nf;                           // steps per round (>= 1); ring_steps >= 8 * per
        const u32 pf = lane % nf, ps = lane / nf;
        const bool worker = ps < per;                      // 32 % nf leftover lanes idle
        SegDesc sd;
        sd.off = 0; sd.len = 0; sd.prefix = 0; sd.out_idx = 0; sd.flags = kSegNoFinal;
        sd = descs[first + pf];
        const u32 f_blocks = (u32)seg_blocks(sd.len, sd.flags);
        const bool final_seg = !(sd.flags & kSegNoFinal);
        const u64 total_len = sd.prefix + sd.len;
        u32 done = 0;                                      // steps published so far
        while (done < steps) {
            // room in the ring for `per` more steps?
            const u32 upto = min(steps, done + per);
            u32 freed;
            // the producer is far faster than the consumer: back off instead of hammering the
            // shared-memory pipe the consumer's loads go through
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
                u64 *slot = ring + (size_t)(b % ring_steps) * kLongSlotWords * nf + pf;
#pragma unroll 1
                for (int grp = 0; grp < 5; grp++) {
#pragma unroll
                    for (int i = 0; i < 16; i++) slot[(size_t)(grp * 16 + i) * nf] = w[i] + c_K512[grp * 16 + i];
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
