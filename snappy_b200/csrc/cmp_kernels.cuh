// cmp_kernels.cuh -- batched byte-wise equality of file pairs, HBM-bound.
//
// Device side of snapgpu_cmp_batch[_device]; replaces streamsEqual / bytes.Equal
// (reference: helpers/cmp.go:61-86, chunk size bufsz = 16 KiB at helpers/cmp.go:27).
//
// Only a boolean per pair leaves the reference, so the order in which chunks are examined
// is free.  Each pair is cut into 16 KiB tiles (the reference's own chunk); the tiles of the
// whole batch form one list, split into contiguous ranges, one per CTA, so that a CTA
// streams long runs of memory.  A warp loads 4 x 512 B coalesced rows of both streams with
// 128-bit loads (8 independent loads in flight per lane), compares, and votes with a warp
// ballot.  Early-out: a mismatch clears equal[pair]; a warp that has seen (or reads at the
// pair's first tile) a cleared flag skips the remaining tiles of that pair.
#pragma once
#include <cstdint>

namespace snapgpu {

constexpr int kCmpThreads = 256;
constexpr int kCmpTileBytes = 16 * 1024;

struct CmpPair {
    uint64_t off;         // byte offset of the pair in both packed buffers
    uint64_t len;         // bytes to compare
    uint64_t first_tile;  // index of this pair's first tile in the batch-wide tile list
};

__device__ __forceinline__ uint4 cmp_ldg_stream(const void *p) {
    uint4 v;   // read-once data: do not keep it in L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t diff_bits(uint4 a, uint4 b) {
    return (a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w);
}

// pairs[npairs] carries a sentinel entry at [npairs] with first_tile = total tiles.
template <bool kAligned16>
__global__ void __launch_bounds__(kCmpThreads)
cmp_pairs_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b,
                 const CmpPair *__restrict__ pairs, uint32_t npairs, uint64_t ntiles,
                 uint8_t *__restrict__ equal /* preset to 1 */) {
    // contiguous tile range of this CTA
    const uint64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    uint64_t tile = per * blockIdx.x;
    uint64_t tile_end = tile + per < ntiles ? tile + per : ntiles;
    if (tile >= tile_end) return;

    // binary search: last pair with first_tile <= tile (and at least one tile)
    uint32_t lo = 0, hi = npairs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (pairs[mid].first_tile <= tile) lo = mid; else hi = mid;
    }
    uint32_t pi = lo;

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;

    while (tile < tile_end) {
        while (pairs[pi + 1].first_tile <= tile) pi++;          // skip empty pairs too
        const CmpPair pr = pairs[pi];
        const uint64_t pair_tiles_end = pairs[pi + 1].first_tile;
        const uint64_t stop = pair_tiles_end < tile_end ? pair_tiles_end : tile_end;
        bool known_diff = equal[pi] == 0;
        for (; tile < stop && !known_diff; tile++) {
            const uint64_t tbase = (tile - pr.first_tile) * (uint64_t)kCmpTileBytes;
            uint32_t d = 0;
            if (kAligned16) {
                // warp w owns bytes [2048 w, 2048 (w+1)) of the tile: 4 rows of 512 B
                const uint64_t wbase = tbase + (uint64_t)warp * 2048 + lane * 16;
                uint4 va[4], vb[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint64_t o = wbase + r * 512;
                    va[r] = make_uint4(0, 0, 0, 0);
                    vb[r] = make_uint4(0, 0, 0, 0);
                    if (o < pr.len) {
                        va[r] = cmp_ldg_stream(a + pr.off + o);
                        vb[r] = cmp_ldg_stream(b + pr.off + o);
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint64_t o = wbase + r * 512;
                    uint32_t x = diff_bits(va[r], vb[r]);
                    if (o + 16 > pr.len && o < pr.len) {
                        // last, partial 16 bytes of the pair: ignore what lies past the end
                        const int keep = (int)(pr.len - o);           // 1..15 bytes
                        uint32_t w[4] = {va[r].x ^ vb[r].x, va[r].y ^ vb[r].y,
                                         va[r].z ^ vb[r].z, va[r].w ^ vb[r].w};
                        x = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            int nb = keep - 4 * k;
                            nb = nb < 0 ? 0 : (nb > 4 ? 4 : nb);
                            uint32_t m = nb == 0 ? 0u : (nb == 4 ? 0xffffffffu : ((1u << (8 * nb)) - 1u));
                            x |= w[k] & m;
                        }
                    }
                    d |= x;
                }
            } else {
                // any alignment: bytes, 64 per thread and tile
                for (int k = 0; k < kCmpTileBytes / kCmpThreads; k++) {
                    const uint64_t o = tbase + (uint64_t)k * kCmpThreads + threadIdx.x;
                    if (o < pr.len) d |= (uint32_t)(a[pr.off + o] ^ b[pr.off + o]);
                }
            }
            if (__ballot_sync(0xffffffffu, d != 0) != 0) {
                if (lane == 0) equal[pi] = 0;
                known_diff = true;
            }
        }
        tile = stop;
    }
}

// ---- synthetic content, generated in place in HBM ----------------

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

struct SynthFile {
    uint64_t off;
    uint64_t len;
};

// one CTA per file (grid-stride); word j of file i = splitmix64(seed + i*GOLDEN + j)
__global__ void synth_fill_kernel(uint8_t *__restrict__ data, const SynthFile *__restrict__ files,
                                  uint32_t nfiles, uint64_t first_index, uint64_t seed) {
    for (uint32_t f = blockIdx.x; f < nfiles; f += gridDim.x) {
        const SynthFile sf = files[f];
        const uint64_t base = seed + (first_index + f) * 0x9E3779B97F4A7C15ULL;
        uint8_t *dst = data + sf.off;
        const uint64_t nwords = sf.len >> 3;
        if (((uintptr_t)dst & 7) == 0) {
            uint64_t *d64 = reinterpret_cast<uint64_t *>(dst);
            for (uint64_t j = threadIdx.x; j < nwords; j += blockDim.x) d64[j] = splitmix64(base + j);
        } else {
            for (uint64_t j = threadIdx.x; j < nwords; j += blockDim.x) {
                uint64_t v = splitmix64(base + j);
                for (int k = 0; k < 8; k++) dst[8 * j + k] = (uint8_t)(v >> (8 * k));
            }
        }
        if (threadIdx.x == 0) {
            uint64_t v = splitmix64(base + nwords);
            for (uint64_t k = 0; k < (sf.len & 7); k++) dst[8 * nwords + k] = (uint8_t)(v >> (8 * k));
        }
    }
}

}  // namespace snapgpu
