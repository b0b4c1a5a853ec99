// plan_kernels.cuh -- length binning on the device.
//
// The SHA-512 kernel wants the files of a launch ordered by 128-byte block count, longest
// first, so that the 32 lanes of a warp run the same number of blocks and the persistent grid
// is scheduled longest-processing-time-first (sha512_kernels.cuh).  The host only streams the
// descriptors into pinned memory; the ordering is a counting sort done here, three small
// launches in front of the hashing kernel on the same stream:
//
//   plan_hist_kernel     hist[min(blocks, top)]++            (warp-aggregated atomics)
//   plan_scan_kernel     hist[k] := number of files with a larger key   (descending starts)
//   plan_scatter_kernel  order[hist[key]++] = i              (warp-aggregated atomics)
//
// The order inside a bucket is whatever the atomics produce; digests do not depend on it.
#pragma once
#include "sha512_kernels.cuh"

namespace snapgpu {

constexpr int kPlanThreads = 256;
constexpr int kPlanScanThreads = 1024;
constexpr u32 kPlanTopMax = 65535;      // files of >= 65535 blocks (8 MiB) share the first bucket

__device__ __forceinline__ u32 plan_key(const SegDesc *__restrict__ descs, u32 i, u32 top) {
    const uint4 *q = reinterpret_cast<const uint4 *>(descs + i);
    const uint4 q0 = q[0], q1 = q[1];
    const u64 nb = seg_blocks(pack64(q0.z, q0.w), q1.w);
    return nb < (u64)top ? (u32)nb : top;
}

__global__ void __launch_bounds__(kPlanThreads)
plan_hist_kernel(const SegDesc *__restrict__ descs, u32 n, u32 top, u32 *__restrict__ hist) {
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * kPlanThreads + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * kPlanThreads) >> 5;
    for (u32 base = warp * 32; base < n; base += nwarps * 32) {       // uniform per warp
        const u32 i = base + lane;
        const u32 key = i < n ? plan_key(descs, i, top) : 0xffffffffu;
        const u32 peers = __match_any_sync(0xffffffffu, key);
        if (i < n && lane == (u32)__ffs(peers) - 1) atomicAdd(&hist[key], (u32)__popc(peers));
    }
}

// One CTA.  In: hist[k] = count of key k, k in [0, nbuckets).  Out: hist[k] = number of files
// whose key is larger than k, i.e. the first slot of bucket k when buckets are laid out from
// the largest key down.
__global__ void __launch_bounds__(kPlanScanThreads)
plan_scan_kernel(u32 *__restrict__ hist, u32 nbuckets) {
    __shared__ u32 part[kPlanScanThreads];
    const u32 t = threadIdx.x;
    const u32 per = (nbuckets + kPlanScanThreads - 1) / kPlanScanThreads;
    const u32 r0 = t * per, r1 = min(nbuckets, r0 + per);            // positions counted from the top bucket
    u32 sum = 0;
    for (u32 r = r0; r < r1; r++) sum += hist[nbuckets - 1 - r];
    part[t] = sum;
    __syncthreads();
    for (u32 d = 1; d < kPlanScanThreads; d <<= 1) {                  // inclusive scan of the partial sums
        const u32 v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    u32 run = part[t] - sum;                                          // exclusive prefix of this thread's range
    for (u32 r = r0; r < r1; r++) {
        const u32 k = nbuckets - 1 - r;
        const u32 c = hist[k];
        hist[k] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(kPlanThreads)
plan_scatter_kernel(const SegDesc *__restrict__ descs, u32 n, u32 top, u32 *__restrict__ cursor,
                    u32 *__restrict__ order) {
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * kPlanThreads + threadIdx.x) >> 5;
    const u32 nwarps = (gridDim.x * kPlanThreads) >> 5;
    for (u32 base = warp * 32; base < n; base += nwarps * 32) {
        const u32 i = base + lane;
        const u32 key = i < n ? plan_key(descs, i, top) : 0xffffffffu;
        const u32 peers = __match_any_sync(0xffffffffu, key);
        const u32 leader = (u32)__ffs(peers) - 1;
        u32 first = 0;
        if (i < n && lane == leader) first = atomicAdd(&cursor[key], (u32)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (i < n) order[first + __popc(peers & ((1u << lane) - 1u))] = i;
    }
}

}  // namespace snapgpu
