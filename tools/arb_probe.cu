// arb_probe.cu -- how does an SM sub-partition share its ALU pipe between co-resident warps?
// R CTAs of 4 warps per SM run the same ALU-bound loop for a fixed number of cycles; every warp
// reports how many iterations it got, with its SM, hardware warp slot and CTA index.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/arb_probe tools/arb_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <map>
#include <cuda_runtime.h>

struct Rec { unsigned iters, smid, hwwarp, cta, warp; unsigned long long t0, t1; };

__global__ void __launch_bounds__(128) arb_kernel(Rec *out, unsigned long long cycles, unsigned y_in) {
    unsigned x[8];
    const unsigned y = y_in ^ threadIdx.x;
#pragma unroll
    for (int c = 0; c < 8; c++) x[c] = threadIdx.x * 2654435761u + c * 40503u;
    const unsigned long long t0 = clock64();
    unsigned iters = 0;
    for (;;) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (c & 1) asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(x[(c + 1) & 7]));
                else x[c] = __funnelshift_r(x[c], y, 13);
            }
        }
        iters++;
        if ((iters & 7) == 0 && clock64() - t0 > cycles) break;
    }
    unsigned r = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) r ^= x[c];
    if ((threadIdx.x & 31) == 0) {
        Rec rec;
        rec.iters = iters + (r == 0x12345678u);
        asm volatile("mov.u32 %0, %%smid;" : "=r"(rec.smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(rec.hwwarp));
        rec.cta = blockIdx.x;
        rec.warp = threadIdx.x >> 5;
        rec.t0 = t0;
        rec.t1 = clock64();
        out[blockIdx.x * 4 + rec.warp] = rec;
    }
}

int main(int argc, char **argv) {
    int R = argc > 1 ? atoi(argv[1]) : 2;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * R;
    Rec *d;
    cudaMalloc(&d, sizeof(Rec) * grid * 4);
    for (int rep = 0; rep < 2; rep++) {
        arb_kernel<<<grid, 128>>>(d, 3000000ull, 0x9e3779b9u);
        cudaDeviceSynchronize();
    }
    std::vector<Rec> h(grid * 4);
    cudaMemcpy(h.data(), d, sizeof(Rec) * h.size(), cudaMemcpyDeviceToHost);
    // group by (smid, hwwarp & 3)
    std::map<std::pair<unsigned, unsigned>, std::vector<Rec>> g;
    for (auto &r : h) g[{r.smid, r.hwwarp & 3}].push_back(r);
    printf("R=%d CTAs/SM, %d SMs; first groups (smid, hwwarp&3): [cta warp hwwarp iters]\n", R, p.multiProcessorCount);
    int shown = 0;
    double share_by_rank[8] = {0};
    long groups = 0, sizes[8] = {0};
    for (auto &kv : g) {
        auto v = kv.second;
        std::sort(v.begin(), v.end(), [](const Rec &a, const Rec &b) { return a.hwwarp < b.hwwarp; });
        if (shown < 8) {
            printf("  sm %3u q%u:", kv.first.first, kv.first.second);
            for (auto &r : v) printf("  [cta %4u w%u hw %2u  %6u]", r.cta, r.warp, r.hwwarp, r.iters);
            printf("\n");
            shown++;
        }
        sizes[std::min<size_t>(v.size(), 7)]++;
        if ((int)v.size() == R) {
            double tot = 0;
            for (auto &r : v) tot += r.iters;
            for (int k = 0; k < R; k++) share_by_rank[k] += v[k].iters / tot;
            groups++;
        }
    }
    printf("group sizes:");
    for (int k = 0; k < 8; k++) if (sizes[k]) printf("  %d warps: %ld groups", k, sizes[k]);
    printf("\nmean share of the sub-partition by hardware-warp-slot rank (lowest slot first), %ld groups of %d:", groups, R);
    for (int k = 0; k < R; k++) printf("  %.3f", share_by_rank[k] / groups);
    // is the lowest slot the oldest CTA?
    long low_is_oldest = 0;
    for (auto &kv : g) {
        auto v = kv.second;
        if ((int)v.size() != R) continue;
        auto lo = *std::min_element(v.begin(), v.end(), [](const Rec &a, const Rec &b) { return a.hwwarp < b.hwwarp; });
        auto old = *std::min_element(v.begin(), v.end(), [](const Rec &a, const Rec &b) { return a.cta < b.cta; });
        low_is_oldest += lo.cta == old.cta;
    }
    printf("\nlowest slot holds the lowest CTA index in %ld of %ld groups\n", low_is_oldest, groups);
    return 0;
}
