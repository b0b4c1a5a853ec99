// pipe_microbench.cuh -- register-only issue-rate probes that define the measured roofline
// of the SHA-512 kernel (SURVEY.md section 8d "step 0").
//
// Each probe runs kChains independent dependency chains per thread of one instruction class,
// so a warp always has independent work and the rate that comes out is the pipe's, not the
// latency's.  The SASS of every probe is checked in tests/test_sass.py (cuobjdump, no GPU).
#pragma once
#include <cstdint>

namespace snapgpu {

constexpr int kProbeChains = 8;
constexpr int kProbeUnroll = 16;
constexpr int kProbeThreads = 128;

enum ProbeKind : int {
    kProbeIadd3 = 0,     // IADD3   (ALU pipe)
    kProbeLop3 = 1,      // LOP3    (ALU pipe)
    kProbeShf = 2,       // SHF.R.W (ALU pipe)
    kProbeImad = 3,      // IMAD    (FMA pipe)
    kProbeImadWide = 4,  // IMAD.WIDE.U32 with 64-bit addend (FMA pipe)
    kProbeAluImad = 5,   // IADD3 / IMAD alternating, 1:1
    kProbeAluWide = 6,   // LOP3 / IMAD.WIDE alternating, 1:1
    kProbeShaMix = 7,    // 10 ALU : 3 IMAD : 3 IMAD.WIDE, the mix of the FMA-add SHA-512 kernel
    kProbeImadHi = 8,    // IMAD.HI.U32 (FMA pipe; the ">> n" half of a multiply-based rotate)
    kProbeAluImadHi = 9, // LOP3 / IMAD.HI alternating, 1:1
    kProbeRotMix = 10,   // SHF : IMAD : IMAD.HI = 2 : 1 : 1 (half of the rotates done by multiplies)
    kProbeAddX = 11,     // 64-bit add as IADD3 (carry out) + IMAD.X (carry in)
    kProbeCount = 12
};

// warp instructions of the probed classes issued per thread and outer iteration
__host__ __device__ constexpr int probe_ops_per_iter(int kind) {
    return kind == kProbeShaMix ? kProbeUnroll * 16 : kProbeUnroll * kProbeChains;
}

template <int kKind>
__global__ void __launch_bounds__(kProbeThreads)
pipe_probe_kernel(uint32_t *out, int iters, uint32_t m_in, uint32_t y_in, unsigned long long *clocks) {
    // per-thread copies, so that the operands are plain registers (not constant-bank reads)
    const uint32_t m = m_in + (threadIdx.x >> 10);
    const uint32_t y = y_in ^ threadIdx.x;
    const uint32_t m2 = (m_in << 19) | (threadIdx.x >> 10);   // 2^19: a rotate-by-13 multiplier
    uint32_t x[kProbeChains];
    uint64_t acc[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; c++) {
        x[c] = threadIdx.x * 2654435761u + c * 40503u + y;
        acc[c] = ((uint64_t)x[c] << 32) | (c + 1);
    }
    unsigned long long t0 = 0, g0 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < kProbeUnroll; u++) {
            if (kKind == kProbeShaMix) {
                // 10 ALU ops (6 SHF + 4 LOP3), 3 IMAD, 3 IMAD.WIDE: the ratio of the FMA-add kernel
#pragma unroll
                for (int c = 0; c < 6; c++) x[c] = __funnelshift_r(x[c], x[(c + 1) & 7], 7 + c);
#pragma unroll
                for (int c = 0; c < 1; c++)
                    asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c + 6]) : "r"(x[c]), "r"(y));
#pragma unroll
                for (int c = 0; c < 3; c++)
                    asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m), "r"(y));
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    uint64_t p;
                    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c + 3]), "r"(m));
                    acc[c] = p + acc[c];
                    x[c + 3] ^= (uint32_t)(acc[c] >> 32);
                }
            } else {
#pragma unroll
                for (int c = 0; c < kProbeChains; c++) {
                    const bool alt = (c & 1) != 0;
                    if (kKind == kProbeIadd3 || (kKind == kProbeAluImad && !alt)) {
                        asm("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(x[c]) : "r"(y), "r"(m));
                    } else if (kKind == kProbeLop3 || (kKind == kProbeAluWide && !alt)) {
                        asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(m));
                    } else if (kKind == kProbeShf) {
                        x[c] = __funnelshift_r(x[c], y, 13);
                    } else if (kKind == kProbeImad || (kKind == kProbeAluImad && alt)) {
                        asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m), "r"(y));
                    } else if (kKind == kProbeImadHi || (kKind == kProbeAluImadHi && alt)) {
                        asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m2), "r"(y));
                    } else if (kKind == kProbeAluImadHi) {
                        asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(m));
                    } else if (kKind == kProbeRotMix) {
                        if ((c & 3) < 2) x[c] = __funnelshift_r(x[c], y, 13);
                        else if ((c & 3) == 2) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m), "r"(y));
                        else asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(m2), "r"(y));
                    } else if (kKind == kProbeAddX) {
                        // (x[c], x[c^1]) as a 64-bit accumulator for even c; odd c is its high half
                        if (!alt)
                            asm("{add.cc.u32 %0, %0, %2;\n\tmadc.lo.u32 %1, %1, %3, %4;}"
                                : "+r"(x[c]), "+r"(x[c + 1]) : "r"(y), "r"(m), "r"(m2));
                    } else {   // IMAD.WIDE.U32 with a register-pair addend
                        uint64_t p;
                        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"((uint32_t)acc[c]), "r"(m));
                        acc[c] = p + acc[c];   // the chain runs through the 64-bit accumulator
                    }
                }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < kProbeChains; c++) r ^= x[c] ^ (uint32_t)acc[c] ^ (uint32_t)(acc[c] >> 32);
    if (threadIdx.x == 0 && clocks) {
        unsigned long long t1 = clock64(), g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        if (blockIdx.x == 0) {
            clocks[0] = t1 - t0;
            clocks[1] = g1 - g0;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace snapgpu
