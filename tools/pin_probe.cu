// pin_probe -- what pinning host memory costs, by how the memory was obtained: cudaHostAlloc against
// cudaHostRegister of an anonymous mapping with 4 KiB pages and with transparent huge pages; and the
// host-to-device copy rate from each.   nvcc -O2 -o pin_probe pin_probe.cu;  pin_probe [MiB]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <time.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char **argv) {
    const size_t mib = argc > 1 ? atoi(argv[1]) : 64, bytes = mib << 20;
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    void *d;
    CK(cudaMalloc(&d, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto rate = [&](void *h) { float ms; cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e0); for (int i = 0; i < 4; i++) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); return 4.0 * bytes / ms / 1e6; };
    for (int rep = 0; rep < 3; rep++) {
        double t0 = now();
        void *h;
        CK(cudaHostAlloc(&h, bytes, cudaHostAllocPortable));
        double t1 = now();
        memset(h, 1, bytes);
        double t2 = now();
        printf("{\"how\": \"cudaHostAlloc\", \"mib\": %zu, \"alloc_ms\": %.2f, \"first_touch_ms\": %.2f, \"h2d_gbs\": %.1f}\n", mib, t1 - t0, t2 - t1, rate(h));
        t0 = now(); cudaFreeHost(h); printf("  free %.2f ms\n", now() - t0);
        for (int huge = 0; huge < 2; huge++) {
            t0 = now();
            void *m = mmap(nullptr, bytes + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            void *a = (void *)(((uintptr_t)m + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
            int adv = huge ? madvise(a, bytes, MADV_HUGEPAGE) : madvise(a, bytes, MADV_NOHUGEPAGE);
            memset(a, 1, bytes);
            t1 = now();
            cudaError_t e = cudaHostRegister(a, bytes, cudaHostRegisterPortable);
            t2 = now();
            printf("{\"how\": \"mmap + touch + cudaHostRegister, %s\", \"mib\": %zu, \"madvise_rc\": %d, \"map_touch_ms\": %.2f, \"register_ms\": %.2f, \"rc\": \"%s\", \"h2d_gbs\": %.1f}\n",
                   huge ? "MADV_HUGEPAGE" : "4 KiB pages", mib, adv, t1 - t0, t2 - t1, cudaGetErrorString(e), e == cudaSuccess ? rate(a) : 0.0);
            t0 = now(); if (e == cudaSuccess) cudaHostUnregister(a); munmap(m, bytes + (2 << 20)); printf("  unregister+unmap %.2f ms\n", now() - t0);
        }
    }
    FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    if (f) { char line[128]; if (fgets(line, sizeof line, f)) printf("thp enabled: %s", line); fclose(f); }
    return 0;
}
