// snapgpu.cu -- CUDA runtime layer of libsnapgpu: device contexts, length-binned launch
// plans, the double-buffered host->device pipeline, the file-list sharder for 1/2/4/8 GPUs,
// and the extern "C" batch entry points declared in include/snapgpu.h.
//
// There is deliberately no CPU implementation of either hot op in this library: if CUDA is
// unavailable every compute entry point returns SNAPGPU_ECUDA / SNAPGPU_ENOINIT.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include <time.h>
#include <sys/mman.h>
#include <map>

#include "cmp_kernels.cuh"
#include "pipe_microbench.cuh"
#include "plan_kernels.cuh"
#include "runtime.hpp"
#include "sha512_kernels.cuh"
#include "sha512_long.cuh"
#include "sha512_pair.cuh"

namespace snapgpu {

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define SG_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t sg_e_ = (call);                                                           \
        if (sg_e_ != cudaSuccess)                                                             \
            return fail(SNAPGPU_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(sg_e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

std::string hex_lower(const uint8_t *p, size_t n) {
    static const char d[] = "0123456789abcdef";
    std::string s(2 * n, '0');
    for (size_t i = 0; i < n; i++) {
        s[2 * i] = d[p[i] >> 4];
        s[2 * i + 1] = d[p[i] & 15];
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// device context
// ------------------------------------------------------------------------------------------

constexpr int kPlanSlots = 4;
constexpr int kStageBufs = 4;
constexpr size_t kCounterBytes = 64;                     // the launch's unit counter
constexpr int kFeeders = 8;                     // at most this many host threads move pageable memory into pinned bounce buffers
constexpr size_t kBounceBytes = 4u << 20;
constexpr size_t kMaxChunkItems = 1u << 20;
constexpr uint64_t kMaxSegBytes = 1ULL << 39;   // block counts stay below 2^32
constexpr size_t kStageSlack = 256;

struct PlanSlot {
    void *h_buf = nullptr;    // pinned
    void *d_buf = nullptr;
    size_t cap = 0;           // bytes
    u32 *d_counter = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t uploaded = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;   // around the long-file kernel on its own stream
    bool in_flight = false;
};

struct TimedLaunch {
    cudaEvent_t beg = nullptr, end = nullptr;
    cudaEvent_t plan_ready = nullptr;   // trace only: the slot's "uploaded" event of this launch
    cudaEvent_t prev_end = nullptr;     // trace only: end event of the launch before
    bool pending = false;
    bool recorded = false;
};

// One pipeline on one GPU: streams, plan slots, staging and result buffers, timing ring.  A device
// has kPipes of them so that concurrent callers (goroutines through cgo, the archive's chain
// beside the tree's batches) overlap one's copies with another's kernels instead of queueing
// behind a per-device lock.  Everything in a Pipe is used under its own mutex only.
struct Pipe {
    int ordinal = -1;
    int sm_count = 0;
    std::mutex mu;
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr, long_stream = nullptr;
    cudaStream_t slot_long_stream[kStageBufs] = {};
    cudaStream_t slot_stream[kStageBufs] = {};   // batch sessions: one compute stream per staging buffer, so that the
                                                 // kernels of consecutive batches (each at least as long as the chain of
                                                 // its longest file, mostly on a fraction of the SMs) run side by side
    PlanSlot slots[kPlanSlots];
    int next_slot = 0;
    // host-buffer pipeline (allocated on first use)
    uint8_t *d_stage[kStageBufs] = {};   // the batch calls use two, a batch session all of them
    size_t stage_cap = 0;
    int stage_n = 0;          // buffers allocated at stage_cap
    uint8_t *d_out[kStageBufs] = {};
    uint8_t *h_out[kStageBufs] = {};
    size_t out_cap = 0;       // bytes
    int out_n = 0;
    cudaEvent_t ev_copied[kStageBufs] = {}, ev_done[kStageBufs] = {};
    // SNAPGPU_TRACE only: device timestamps of a session batch (copy begin / end, kernel begin, all done) and their base
    cudaEvent_t tr_ev[kStageBufs][4] = {}, tr_base = nullptr;
    double tr_base_host = 0;
    // bounce buffers for callers whose host buffer is ordinary pageable memory (allocated on first use)
    uint8_t *bounce[kFeeders][2] = {};
    cudaEvent_t bounce_free[kFeeders][2] = {};
    cudaEvent_t feeder_done[kFeeders] = {};
    cudaEvent_t feeder_fork = nullptr;
    cudaStream_t feeder_stream[kFeeders] = {};
    // kernel timing ring (events on the launching stream)
    TimedLaunch sha_t[8], cmp_t[8];
    int sha_ti = 0, cmp_ti = 0;
    double sha_ms_sum = 0, cmp_ms_sum = 0, sha_ms_last = 0, cmp_ms_last = 0;
    uint64_t sha_ms_n = 0, cmp_ms_n = 0;
};

constexpr int kPipes = 3;

struct Device {
    int ordinal = -1;
    int sm_count = 0;
    std::unique_ptr<Pipe> pipes[kPipes];
    std::mutex lease_mu;                  // guards busy[] and sessions
    std::condition_variable lease_cv;
    bool busy[kPipes] = {};
    int sessions = 0;
};

// A caller's hold on one pipe of a device: the lowest-numbered free one (a lone caller always
// gets pipe 0 and its warm buffers); when all are taken it waits for whichever is released
// first.  A batch session keeps its pipe for a long time and its owner makes short calls of its
// own meanwhile (the archive's chain), so sessions may hold at most kPipes - 1 pipes of a
// device: one is always left to the calls that come and go.
struct PipeLease {
    Device &dev;
    Pipe *pipe = nullptr;
    int index = -1;
    const bool session;
    explicit PipeLease(Device &d, bool for_session = false) : dev(d), session(for_session) {
        {
            std::unique_lock<std::mutex> lk(dev.lease_mu);
            dev.lease_cv.wait(lk, [&] {
                if (session && dev.sessions >= kPipes - 1) return false;
                for (int i = 0; i < kPipes; i++)
                    if (!dev.busy[i]) {
                        index = i;
                        return true;
                    }
                return false;
            });
            dev.busy[index] = true;
            if (session) dev.sessions++;
        }
        pipe = dev.pipes[index].get();
        pipe->mu.lock();                  // contended only by the statistics readers
    }
    ~PipeLease() {
        pipe->mu.unlock();
        {
            std::lock_guard<std::mutex> lk(dev.lease_mu);
            dev.busy[index] = false;
            if (session) dev.sessions--;
        }
        dev.lease_cv.notify_all();
    }
    PipeLease(const PipeLease &) = delete;
    PipeLease &operator=(const PipeLease &) = delete;
};

struct Options {
    std::atomic<long long> staging_bytes{1024ll << 20};   // fewer, larger H2D copies: 55 vs 47 GB/s measured
    std::atomic<long long> sha_warps_per_sm{0};
    std::atomic<long long> sha_variant{0};
    std::atomic<long long> cmp_ctas_per_sm{0};
    std::atomic<long long> time_kernels{1};
    std::atomic<long long> feeders{0};          // bounce-buffer threads per device for pageable input, 0 = auto
    std::atomic<long long> long_kernel{2};      // 0 off, 1 one lane per file, 2 a lane pair per file
    std::atomic<long long> two_ended{1};        // several CTAs per SM: slow warps claim from the short end of the plan
    std::atomic<long long> pair_form{0};        // lane-pair kernel: 0 lanes exchange through mailboxes, 1 by shuffle
    std::atomic<long long> taper{1};            // host-buffer pipeline: small last chunk (see taper_cap)
    std::atomic<long long> long_min_blocks{0};  // long-file bin: candidates have at least this many blocks (0 = by form)
    std::atomic<long long> pair_files_per_cta{0};   // lane-pair kernel: 0 = spread over the SMs, 1..16 = exactly this many
};

// devs is written only by snapgpu_init / snapgpu_shutdown, under mu AND devs_mu held exclusively;
// every entry point that touches a device holds devs_mu shared for the duration of its call, so
// a first call racing with snapgpu_init sees either no devices or all of them, and a re-init or
// shutdown waits for the calls in flight instead of destroying pipes under them.
struct Runtime {
    std::mutex mu;
    std::shared_mutex devs_mu;
    std::vector<std::unique_ptr<Device>> devs;
    Options opt;
    std::atomic<uint64_t> kernel_launches{0}, sha_launches{0}, sha_long_launches{0}, cmp_launches{0}, h2d_bytes{0},
        d2h_bytes{0};
};

// Pinned host memory.  cudaHostAlloc pins page by page: 0.45 ms per MiB on this pool's boxes (114 ms for 256 MiB,
// tools/pin_probe.cu), which is what the first large tree of a process used to wait for.  From 2 MiB up the memory is
// instead an anonymous mapping on transparent huge pages (madvise; 512x fewer pages to pin), faulted in by a few
// threads at once and then registered: 256 MiB in ~10 ms, same copy rate (55 GB/s).  Anything that fails on the way
// falls back to cudaHostAlloc.  SNAPGPU_PIN=hostalloc forces the old path, SNAPGPU_PIN=nohuge the mapping without huge
// pages (for A/B runs: warm, the three are indistinguishable; cold, they are 115 / 10-50 / 65 ms per 256 MiB).
namespace {
struct PinnedMapping {
    void *map_base;
    size_t map_len;
};
std::mutex g_pinned_mu;
std::map<void *, PinnedMapping> g_pinned_maps;       // registered mappings by user pointer

void *pinned_huge_alloc(size_t bytes) {
    constexpr size_t kHuge = (size_t)2 << 20;
    const size_t len = (bytes + kHuge - 1) & ~(kHuge - 1), map_len = len + kHuge;
    void *m = mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) return nullptr;
    uint8_t *a = reinterpret_cast<uint8_t *>(((uintptr_t)m + kHuge - 1) & ~(uintptr_t)(kHuge - 1));
#ifdef MADV_HUGEPAGE
    static const bool no_huge = [] { const char *e = getenv("SNAPGPU_PIN"); return e && !strcmp(e, "nohuge"); }();
    if (!no_huge) madvise(a, len, MADV_HUGEPAGE);    // refused or unsupported: 4 KiB pages, still correct
#endif
#ifdef MADV_DONTFORK
    madvise(a, len, MADV_DONTFORK);                  // a child (tar, gzip ...) neither copies nor shares DMA memory
#endif
    // fault the pages in (zeroing them is the cost) on several threads; one touch per 4 KiB page covers both page sizes
    const size_t nthreads = std::max<size_t>(1, std::min<size_t>({(size_t)8, len / ((size_t)16 << 20),
                                                                  (size_t)std::max(1u, std::thread::hardware_concurrency())}));
    auto touch = [a, len, nthreads](size_t t) {
        const size_t lo = len / nthreads * t, hi = t + 1 == nthreads ? len : len / nthreads * (t + 1);
        for (size_t o = lo; o < hi; o += 4096) reinterpret_cast<volatile uint8_t *>(a)[o] = 0;
    };
    std::vector<std::thread> th;
    for (size_t t = 1; t < nthreads; t++) th.emplace_back(touch, t);
    touch(0);
    for (auto &x : th) x.join();
    if (cudaHostRegister(a, len, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(m, map_len);
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned_maps[a] = PinnedMapping{m, map_len};
    return a;
}

// pinned memory for the runtime's own buffers (plan slots, digest buffers, bounce buffers)
cudaError_t host_pinned_alloc(void **out, size_t bytes) {
    static const bool force_hostalloc = [] { const char *e = getenv("SNAPGPU_PIN"); return e && !strcmp(e, "hostalloc"); }();
    if (bytes >= ((size_t)2 << 20) && !force_hostalloc && (*out = pinned_huge_alloc(bytes))) return cudaSuccess;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
}
void host_pinned_free(void *p) {
    if (!p) return;
    PinnedMapping pm{nullptr, 0};
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        auto it = g_pinned_maps.find(p);
        if (it != g_pinned_maps.end()) {
            pm = it->second;
            g_pinned_maps.erase(it);
        }
    }
    if (pm.map_base) {
        cudaHostUnregister(p);
        munmap(pm.map_base, pm.map_len);
    } else {
        cudaFreeHost(p);
    }
}
}  // namespace


static Runtime &rt() {
    static Runtime r;
    return r;
}

typedef std::shared_lock<std::shared_mutex> DevsInUse;

bool runtime_ready() {
    DevsInUse use(rt().devs_mu);
    return !rt().devs.empty();
}
size_t staging_bytes() { return (size_t)rt().opt.staging_bytes.load(); }

static void destroy_pipe(Pipe &D) {
    if (D.ordinal < 0) return;
    cudaSetDevice(D.ordinal);
    cudaDeviceSynchronize();
    for (auto &s : D.slots) {
        if (s.h_buf) host_pinned_free(s.h_buf);
        if (s.d_buf) cudaFree(s.d_buf);
        if (s.d_counter) cudaFree(s.d_counter);
        if (s.done) cudaEventDestroy(s.done);
        if (s.uploaded) cudaEventDestroy(s.uploaded);
        if (s.fork) cudaEventDestroy(s.fork);
        if (s.join) cudaEventDestroy(s.join);
    }
    for (int b = 0; b < kStageBufs; b++) {
        if (D.d_stage[b]) cudaFree(D.d_stage[b]);
        if (D.d_out[b]) cudaFree(D.d_out[b]);
        if (D.h_out[b]) host_pinned_free(D.h_out[b]);
        if (D.ev_copied[b]) cudaEventDestroy(D.ev_copied[b]);
        for (auto &e : D.tr_ev[b])
            if (e) { cudaEventDestroy(e); e = nullptr; }
        if (D.ev_done[b]) cudaEventDestroy(D.ev_done[b]);
    }
    if (D.tr_base) { cudaEventDestroy(D.tr_base); D.tr_base = nullptr; }
    for (auto &t : D.sha_t) { if (t.beg) cudaEventDestroy(t.beg); if (t.end) cudaEventDestroy(t.end); }
    for (auto &t : D.cmp_t) { if (t.beg) cudaEventDestroy(t.beg); if (t.end) cudaEventDestroy(t.end); }
    for (int f = 0; f < kFeeders; f++) {
        for (int k = 0; k < 2; k++) {
            if (D.bounce[f][k]) host_pinned_free(D.bounce[f][k]);
            if (D.bounce_free[f][k]) cudaEventDestroy(D.bounce_free[f][k]);
            D.bounce[f][k] = nullptr;
            D.bounce_free[f][k] = nullptr;
        }
        if (f == 0 && D.feeder_fork) { cudaEventDestroy(D.feeder_fork); D.feeder_fork = nullptr; }
        if (D.feeder_done[f]) cudaEventDestroy(D.feeder_done[f]);
        if (D.feeder_stream[f]) cudaStreamDestroy(D.feeder_stream[f]);
        D.feeder_done[f] = nullptr;
        D.feeder_stream[f] = nullptr;
    }
    if (D.copy_stream) cudaStreamDestroy(D.copy_stream);
    if (D.compute_stream) cudaStreamDestroy(D.compute_stream);
    if (D.long_stream) cudaStreamDestroy(D.long_stream);
    for (auto &st : D.slot_stream)
        if (st) cudaStreamDestroy(st);
    for (auto &st : D.slot_long_stream)
        if (st) cudaStreamDestroy(st);
    D.ordinal = -1;
}

static int init_pipe(Pipe &D, int ordinal, int sm_count) {
    D.ordinal = ordinal;
    D.sm_count = sm_count;
    SG_CUDA(cudaSetDevice(ordinal));
    SG_CUDA(cudaStreamCreateWithFlags(&D.copy_stream, cudaStreamNonBlocking));
    SG_CUDA(cudaStreamCreateWithFlags(&D.compute_stream, cudaStreamNonBlocking));
    SG_CUDA(cudaStreamCreateWithFlags(&D.long_stream, cudaStreamNonBlocking));
    for (auto &st : D.slot_stream) SG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &st : D.slot_long_stream) SG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (auto &s : D.slots) {
        SG_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        SG_CUDA(cudaEventCreate(&s.uploaded));
        SG_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        SG_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
        SG_CUDA(cudaMalloc(&s.d_counter, kCounterBytes));
    }
    for (int b = 0; b < kStageBufs; b++) {
        SG_CUDA(cudaEventCreateWithFlags(&D.ev_copied[b], cudaEventDisableTiming));
        SG_CUDA(cudaEventCreateWithFlags(&D.ev_done[b], cudaEventDisableTiming));
    }
    for (auto &t : D.sha_t) { SG_CUDA(cudaEventCreate(&t.beg)); SG_CUDA(cudaEventCreate(&t.end)); }
    for (auto &t : D.cmp_t) { SG_CUDA(cudaEventCreate(&t.beg)); SG_CUDA(cudaEventCreate(&t.end)); }
    return 0;
}

typedef void (*PairKernel)(const uint8_t *, const SegDesc *, u32, uint8_t *, u32);
// the lane-pair kernel by alignment, exchange form (0 mailboxes -- the default, 1 shuffle) and, for the mailbox form,
// the number of branch-free regions a block's rounds are cut into (sha512_pair.cuh)
static PairKernel pair_kernel_for(bool aligned, int form, int regions) {
    if (form == 1) return aligned ? sha512_pair_kernel<true, true, 2> : sha512_pair_kernel<false, true, 2>;
    if (regions == 1) return aligned ? sha512_pair_kernel<true, false, 1> : sha512_pair_kernel<false, false, 1>;
    return aligned ? sha512_pair_kernel<true, false, 2> : sha512_pair_kernel<false, false, 2>;
}

static void destroy_device(Device &dev) {
    for (auto &p : dev.pipes)
        if (p) {
            std::lock_guard<std::mutex> lock(p->mu);
            destroy_pipe(*p);
        }
    dev.ordinal = -1;
}

static int init_device(Device &dev, int ordinal) {
    dev.ordinal = ordinal;
    SG_CUDA(cudaSetDevice(ordinal));
    cudaDeviceProp prop;
    SG_CUDA(cudaGetDeviceProperties(&prop, ordinal));
    dev.sm_count = prop.multiProcessorCount;
    if (prop.major < 10)
        return fail(SNAPGPU_ECUDA, "device %d is sm_%d%d; libsnapgpu is built for sm_100a only", ordinal,
                    prop.major, prop.minor);
    SG_CUDA(cudaFuncSetAttribute(sha512_long_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmemBytes));
    SG_CUDA(cudaFuncSetAttribute(sha512_long_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLongSmemBytes));
    for (int form = 0; form < 2; form++)
        for (int regions = 1; regions <= 2; regions++) {
            SG_CUDA(cudaFuncSetAttribute(pair_kernel_for(true, form, regions), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes));
            SG_CUDA(cudaFuncSetAttribute(pair_kernel_for(false, form, regions), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemBytes));
        }
    for (auto &p : dev.pipes) {
        p.reset(new Pipe());
        int rc = init_pipe(*p, ordinal, dev.sm_count);
        if (rc) return rc;
    }
    return 0;
}

static int get_device(int dev, Device **out) {
    auto &R = rt();
    if (R.devs.empty()) return fail(SNAPGPU_ENOINIT, "snapgpu_init has not been called (or failed)");
    if (dev < 0 || (size_t)dev >= R.devs.size()) return fail(SNAPGPU_EINVAL, "device index %d out of range", dev);
    *out = R.devs[(size_t)dev].get();
    return 0;
}

int ensure_init() {
    if (runtime_ready()) return 0;
    return snapgpu_init(nullptr, 0);
}

// Reserve a plan slot of at least `bytes`: waits for the launch that used it last.  When a
// slot has to grow, all of them grow together, so the cost of pinning memory is paid in the
// first call of a given size and not again three launches later.
static int acquire_slot(Pipe &D, size_t bytes, PlanSlot **out) {
    PlanSlot &s = D.slots[D.next_slot];
    D.next_slot = (D.next_slot + 1) % kPlanSlots;
    if (s.in_flight) {
        SG_CUDA(cudaEventSynchronize(s.done));
        s.in_flight = false;
    }
    if (s.cap < bytes) {
        const size_t want = std::max<size_t>(bytes + bytes / 4, 1u << 16);
        for (PlanSlot &g : D.slots) {
            if (g.cap >= want) continue;
            if (g.in_flight) {
                SG_CUDA(cudaEventSynchronize(g.done));
                g.in_flight = false;
            }
            if (g.h_buf) host_pinned_free(g.h_buf);
            if (g.d_buf) cudaFree(g.d_buf);
            g.h_buf = g.d_buf = nullptr;
            g.cap = 0;
            SG_CUDA(host_pinned_alloc(&g.h_buf, want));
            SG_CUDA(cudaMalloc(&g.d_buf, want));
            memset(g.h_buf, 0, want);          // touch the pages now, not inside a timed launch
            g.cap = want;
        }
    }
    *out = &s;
    return 0;
}

static bool trace_on();
static void harvest_timings(TimedLaunch *ring, int n, double &sum, uint64_t &cnt, double &last, bool wait) {
    for (int i = 0; i < n; i++) {
        TimedLaunch &t = ring[i];
        if (!t.pending) continue;
        if (wait) cudaEventSynchronize(t.end);
        else if (cudaEventQuery(t.end) != cudaSuccess) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.beg, t.end) == cudaSuccess) {
            sum += ms;
            cnt++;
            last = ms;
        }
        if (trace_on() && t.plan_ready) {
            float since_plan = -1, gap = -1;
            cudaEventElapsedTime(&since_plan, t.plan_ready, t.beg);
            if (t.prev_end) cudaEventElapsedTime(&gap, t.prev_end, t.beg);
            cudaGetLastError();
            fprintf(stderr, "[snapgpu] kernel %.3f ms; started %.3f ms after its plan was ready, %.3f ms after the previous kernel ended\n",
                    ms, since_plan, gap);
        }
        t.pending = false;
    }
}

// ------------------------------------------------------------------------------------------
// SHA-512 launch: length binning + persistent kernel
// ------------------------------------------------------------------------------------------

static double now_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static bool trace_on() {
    static const bool on = getenv("SNAPGPU_TRACE") != nullptr;
    return on;
}


constexpr int kShaCtasPerSmMax = 3;   // 168 registers per thread: no spills with the one-block-ahead prefetch
constexpr int kShaVariants = 6;

// variant 0 (default): compact 16-round loop, cp.async staging through shared memory, plain
//            64-bit adds (ptxas pairs them into 3-input IADD3 / IADD3.X and already sends the
//            high half of every 2-input add to the FMA pipe as IMAD.X)
// variant 1: same, every add split IADD3 (low half, ALU) / IMAD.X (high half, FMA): loses the
//            3-input merging, so the ALU count does not drop                 (measured 3-5 % slower)
// variant 2: 80 rounds fully unrolled, register prefetch, ALU adds            (round-1 first cut)
// variant 3: fully unrolled, every add as IMAD.WIDE + IMAD on the FMA pipe
// variant 4: fully unrolled, round adds on FMA, schedule adds on ALU
// variant 5: compact loop, every add as IMAD.WIDE + IMAD on the FMA pipe.  ALU instructions per
//            block drop from 3425 to ~2700, yet it measured 19-29 % slower (and every partial mix
//            between 0 and 5 fell in between, profiles/r01_sweep_fma_add_variants.txt): IMAD.WIDE
//            issues at half the ALU rate and holds up the ALU instructions issued next to it (a 1:1 mix
//            of LOP3 and IMAD.WIDE runs at 0.75 + 0.75 per clock per SM, pipe_microbench.cuh).
typedef void (*ShaKernel)(const uint8_t *, const SegDesc *, const u32 *, u32, uint8_t *, u32 *, u32, u32);
static ShaKernel sha_kernel_for(int variant, bool aligned) {
    if (!aligned) {
        // any byte alignment: the staged kernel reading from each file's own phase; variant 2
        // selects the register-load kernel (33 aligned words + funnel shifts) it replaced
        if (variant == 2) return sha512_segments_kernel<0x00, 0x0, false, kShaCtasPerSmMax>;
        return sha512_segments_kernel_v2<0, kShaCtasPerSmMax, false>;
    }
    switch (variant) {
    case 1: return sha512_segments_kernel_v2<1, kShaCtasPerSmMax>;
    case 2: return sha512_segments_kernel<0x00, 0x0, true, kShaCtasPerSmMax>;
    case 3: return sha512_segments_kernel<0x7f, 0x7, true, kShaCtasPerSmMax>;
    case 4: return sha512_segments_kernel<0x7f, 0x0, true, kShaCtasPerSmMax>;
    case 5: return sha512_segments_kernel_v2<0x1000 | 0x700 | 0x7f, kShaCtasPerSmMax>;
    default: return sha512_segments_kernel_v2<0, kShaCtasPerSmMax>;
    }
}

// Host half of the launch plan: the `n` descriptors produced by get(i) are streamed into
// pinned memory in the caller's order and checked; the ordering by length is done on the
// device (plan_kernels.cuh).  Multi-million-file shards are written by several host threads.
// The long-file bin takes the files whose chain would dominate the launch: at least `long_min_blocks`
// blocks AND at least kLongDominance times the launch's blocks per lane AND (lane-pair form) at least
// kPairBalance of the longest file -- if there are no more of them than one CTA per SM of the lane-pair
// kernel can take (16 files each: 2368 on a B200; 256 for the one-lane form).  Up to there every chain
// runs at 1.8-1.9 us per block instead of 3.8 while the batched kernel would need the same single wave;
// beyond it the batched kernel keeps everything -- it has the higher throughput.
//   long_min_blocks: 256 (32 KiB) for the lane-pair form -- a launch with nothing longer is over in
// under a millisecond whatever hashes it; what this buys is the chain-bound launch of a small batch (the
// last chunk of a host-buffer call, a tree batch, a small request): with 64 KiB files in it 1.95 ms in
// the batched kernel, 0.97 ms with the top files on lane pairs.  The one-lane form gains nothing over a
// lone warp of the batched kernel below 128 KiB and keeps 1024.
//   kPairBalance: the bin's chains run 2.1x faster than those left behind, so files shorter than
// 31/64 of the longest would finish before it in the batched kernel anyway.
constexpr uint64_t kLongMinBlocksPair = 256, kLongMinBlocksLane = 1024;
constexpr uint64_t kPairBalanceNum = 31, kPairBalanceDen = 64;
constexpr uint64_t kLongDominance = 4;
constexpr size_t kLongMaxFiles = 2400;          // descriptor room in a plan slot
constexpr size_t kLongMaxCandidates = 4096;

struct PlanInfo {
    uint64_t total_blocks = 0, max_blocks = 0;
    bool aligned = true;
    int bad = 0;            // 1: too large, 2: non-final segment not a multiple of 128
    size_t bad_index = 0;
    uint64_t min_long_blocks = kLongMinBlocksLane;   // in: what counts as a candidate for the long-file bin
    uint64_t max_short_blocks = 0;      // longest item below min_long_blocks
    size_t n_long = 0;                  // items of >= min_long_blocks blocks
    std::vector<u32> long_idx;          // their indices (the first kLongMaxCandidates)
    // second tier, for a launch with too many candidates of the first (50 000 files of which 4 000 are longer than
    // 32 KiB, and four of 1 GiB): the items of at least kLongMinBlocksLane blocks alone
    uint64_t max_mid_blocks = 0;        // longest item of the first tier only
    size_t n_very = 0;
    std::vector<u32> very_idx;
};

template <typename Get>
static void write_descriptors(Get get, size_t n, SegDesc *descs, PlanInfo *info) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nthreads = n < (1u << 19) ? 1 : std::min<size_t>({(size_t)hw, (size_t)8, n >> 18});
    std::vector<PlanInfo> part(nthreads);
    const uint64_t min_long = info->min_long_blocks;
    auto body = [&](size_t t) {
        PlanInfo &pi = part[t];
        const size_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        for (size_t i = lo; i < hi; i++) {
            const SegDesc d = get(i);
            descs[i] = d;
            if (d.len >= kMaxSegBytes && !pi.bad) { pi.bad = 1; pi.bad_index = i; }
            if ((d.flags & kSegNoFinal) && (d.len & 127) && !pi.bad) { pi.bad = 2; pi.bad_index = i; }
            pi.aligned = pi.aligned && ((d.off & 15) == 0);
            const uint64_t nb = seg_blocks(d.len, d.flags);
            pi.total_blocks += nb;
            pi.max_blocks = std::max(pi.max_blocks, nb);
            if (nb >= min_long) {
                if (pi.long_idx.size() < kLongMaxCandidates) pi.long_idx.push_back((u32)i);
                pi.n_long++;
                if (nb >= kLongMinBlocksLane) {
                    if (pi.very_idx.size() < kLongMaxCandidates) pi.very_idx.push_back((u32)i);
                    pi.n_very++;
                } else {
                    pi.max_mid_blocks = std::max(pi.max_mid_blocks, nb);
                }
            } else {
                pi.max_short_blocks = std::max(pi.max_short_blocks, nb);
            }
        }
    };
    if (nthreads == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (size_t t = 1; t < nthreads; t++) th.emplace_back(body, t);
        body(0);
        for (auto &x : th) x.join();
    }
    for (const PlanInfo &pi : part) {
        info->total_blocks += pi.total_blocks;
        info->max_blocks = std::max(info->max_blocks, pi.max_blocks);
        info->aligned = info->aligned && pi.aligned;
        if (pi.bad && !info->bad) { info->bad = pi.bad; info->bad_index = pi.bad_index; }
        info->n_long += pi.n_long;
        info->max_short_blocks = std::max(info->max_short_blocks, pi.max_short_blocks);
        for (u32 i : pi.long_idx)
            if (info->long_idx.size() < kLongMaxCandidates) info->long_idx.push_back(i);
        info->n_very += pi.n_very;
        info->max_mid_blocks = std::max(info->max_mid_blocks, pi.max_mid_blocks);
        for (u32 i : pi.very_idx)
            if (info->very_idx.size() < kLongMaxCandidates) info->very_idx.push_back(i);
    }
}

// Device half: order[] = indices sorted by block count, longest first.  Layout of the plan
// slot (host and device): SegDesc[n] | SegDesc long[kLongMaxFiles] | u32 order[n] | u32 hist[nbuckets].
struct DevicePlan {
    const SegDesc *descs;
    const SegDesc *long_descs;
    u32 *order;
    u32 *hist;
};

static DevicePlan plan_layout(void *d_buf, size_t n) {
    SegDesc *descs = static_cast<SegDesc *>(d_buf);
    u32 *order = reinterpret_cast<u32 *>(descs + n + kLongMaxFiles);
    return DevicePlan{descs, descs + n, order, order + n};
}

static size_t plan_bytes(size_t n, size_t nbuckets) {
    return (n + kLongMaxFiles) * sizeof(SegDesc) + n * sizeof(u32) + nbuckets * sizeof(u32);
}

static int enqueue_length_binning(Pipe &D, cudaStream_t stream, const DevicePlan &p, size_t n, uint64_t max_blocks) {
    const u32 top = (u32)std::min<uint64_t>(max_blocks, kPlanTopMax);
    const u32 nbuckets = top + 1;
    SG_CUDA(cudaMemsetAsync(p.hist, 0, nbuckets * sizeof(u32), stream));
    const u32 grid = (u32)std::min<size_t>((n + kPlanThreads - 1) / kPlanThreads, (size_t)D.sm_count * 8);
    plan_hist_kernel<<<grid, kPlanThreads, 0, stream>>>(p.descs, (u32)n, top, p.hist);
    plan_scan_kernel<<<1, kPlanScanThreads, 0, stream>>>(p.hist, nbuckets);
    plan_scatter_kernel<<<grid, kPlanThreads, 0, stream>>>(p.descs, (u32)n, top, p.hist, p.order);
    SG_CUDA(cudaGetLastError());
    rt().kernel_launches += 3;
    return 0;
}

// Enqueue the hashing of `n` segments (get(i) -> SegDesc) of `d_data` on `stream`.
// Which files of a launch go to the long-file bin (rule above): marks them kSegSkip in descs[0..n), appends their
// descriptors at descs[n..), longest first, takes their blocks out of *total_blocks and sets *max_blocks to the longest
// file left to the batched kernel.  Returns how many.  Pure host logic (tests/test_host_logic.py drives it through
// snapgpu_test_long_bin).
static size_t select_long_bin(const PlanInfo &info, SegDesc *h_descs, size_t n, int sm_count, long long long_mode,
                              uint64_t *total_blocks_io, uint64_t *max_blocks_io) {
    const bool pair_bin = long_mode >= 2;
    uint64_t &total_blocks = *total_blocks_io, &max_blocks = *max_blocks_io;
    size_t n_long = 0;
    const std::vector<u32> *candidates = nullptr;
    uint64_t floor_blocks = 0, rest_blocks = 0;           // smallest candidate; longest item that is not one
    if (long_mode && info.n_long >= 1 && info.n_long <= kLongMaxCandidates) {
        candidates = &info.long_idx;
        floor_blocks = info.min_long_blocks;
        rest_blocks = info.max_short_blocks;
    } else if (long_mode && info.n_very >= 1 && info.n_very <= kLongMaxCandidates) {
        candidates = &info.very_idx;
        floor_blocks = std::max(info.min_long_blocks, kLongMinBlocksLane);
        rest_blocks = std::max(info.max_short_blocks, info.max_mid_blocks);
    }
    if (candidates) {
        const uint64_t lanes = (uint64_t)sm_count * kShaThreads;
        uint64_t threshold = std::max<uint64_t>(floor_blocks, kLongDominance * (total_blocks / lanes));
        if (pair_bin) threshold = std::max(threshold, max_blocks * kPairBalanceNum / kPairBalanceDen);
        size_t dominant = 0;
        for (u32 i : *candidates) dominant += seg_blocks(h_descs[i].len, h_descs[i].flags) >= threshold;
        const size_t room = pair_bin ? std::min<size_t>(kLongMaxFiles, (size_t)sm_count * kPairFilesPerCta) : 256;
        if (dominant >= 1 && dominant <= room) {
            SegDesc *h_long = h_descs + n;
            uint64_t max_rest = rest_blocks;
            for (u32 i : *candidates) {
                SegDesc &d = h_descs[i];
                const uint64_t nb = seg_blocks(d.len, d.flags);
                if (nb < threshold) {
                    max_rest = std::max(max_rest, nb);
                    continue;
                }
                h_long[n_long++] = d;
                total_blocks -= nb;
                d.flags |= kSegSkip;
            }
            std::sort(h_long, h_long + n_long, [](const SegDesc &a, const SegDesc &b) {
                return seg_blocks(a.len, a.flags) > seg_blocks(b.len, b.flags);
            });
            max_blocks = max_rest;
        }
    }
    return n_long;
}

// Files per CTA of the lane-pair kernel for n_long files (forced: option pair_files_per_cta, 0 = this rule).  A CTA of
// that kernel has its SM to itself (sha512_pair.cuh: kPairSmemBytes), so its consumer warp shares the ALU pipe with
// nobody -- and the batched kernel has that many SMs less.  Up to a quarter of the SMs go to the bin: one file per CTA
// while that lasts (then the rounds of a block are one branch-free region, 1.81 us per block), two (1.88), and from
// there 16 per CTA (two regions, 1.90 whatever the count) on as many CTAs as it takes.
static u32 pair_cta_shape(size_t n_long, int sm_count, u32 forced) {
    const u32 budget = std::max<u32>(1, (u32)sm_count / 4);
    u32 per_cta = n_long <= budget ? 1 : n_long <= 2 * (size_t)budget ? 2 : (u32)kPairFilesPerCta;
    if (forced) per_cta = forced;
    return std::min<u32>(std::max<u32>(per_cta, 1), kPairFilesPerCta);
}

// Caller holds D.mu.  The plan goes up on the copy stream so that it overlaps whatever the
// caller's stream is still running.
template <typename Get>
static int launch_sha512(Pipe &D, cudaStream_t stream, const uint8_t *d_data, Get get, size_t n,
                         uint8_t *d_digests, cudaStream_t long_stream = nullptr) {
    if (!long_stream) long_stream = D.long_stream;
    if (n == 0) return 0;
    if (n > 0xfffffff0ull) return fail(SNAPGPU_EINVAL, "too many files in one launch (%zu)", n);
    auto &R = rt();
    PlanSlot *slot;
    // descriptors are written before the slot's final size is known: the histogram needs at
    // most kPlanTopMax + 1 buckets
    const double t_plan = now_ms();
    int rc = acquire_slot(D, plan_bytes(n, kPlanTopMax + 1), &slot);
    if (rc) return rc;
    PlanInfo info;
    const long long long_mode = R.opt.long_kernel.load();
    const bool pair_bin = long_mode >= 2;
    {
        const long long forced = R.opt.long_min_blocks.load();
        info.min_long_blocks = forced > 0 ? (uint64_t)forced : pair_bin ? kLongMinBlocksPair : kLongMinBlocksLane;
    }
    if (trace_on()) fprintf(stderr, "[snapgpu] plan slot acquired after %.3f ms\n", now_ms() - t_plan);
    write_descriptors(get, n, static_cast<SegDesc *>(slot->h_buf), &info);
    if (trace_on()) fprintf(stderr, "[snapgpu] %zu descriptors written in %.3f ms\n", n, now_ms() - t_plan);
    if (info.bad == 1) return fail(SNAPGPU_EINVAL, "file %zu too large", info.bad_index);
    if (info.bad == 2)
        return fail(SNAPGPU_EINVAL, "non-final segment %zu is not a multiple of 128 bytes", info.bad_index);
    const bool aligned = info.aligned && ((uintptr_t)d_data & 15) == 0;
    uint64_t total_blocks = info.total_blocks, max_blocks = info.max_blocks;
    const DevicePlan plan = plan_layout(slot->d_buf, n);

    // The long-file bin: a handful of files far longer than the rest leave the batched kernel
    // (their descriptors are marked kSegSkip there) and go to sha512_long_kernel, which runs
    // beside it on its own stream.  With many long files the batched kernel keeps them: it has
    // the higher throughput, the long kernel only the shorter chain.
    SegDesc *h_descs = static_cast<SegDesc *>(slot->h_buf);
    const size_t n_long = select_long_bin(info, h_descs, n, D.sm_count, long_mode, &total_blocks, &max_blocks);
    const size_t n_main = n - n_long;

    // upload and binning run on the copy stream, i.e. beside whatever the caller's stream is
    // still hashing; the hashing kernel waits for the finished plan
    SG_CUDA(cudaMemcpyAsync(slot->d_buf, slot->h_buf, (n + n_long) * sizeof(SegDesc), cudaMemcpyHostToDevice,
                            D.copy_stream));
    SG_CUDA(cudaMemsetAsync(slot->d_counter, 0, kCounterBytes, D.copy_stream));
    if (n_main && (rc = enqueue_length_binning(D, D.copy_stream, plan, n, max_blocks))) return rc;
    SG_CUDA(cudaEventRecord(slot->uploaded, D.copy_stream));
    SG_CUDA(cudaStreamWaitEvent(stream, slot->uploaded, 0));

    // warps per SM sub-partition: more hides latency better, fewer shortens the makespan when
    // one file is a large share of a lane's work (see DESIGN.md "makespan").
    int per_sm = (int)R.opt.sha_warps_per_sm.load();
    if (per_sm <= 0) {
        const uint64_t lanes = (uint64_t)D.sm_count * kShaThreads;
        // depth = how many times the longest file fits into a lane's share of the launch.  Below 2
        // the longest chain is the makespan and must have its sub-partition to itself; three
        // warps only pay off for deep launches (measured over seven length distributions,
        // profiles/r01_shape_probe.jsonl: with depth 3..5 two warps are 1-10 % faster than three)
        const uint64_t depth = max_blocks ? total_blocks / (lanes * max_blocks) : 1;
        per_sm = depth < 2 ? 1 : depth < 6 ? 2 : 3;
    }
    per_sm = std::min(per_sm, kShaCtasPerSmMax);
    const u32 nunits = (u32)((n + 31) / 32);
    const u32 want_ctas = (nunits + kShaWarpsPerCta - 1) / kShaWarpsPerCta;
    const u32 grid = std::max<u32>(1, std::min<u32>((u32)(D.sm_count * per_sm), want_ctas));

    ShaKernel k = sha_kernel_for((int)R.opt.sha_variant.load(), aligned);
    TimedLaunch *tl = nullptr;
    if (R.opt.time_kernels.load()) {
        harvest_timings(D.sha_t, 8, D.sha_ms_sum, D.sha_ms_n, D.sha_ms_last, false);
        tl = &D.sha_t[D.sha_ti];
        D.sha_ti = (D.sha_ti + 1) % 8;
        if (tl->pending) harvest_timings(tl, 1, D.sha_ms_sum, D.sha_ms_n, D.sha_ms_last, true);
        tl->plan_ready = slot->uploaded;
        const TimedLaunch &before = D.sha_t[(D.sha_ti + 6) % 8];      // the launch before this one
        tl->prev_end = before.recorded ? before.end : nullptr;
        SG_CUDA(cudaEventRecord(tl->beg, stream));
    }
    if (n_long) {
        // everything `stream` has been asked to do so far (the data, a chaining value) comes first
        SG_CUDA(cudaEventRecord(slot->fork, stream));
        SG_CUDA(cudaStreamWaitEvent(long_stream, slot->fork, 0));
        if (pair_bin) {                                     // one chain per lane pair (sha512_pair.cuh)
            const u32 per_cta = pair_cta_shape(n_long, D.sm_count, (u32)R.opt.pair_files_per_cta.load());
            const u32 long_grid = (u32)((n_long + per_cta - 1) / per_cta);
            pair_kernel_for(aligned, (int)R.opt.pair_form.load(), per_cta <= 2 ? 1 : 2)
                <<<long_grid, kLongThreads, kPairSmemBytes, long_stream>>>(d_data, plan.long_descs, (u32)n_long, d_digests,
                                                                           per_cta);
        } else {                                            // one chain per lane (sha512_long.cuh)
            const u32 long_grid = (u32)((n_long + kLongFilesPerCta - 1) / kLongFilesPerCta);
            if (aligned)
                sha512_long_kernel<true><<<long_grid, kLongThreads, kLongSmemBytes, long_stream>>>(
                    d_data, plan.long_descs, (u32)n_long, d_digests);
            else
                sha512_long_kernel<false><<<long_grid, kLongThreads, kLongSmemBytes, long_stream>>>(
                    d_data, plan.long_descs, (u32)n_long, d_digests);
        }
        SG_CUDA(cudaGetLastError());
        SG_CUDA(cudaEventRecord(slot->join, long_stream));
        R.kernel_launches++;
        R.sha_long_launches++;
    }
    if (n_main) {
        // several CTAs per SM, plenty of units and a wide spread of lengths (longest file at least three times the
        // mean): claims from both ends of the sorted plan (sha512_kernels.cuh).  With a narrow spread there is nothing
        // to balance, and a warp that times itself as slow by mistake only disturbs the order: uniform 8-16 KiB files
        // read 0.94-0.98 with it from run to run, 0.981 without (profiles/r02_shape_probe.jsonl).
        const bool wide = max_blocks * (uint64_t)n_main >= 3 * total_blocks;
        const long long two_ended = R.opt.two_ended.load();
        const u32 first_wave = ((two_ended == 2 || (two_ended == 1 && wide)) && grid > (u32)D.sm_count &&
                                nunits >= 8 * grid * kShaWarpsPerCta)
                                   ? (u32)D.sm_count : 0u;
        k<<<grid, kShaThreads, 0, stream>>>(d_data, plan.descs, plan.order, (u32)n, d_digests, slot->d_counter, 1u, first_wave);
        SG_CUDA(cudaGetLastError());
        R.kernel_launches++;
    }
    if (n_long) SG_CUDA(cudaStreamWaitEvent(stream, slot->join, 0));
    if (tl) {
        SG_CUDA(cudaEventRecord(tl->end, stream));
        tl->pending = true;
        tl->recorded = true;
    }
    SG_CUDA(cudaEventRecord(slot->done, stream));
    slot->in_flight = true;
    R.sha_launches++;
    if (trace_on()) fprintf(stderr, "[snapgpu] sha512 launch of %zu files enqueued in %.3f ms (host)\n", n, now_ms() - t_plan);
    return 0;
}

// ------------------------------------------------------------------------------------------
// cmp launch
// ------------------------------------------------------------------------------------------

struct CmpItem {
    uint64_t off, len;
};

static int launch_cmp(Pipe &D, cudaStream_t stream, const uint8_t *d_a, const uint8_t *d_b, const CmpItem *items,
                      size_t n, uint8_t *d_equal) {
    if (n == 0) return 0;
    if (n > 0xfffffff0ull) return fail(SNAPGPU_EINVAL, "too many pairs in one launch (%zu)", n);
    auto &R = rt();
    PlanSlot *slot;
    int rc = acquire_slot(D, (n + 1) * sizeof(CmpPair), &slot);
    if (rc) return rc;
    CmpPair *hp = static_cast<CmpPair *>(slot->h_buf);
    bool aligned = (((uintptr_t)d_a | (uintptr_t)d_b) & 15) == 0;
    uint64_t tiles = 0;
    for (size_t i = 0; i < n; i++) {
        hp[i].off = items[i].off;
        hp[i].len = items[i].len;
        hp[i].first_tile = tiles;
        tiles += (items[i].len + kCmpTileBytes - 1) / kCmpTileBytes;
        aligned = aligned && ((items[i].off & 15) == 0);
    }
    hp[n].off = 0;
    hp[n].len = 0;
    hp[n].first_tile = tiles;
    SG_CUDA(cudaMemsetAsync(d_equal, 1, n, stream));
    if (tiles == 0) return 0;
    SG_CUDA(cudaMemcpyAsync(slot->d_buf, slot->h_buf, (n + 1) * sizeof(CmpPair), cudaMemcpyHostToDevice, stream));
    // grid = a whole number of full waves: the CTAs that fit one SM (5 at 48 registers) times 4.  A
    // grid that is not a multiple of the resident count leaves a partial last wave (8/SM measured 4 %
    // slower than 5, 10, 15 or 20/SM).
    int per_sm = (int)R.opt.cmp_ctas_per_sm.load();
    if (per_sm <= 0) {
        int resident = 0;
        cudaError_t oe = aligned ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, cmp_pairs_kernel<true>, kCmpThreads, 0)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, cmp_pairs_kernel<false>, kCmpThreads, 0);
        if (oe != cudaSuccess || resident <= 0) {
            cudaGetLastError();
            resident = 5;
        }
        per_sm = 4 * resident;
    }
    const uint64_t grid64 = std::min<uint64_t>((uint64_t)D.sm_count * per_sm, tiles);
    const u32 grid = (u32)std::max<uint64_t>(1, grid64);
    TimedLaunch *tl = nullptr;
    if (R.opt.time_kernels.load()) {
        harvest_timings(D.cmp_t, 8, D.cmp_ms_sum, D.cmp_ms_n, D.cmp_ms_last, false);
        tl = &D.cmp_t[D.cmp_ti];
        D.cmp_ti = (D.cmp_ti + 1) % 8;
        if (tl->pending) harvest_timings(tl, 1, D.cmp_ms_sum, D.cmp_ms_n, D.cmp_ms_last, true);
        SG_CUDA(cudaEventRecord(tl->beg, stream));
    }
    if (aligned)
        cmp_pairs_kernel<true><<<grid, kCmpThreads, 0, stream>>>(d_a, d_b, static_cast<const CmpPair *>(slot->d_buf),
                                                                 (u32)n, tiles, d_equal);
    else
        cmp_pairs_kernel<false><<<grid, kCmpThreads, 0, stream>>>(d_a, d_b, static_cast<const CmpPair *>(slot->d_buf),
                                                                  (u32)n, tiles, d_equal);
    SG_CUDA(cudaGetLastError());
    if (tl) {
        SG_CUDA(cudaEventRecord(tl->end, stream));
        tl->pending = true;
    }
    SG_CUDA(cudaEventRecord(slot->done, stream));
    slot->in_flight = true;
    R.kernel_launches++;
    R.cmp_launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------
// host-buffer pipeline: pack spans -> H2D -> kernel -> D2H, double buffered
// ------------------------------------------------------------------------------------------

// At least `nbuf` staging buffers of stage_bytes and result buffers of out_bytes.  Growing frees
// the old buffers: the caller makes sure nothing in flight still uses them.
static int ensure_staging(Pipe &D, size_t stage_bytes, size_t out_bytes, int nbuf = 2) {
    if (D.stage_cap < stage_bytes) {
        for (int b = 0; b < kStageBufs; b++) {
            if (D.d_stage[b]) cudaFree(D.d_stage[b]);
            D.d_stage[b] = nullptr;
        }
        D.stage_cap = 0;
        D.stage_n = 0;
        D.stage_cap = stage_bytes;
    }
    for (; D.stage_n < nbuf; D.stage_n++) SG_CUDA(cudaMalloc(&D.d_stage[D.stage_n], D.stage_cap + kStageSlack));
    if (D.out_cap < out_bytes) {
        for (int b = 0; b < kStageBufs; b++) {
            if (D.d_out[b]) cudaFree(D.d_out[b]);
            if (D.h_out[b]) host_pinned_free(D.h_out[b]);
            D.d_out[b] = D.h_out[b] = nullptr;
        }
        D.out_n = 0;
        D.out_cap = std::max<size_t>(out_bytes + out_bytes / 2, 1u << 16);
    }
    for (; D.out_n < nbuf; D.out_n++) {
        SG_CUDA(cudaMalloc(&D.d_out[D.out_n], D.out_cap));
        SG_CUDA(host_pinned_alloc(reinterpret_cast<void **>(&D.h_out[D.out_n]), D.out_cap));
    }
    return 0;
}

// Is `p` memory the copy engine can read directly (cudaHostAlloc / cudaHostRegister)?
static bool host_pointer_is_pinned(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

// Host-to-device copy of one span on D.copy_stream (in stream order with what was enqueued before).
// Pinned memory is one DMA.  Pageable memory -- a Go slice handed straight through cgo -- would make
// cudaMemcpyAsync stage through the driver's single bounce buffer at ~10 GB/s; instead kFeeders
// host threads copy 4 MiB pieces into pinned bounce buffers of their own (two each, so the memcpy
// of a piece overlaps the DMA of the previous one) and enqueue the DMAs on their own streams.
static int h2d_span(Pipe &D, uint8_t *dst, const uint8_t *src, size_t bytes, bool pinned) {
    if (bytes == 0) return 0;
    if (pinned || bytes < kBounceBytes) {
        SG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, D.copy_stream));
        return 0;
    }
    const size_t pieces = (bytes + kBounceBytes - 1) / kBounceBytes;
    // threads: one memcpy runs at ~10 GB/s, the link takes ~55; half the cores, shared between the
    // bound devices (each has its own set of feeders when a call is sharded over several)
    int nfeed = (int)rt().opt.feeders.load();
    if (nfeed <= 0) {
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        nfeed = (int)(hw / 2 / std::max<size_t>(1, rt().devs.size()));
    }
    nfeed = std::max(1, std::min<int>({nfeed, kFeeders, (int)pieces}));
    for (int f = 0; f < nfeed; f++) {
        if (D.feeder_stream[f]) continue;
        SG_CUDA(cudaStreamCreateWithFlags(&D.feeder_stream[f], cudaStreamNonBlocking));
        SG_CUDA(cudaEventCreateWithFlags(&D.feeder_done[f], cudaEventDisableTiming));
        for (int k = 0; k < 2; k++) {
            SG_CUDA(host_pinned_alloc(reinterpret_cast<void **>(&D.bounce[f][k]), kBounceBytes));
            SG_CUDA(cudaEventCreateWithFlags(&D.bounce_free[f][k], cudaEventDisableTiming));
        }
    }
    // the DMAs must not overtake what the copy stream was asked to do before (the previous chunk
    // may still be reading the same staging buffer's neighbour; plan uploads): fork from it
    if (!D.feeder_fork) SG_CUDA(cudaEventCreateWithFlags(&D.feeder_fork, cudaEventDisableTiming));
    SG_CUDA(cudaEventRecord(D.feeder_fork, D.copy_stream));
    for (int f = 0; f < nfeed; f++) SG_CUDA(cudaStreamWaitEvent(D.feeder_stream[f], D.feeder_fork, 0));
    int rcs[kFeeders] = {};
    std::string errs[kFeeders];
    auto feed = [&](int f) {
        if (cudaSetDevice(D.ordinal) != cudaSuccess) { rcs[f] = SNAPGPU_ECUDA; errs[f] = "cudaSetDevice failed"; return; }
        int use = 0;
        for (size_t p = (size_t)f; p < pieces; p += (size_t)nfeed, use ^= 1) {
            const size_t off = p * kBounceBytes, len = std::min(kBounceBytes, bytes - off);
            cudaError_t e = cudaEventSynchronize(D.bounce_free[f][use]);       // its previous DMA has drained
            if (e == cudaSuccess) {
                memcpy(D.bounce[f][use], src + off, len);
                e = cudaMemcpyAsync(dst + off, D.bounce[f][use], len, cudaMemcpyHostToDevice, D.feeder_stream[f]);
            }
            if (e == cudaSuccess) e = cudaEventRecord(D.bounce_free[f][use], D.feeder_stream[f]);
            if (e != cudaSuccess) {
                rcs[f] = SNAPGPU_ECUDA;
                errs[f] = cudaGetErrorString(e);
                return;
            }
        }
    };
    std::vector<std::thread> th;
    for (int f = 1; f < nfeed; f++) th.emplace_back(feed, f);
    feed(0);
    for (auto &t : th) t.join();
    for (int f = 0; f < nfeed; f++)
        if (rcs[f]) return fail(rcs[f], "host-to-device copy through the bounce buffers failed: %s", errs[f].c_str());
    for (int f = 0; f < nfeed; f++) {                         // join: the copy stream continues after all pieces
        SG_CUDA(cudaEventRecord(D.feeder_done[f], D.feeder_stream[f]));
        SG_CUDA(cudaStreamWaitEvent(D.copy_stream, D.feeder_done[f], 0));
    }
    return 0;
}

// A unit of pipelined work: a run of items whose bytes form one dense span of the host buffer.
struct WorkItem {
    size_t user_index;     // digest slot / pair index in the caller's arrays
    uint64_t off, len;     // host offsets
    uint64_t prefix;
    uint32_t flags;
};
struct Chunk {
    size_t first, count;   // range in the shard's item list, or in the split-piece list
    bool in_pieces;        // true: `first` indexes the pieces of an item that was split
    uint64_t span_begin, span_end;   // host byte range to copy (span_begin is 16-aligned)
    bool needs_state_in;
    bool dense_out;        // user_index runs first..first+count-1: results move with one memcpy
};

// The items of one shard: an explicit list, or -- the common call, plain files on one device --
// the caller's offsets/lengths arrays used in place (no per-file copy in front of the first DMA).
struct ItemList {
    const WorkItem *items = nullptr;
    const uint64_t *offs = nullptr, *lens = nullptr;
    size_t n = 0;
    ItemList() {}
    ItemList(const std::vector<WorkItem> &v) : items(v.data()), n(v.size()) {}
    ItemList(const uint64_t *o, const uint64_t *l, size_t count) : offs(o), lens(l), n(count) {}
    WorkItem at(size_t k) const { return items ? items[k] : WorkItem{k, offs[k], lens[k], 0, 0}; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
};

struct ChunkPlan {
    std::vector<Chunk> chunks;
    std::vector<WorkItem> pieces;   // continuation segments (SHA) / sub-ranges (cmp) of oversized items
    ItemList base;
    WorkItem item(const Chunk &c, size_t i) const { return c.in_pieces ? pieces[c.first + i] : base.at(c.first + i); }
};

// Cuts a shard's item list into chunks that fit the staging buffer, one chunk per next() call,
// so that the first host-to-device copy starts after a few thousand items have been looked at
// instead of after the whole list.  Items longer than the staging buffer are split into
// continuation segments (SHA) or sub-ranges (cmp); items that fit are referenced in place.
struct ChunkStream {
    const ItemList in;
    const bool is_sha;
    ChunkPlan &plan;
    size_t k = 0;             // next item of `in`
    uint64_t piece_done = 0;  // bytes of in[k] already emitted as pieces (only while splitting)
    bool splitting = false;

    ChunkStream(const ItemList &items, bool sha, ChunkPlan &p) : in(items), is_sha(sha), plan(p) {
        plan.chunks.clear();
        plan.pieces.clear();
        plan.base = in;
    }

    static uint64_t usable_of(size_t cap) { return (cap - 64) & ~(uint64_t)127; }

    // cap_now: bytes this chunk should stay under (the pipeline starts with small chunks);
    // cap_max: the staging buffer.  An item that does not fit cap_now but fits cap_max gets a
    // cap_max chunk; an item that does not fit cap_max is split.
    bool next(size_t cap_now, size_t cap_max, Chunk *out) {
        const uint64_t usable_max = usable_of(cap_max);
        if (!splitting) {
            if (k >= in.size()) return false;
            if (in.at(k).len + 16 > usable_max) {
                splitting = true;
                piece_done = 0;
            }
        }
        if (splitting) {
            const WorkItem w = in.at(k);
            uint64_t piece = std::min<uint64_t>(w.len - piece_done, usable_max - 128);
            piece = (piece_done + piece < w.len) ? (piece & ~(uint64_t)127) : piece;
            WorkItem s = w;
            s.off = w.off + piece_done;
            s.len = piece;
            s.prefix = w.prefix + piece_done;
            if (is_sha) {
                s.flags = w.flags;
                if (piece_done > 0) s.flags |= kSegContinue;
                if (piece_done + piece < w.len) s.flags |= kSegNoFinal;
            }
            plan.pieces.push_back(s);
            *out = Chunk{plan.pieces.size() - 1, 1, true, s.off & ~(uint64_t)15, s.off + s.len,
                         (s.flags & kSegContinue) != 0, true};
            piece_done += piece;
            if (piece_done >= w.len) {
                splitting = false;
                k++;
            }
            plan.chunks.push_back(*out);
            return true;
        }
        const WorkItem head = in.at(k);
        const uint64_t usable = head.len + 16 > usable_of(cap_now) ? usable_max : usable_of(cap_now);
        Chunk cur{k, 0, false, head.off & ~(uint64_t)15, head.off + head.len, false, true};
        size_t prev_user = 0;
        for (; k < in.size(); k++) {
            const WorkItem w = in.at(k);
            const uint64_t begin = w.off & ~(uint64_t)15;
            const uint64_t end = w.off + w.len;
            if (cur.count) {
                const bool fits = w.len + 16 <= usable_max && cur.count < kMaxChunkItems && begin >= cur.span_begin &&
                                  end - cur.span_begin <= usable && w.off <= cur.span_end + (1u << 20);
                if (!fits) break;
                if (w.user_index != prev_user + 1) cur.dense_out = false;
            }
            prev_user = w.user_index;
            cur.span_end = std::max(cur.span_end, end);
            cur.count++;
            cur.needs_state_in = cur.needs_state_in || (is_sha && (w.flags & kSegContinue));
        }
        *out = cur;
        plan.chunks.push_back(cur);
        return true;
    }
};

// All chunks of a list for one fixed staging size (test hook).
static void build_chunks(const ItemList &in, size_t cap, bool is_sha, ChunkPlan &plan) {
    ChunkStream cs(in, is_sha, plan);
    Chunk c;
    while (cs.next(cap, cap, &c)) {}
}

// The host pipeline ramps its chunk size: a small first chunk gets the copy engine going
// while the rest of the list is still being cut, later chunks are large because every launch
// costs at least the serial chain of its longest file (~2 ms for 64 KiB).
static size_t ramp_cap(size_t chunk_index, size_t cap_max) {
    const size_t first = 64u << 20;
    const size_t want = chunk_index >= 8 ? cap_max : (first << (2 * chunk_index));
    return std::min(want, cap_max);
}

// ... and tapers it: nothing overlaps the hashing of the LAST chunk, and a launch costs at least the serial
// chain of its longest file however few files it has (with 64 KiB files: ~2 ms in the batched kernel, ~1 ms
// once the long-file bin takes them, which it does for chunks up to ~256 MiB).  So the rest of a shard does
// not go out as one large chunk.  The last one is 64 MiB: small enough for the bin and a short copy-back of
// digests.  The one before it is 128 MiB: its launch (1.15 ms) is over before the last chunk has been copied
// (1.2 ms) -- the bin's CTAs want SMs to themselves and would otherwise wait for the batched kernel of a
// large chunk to leave them (measured: 1.6 ms instead of 1.0 for the last launch) -- and its own copy (2.4 ms)
// covers the ~2 ms launch of the large chunk before it.  Chunks alternate between two compute streams, so a
// launch never queues behind the previous one either.
constexpr uint64_t kTailLast = 64u << 20, kTailBefore = 128u << 20;
static size_t taper_cap(size_t want, const ItemList &shard, size_t next_item) {
    if (next_item >= shard.size() || !rt().opt.taper.load()) return want;
    const WorkItem last = shard.at(shard.size() - 1);
    const uint64_t pos = shard.at(next_item).off, end = last.off + last.len;
    if (end <= pos) return want;                          // not laid out in order: no estimate
    const uint64_t left = end - pos;
    if (left <= kTailLast + kTailLast / 2) return want;                                   // the last chunk
    if (left <= kTailLast + kTailBefore + kTailBefore / 2) return (size_t)std::min<uint64_t>(want, left - kTailLast);
    const uint64_t body = left - (kTailLast + kTailBefore);
    return body <= want ? (size_t)body : want;
}

// Runs one device's shard of a host-buffer SHA-512 batch.  digests: caller's n*64 array.
// The staging buffers a shard needs: the configured size, or less when the whole shard is smaller
// (a concurrent caller hashing a few MiB on a second pipe should not allocate two 1 GiB buffers).
static size_t staging_for(const ItemList &shard, size_t cap_max) {
    if (shard.size() > (1u << 16)) return cap_max;
    uint64_t lo = ~0ull, hi = 0, sum = 0;
    for (size_t k = 0; k < shard.size(); k++) {
        const WorkItem w = shard.at(k);
        lo = std::min(lo, w.off);
        hi = std::max(hi, w.off + w.len);
        sum += w.len + 32;
    }
    const uint64_t need = std::min<uint64_t>(hi - lo, sum) + 4096;     // dense spans never exceed either
    size_t cap = (size_t)16 << 20;
    while (cap < need && cap < cap_max) cap <<= 1;
    return std::min(cap, cap_max);
}

static int sha512_shard(Device &dev, const uint8_t *data, const ItemList &shard, uint8_t *digests) {
    if (shard.empty()) return 0;
    PipeLease lease(dev);
    Pipe &D = *lease.pipe;
    SG_CUDA(cudaSetDevice(D.ordinal));
    auto &R = rt();
    const double t_begin = now_ms();
    const size_t cap = std::min(staging_bytes(), std::max(D.stage_cap, staging_for(shard, staging_bytes())));
    ChunkPlan plan;
    ChunkStream stream(shard, true, plan);
    const std::vector<Chunk> &chunks = plan.chunks;      // grows as the stream is consumed
    int rc = ensure_staging(D, cap, std::min(shard.size(), kMaxChunkItems) * 64);
    if (rc) return rc;

    const size_t phase = (uintptr_t)data & 15;   // keep (data + off) mod 16 on the device
    const bool pinned = host_pointer_is_pinned(data);
    int scatter_pending[2] = {-1, -1};            // chunk index whose digests wait in h_out[b]
    auto scatter = [&](int b) -> int {
        if (scatter_pending[b] < 0) return 0;
        SG_CUDA(cudaEventSynchronize(D.ev_done[b]));
        const Chunk c = chunks[(size_t)scatter_pending[b]];
        if (c.dense_out) {
            memcpy(digests + 64 * plan.item(c, 0).user_index, D.h_out[b], 64 * c.count);
        } else {
            for (size_t i = 0; i < c.count; i++)
                memcpy(digests + 64 * plan.item(c, i).user_index, D.h_out[b] + 64 * i, 64);
        }
        scatter_pending[b] = -1;
        return 0;
    };

    // trace only: when each chunk's copy began and ended on the copy stream, relative to the first
    std::vector<cudaEvent_t> tr;
    auto mark = [&]() {
        if (!trace_on()) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, D.copy_stream);
        tr.push_back(e);
    };
    Chunk c;
    for (size_t ci = 0; stream.next(taper_cap(ramp_cap(ci, cap), shard, stream.k), cap, &c); ci++) {
        const int b = (int)(ci & 1);
        // one compute stream per staging buffer: launches of small chunks are bound by the chain of their longest
        // file, not by throughput, so the last chunk's should not queue behind the one before it
        const cudaStream_t cs = D.slot_stream[b];
        const double t_chunk = now_ms();
        if ((rc = scatter(b))) return rc;        // buffer b (stage, out) is free again
        const double t_free = now_ms();
        const size_t span = (size_t)(c.span_end - c.span_begin);
        mark();
        if (span) {
            if ((rc = h2d_span(D, D.d_stage[b] + phase, data + c.span_begin, span, pinned))) return rc;
            R.h2d_bytes += span;
        }
        mark();
        SG_CUDA(cudaEventRecord(D.ev_copied[b], D.copy_stream));
        if (c.needs_state_in) {
            // chaining values come from the previous chunk: finish it first
            if ((rc = scatter(b ^ 1))) return rc;
            for (size_t i = 0; i < c.count; i++)
                memcpy(D.h_out[b] + 64 * i, digests + 64 * plan.item(c, i).user_index, 64);
            SG_CUDA(cudaMemcpyAsync(D.d_out[b], D.h_out[b], c.count * 64, cudaMemcpyHostToDevice, cs));
            R.h2d_bytes += c.count * 64;
        }
        SG_CUDA(cudaStreamWaitEvent(cs, D.ev_copied[b], 0));
        {
            const uint64_t rebase = phase - c.span_begin;       // host offset -> staging offset
            auto get = [&plan, &c, rebase](size_t i) {
                const WorkItem w = plan.item(c, i);
                return SegDesc{w.off + rebase, w.len, w.prefix, (u32)i, w.flags};
            };
            if ((rc = launch_sha512(D, cs, D.d_stage[b], get, c.count, D.d_out[b], D.slot_long_stream[b]))) return rc;
        }
        SG_CUDA(cudaMemcpyAsync(D.h_out[b], D.d_out[b], c.count * 64, cudaMemcpyDeviceToHost, cs));
        R.d2h_bytes += c.count * 64;
        SG_CUDA(cudaEventRecord(D.ev_done[b], cs));
        scatter_pending[b] = (int)ci;   // scatter(b) syncs on ev_done[b] before buffer b is reused
        if (trace_on())
            fprintf(stderr, "[snapgpu] chunk %zu: %zu items, span %.1f MiB, waited %.2f ms for the buffer, enqueue %.2f ms\n",
                    ci, c.count, (double)(c.span_end - c.span_begin) / (1 << 20), t_free - t_chunk, now_ms() - t_free);
    }
    if ((rc = scatter(0))) return rc;
    if ((rc = scatter(1))) return rc;
    if (trace_on()) {
        const double t_end = now_ms();
        std::string line;
        for (size_t i = 0; i + 1 < tr.size(); i += 2) {
            float t0 = 0, t1 = 0;
            cudaEventElapsedTime(&t0, tr[0], tr[i]);
            cudaEventElapsedTime(&t1, tr[0], tr[i + 1]);
            char buf[64];
            snprintf(buf, sizeof buf, " [%.2f-%.2f]", t0, t1);
            line += buf;
        }
        for (cudaEvent_t e : tr) cudaEventDestroy(e);
        fprintf(stderr, "[snapgpu] dev %d: shard done in %.2f ms; copies on the copy stream (ms from the first):%s\n", D.ordinal,
                t_end - t_begin, line.c_str());
    }
    return 0;
}

static int cmp_shard(Device &dev, const uint8_t *a, const uint8_t *b_host, const ItemList &shard,
                     uint8_t *equal) {
    if (shard.empty()) return 0;
    PipeLease lease(dev);
    Pipe &D = *lease.pipe;
    SG_CUDA(cudaSetDevice(D.ordinal));
    auto &R = rt();
    const size_t cap = std::min(staging_bytes(), std::max(D.stage_cap, 2 * staging_for(shard, staging_bytes())));
    const size_t half = (cap / 2) & ~(size_t)255;
    ChunkPlan plan;
    ChunkStream stream(shard, false, plan);
    const std::vector<Chunk> &chunks = plan.chunks;
    int rc = ensure_staging(D, cap, std::min(shard.size(), kMaxChunkItems));
    if (rc) return rc;
    // the two streams may sit at different phases mod 16; the kernel then takes the byte path
    const size_t pa = (uintptr_t)a & 15, pb = (uintptr_t)b_host & 15;
    const bool a_pinned = host_pointer_is_pinned(a), b_pinned = host_pointer_is_pinned(b_host);
    std::vector<CmpItem> ci_items;
    int pending[2] = {-1, -1};
    auto gather = [&](int b) -> int {
        if (pending[b] < 0) return 0;
        SG_CUDA(cudaEventSynchronize(D.ev_done[b]));
        const Chunk c = chunks[(size_t)pending[b]];
        for (size_t i = 0; i < c.count; i++)
            if (D.h_out[b][i] == 0) equal[plan.item(c, i).user_index] = 0;
        pending[b] = -1;
        return 0;
    };
    Chunk c;
    for (size_t ci = 0; stream.next(ramp_cap(ci, half), half, &c); ci++) {
        const int b = (int)(ci & 1);
        if ((rc = gather(b))) return rc;
        // host-side early-out: a pair already known to differ is not copied again
        if (c.count == 1 && equal[plan.item(c, 0).user_index] == 0) continue;
        const size_t span = (size_t)(c.span_end - c.span_begin);
        uint8_t *da = D.d_stage[b] + pa, *db = D.d_stage[b] + half + 128 + pb;
        if (span) {
            if ((rc = h2d_span(D, da, a + c.span_begin, span, a_pinned))) return rc;
            if ((rc = h2d_span(D, db, b_host + c.span_begin, span, b_pinned))) return rc;
            R.h2d_bytes += 2 * span;
        }
        SG_CUDA(cudaEventRecord(D.ev_copied[b], D.copy_stream));
        ci_items.resize(c.count);
        for (size_t i = 0; i < c.count; i++) {
            ci_items[i].off = plan.item(c, i).off - c.span_begin;
            ci_items[i].len = plan.item(c, i).len;
        }
        SG_CUDA(cudaStreamWaitEvent(D.compute_stream, D.ev_copied[b], 0));
        // offsets are relative to da/db, whose phases differ only if the host pointers' do
        if ((rc = launch_cmp(D, D.compute_stream, da, db, ci_items.data(), c.count, D.d_out[b]))) return rc;
        SG_CUDA(cudaMemcpyAsync(D.h_out[b], D.d_out[b], c.count, cudaMemcpyDeviceToHost, D.compute_stream));
        R.d2h_bytes += c.count;
        SG_CUDA(cudaEventRecord(D.ev_done[b], D.compute_stream));
        pending[b] = (int)ci;
    }
    if ((rc = gather(0))) return rc;
    if ((rc = gather(1))) return rc;
    return 0;
}

// ------------------------------------------------------------------------------------------
// batch session (runtime.hpp): the same pipeline as sha512_shard, fed batch by batch
// ------------------------------------------------------------------------------------------

}  // namespace snapgpu

class snapgpu::BatchSession {
public:
    struct Slot {
        bool busy = false, copy_reported = false;
        uint64_t ticket = 0;
        std::vector<uint8_t *> dst;
    };
    struct Lane {
        Device *dev = nullptr;
        std::unique_ptr<PipeLease> lease;
        Slot slot[kStageBufs];
        int busy() const {
            int n = 0;
            for (const Slot &x : slot) n += x.busy;
            return n;
        }
    };
    DevsInUse use;
    std::vector<Lane> lanes;
    size_t max_batch_bytes = 0;
    uint64_t next_ticket = 1;
    size_t in_flight = 0;
    std::vector<uint64_t> span_base;
    double t_open = 0;

    BatchSession() : use(rt().devs_mu) {}
};

namespace snapgpu {

// Finish what can be finished on one slot.  wait: block for it.
static int session_retire(BatchSession *s, BatchSession::Lane &L, int b, std::vector<uint64_t> *copied, bool wait) {
    BatchSession::Slot &S = L.slot[b];
    if (!S.busy) return 0;
    Pipe &P = *L.lease->pipe;
    if (!S.copy_reported) {
        cudaError_t e = wait ? cudaEventSynchronize(P.ev_copied[b]) : cudaEventQuery(P.ev_copied[b]);
        if (e == cudaErrorNotReady) return 0;
        if (e != cudaSuccess) return fail(SNAPGPU_ECUDA, "host-to-device copy failed: %s", cudaGetErrorString(e));
        S.copy_reported = true;
        if (copied) copied->push_back(S.ticket);
        if (trace_on()) fprintf(stderr, "[snapgpu] session: batch %llu copied, seen at %.2f ms\n", (unsigned long long)S.ticket, now_ms() - s->t_open);
    }
    cudaError_t e = wait ? cudaEventSynchronize(P.ev_done[b]) : cudaEventQuery(P.ev_done[b]);
    if (e == cudaErrorNotReady) return 0;
    if (e != cudaSuccess) return fail(SNAPGPU_ECUDA, "SHA-512 batch failed: %s", cudaGetErrorString(e));
    const uint8_t *src = P.h_out[b];
    for (size_t i = 0; i < S.dst.size(); i++) memcpy(S.dst[i], src + 64 * i, 64);
    S.busy = false;
    s->in_flight--;
    if (trace_on()) {
        float t[4] = {0, 0, 0, 0};
        if (P.tr_base) {
            cudaEventSynchronize(P.tr_ev[b][3]);
            for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], P.tr_base, P.tr_ev[b][k]);
        }
        const double o = P.tr_base_host;
        fprintf(stderr, "[snapgpu] session: batch %llu done, seen at %.2f ms; on the device: copy %.2f-%.2f, kernels end %.2f, digests back %.2f\n",
                (unsigned long long)S.ticket, now_ms() - s->t_open, t[0] + o, t[1] + o, t[2] + o, t[3] + o);
    }
    return 0;
}

int session_open(BatchSession **out, size_t max_batch_bytes) {
    *out = nullptr;
    std::unique_ptr<BatchSession> s(new BatchSession());
    auto &R = rt();
    if (R.devs.empty()) return fail(SNAPGPU_ENOINIT, "snapgpu_init has not been called (or failed)");
    s->max_batch_bytes = std::max<size_t>(max_batch_bytes, 1u << 20);
    s->t_open = now_ms();
    s->lanes.resize(R.devs.size());
    for (size_t d = 0; d < R.devs.size(); d++) {
        s->lanes[d].dev = R.devs[d].get();
        s->lanes[d].lease.reset(new PipeLease(*R.devs[d], true));
    }
    *out = s.release();
    return 0;
}

void session_close(BatchSession *s) {
    if (!s) return;
    for (auto &L : s->lanes) {
        cudaSetDevice(L.dev->ordinal);
        for (int b = 0; b < kStageBufs; b++)
            if (L.slot[b].busy) {                    // abandoned after an error: let the GPU finish with the buffers
                cudaEventSynchronize(L.lease->pipe->ev_done[b]);
                L.slot[b].busy = false;
            }
    }
    delete s;
}

size_t session_in_flight(const BatchSession *s) { return s->in_flight; }
size_t session_capacity(const BatchSession *s) { return kStageBufs * s->lanes.size(); }

int session_poll(BatchSession *s, std::vector<uint64_t> *copied, bool wait_all) {
    for (auto &L : s->lanes) {
        if (!L.busy()) continue;
        SG_CUDA(cudaSetDevice(L.dev->ordinal));
        // a lane's batches in the order they were enqueued (tickets grow), so that "copied" is
        // reported in order
        int order[kStageBufs];
        for (int k = 0; k < kStageBufs; k++) order[k] = k;
        std::sort(order, order + kStageBufs, [&](int x, int y) { return L.slot[x].ticket < L.slot[y].ticket; });
        for (int k = 0; k < kStageBufs; k++) {
            int rc = session_retire(s, L, order[k], copied, wait_all);
            if (rc) return rc;
        }
    }
    return 0;
}

int session_submit(BatchSession *s, const HostSpan *spans, size_t nspans, const SpanSeg *segs,
                   uint8_t *const *digest_dst, size_t nsegs, uint64_t *ticket, std::vector<uint64_t> *copied) {
    if (nsegs == 0) return fail(SNAPGPU_EINVAL, "empty batch");
    auto &R = rt();
    // a lane with a free slot, the one with the fewest batches in flight; none: wait for the oldest batch
    BatchSession::Lane *lane = nullptr;
    int b = -1;
    for (int pass = 0; pass < 2 && !lane; pass++) {
        int best_busy = kStageBufs;
        for (auto &L : s->lanes) {
            const int busy = L.busy();
            if (busy < best_busy) {
                best_busy = busy;
                lane = &L;
            }
        }
        if (lane) break;
        BatchSession::Lane *oldest = nullptr;
        int ob = 0;
        for (auto &L : s->lanes)
            for (int k = 0; k < kStageBufs; k++)
                if (L.slot[k].busy && (!oldest || L.slot[k].ticket < oldest->slot[ob].ticket)) {
                    oldest = &L;
                    ob = k;
                }
        SG_CUDA(cudaSetDevice(oldest->dev->ordinal));
        int rc = session_retire(s, *oldest, ob, copied, true);
        if (rc) return rc;
    }
    Pipe &P = *lane->lease->pipe;
    SG_CUDA(cudaSetDevice(P.ordinal));
    for (b = 0; lane->slot[b].busy; b++) {}                  // a free staging buffer of the lane

    s->span_base.resize(nspans);
    size_t total = 0;
    for (size_t k = 0; k < nspans; k++) {
        s->span_base[k] = total;
        total += (spans[k].bytes + 255) & ~(size_t)255;
    }
    if (total > s->max_batch_bytes) return fail(SNAPGPU_EINVAL, "batch of %zu bytes exceeds the session's %zu", total, s->max_batch_bytes);
    if (P.stage_cap < total || P.out_cap < nsegs * 64 || P.stage_n < kStageBufs || P.out_n < kStageBufs) {
        // growing frees the old buffers: nothing of this lane may still be using them
        if (P.stage_cap < total || P.out_cap < nsegs * 64) {
            for (int k = 0; k < kStageBufs; k++) {
                int rc = session_retire(s, *lane, k, copied, true);
                if (rc) return rc;
            }
            b = 0;
        }
        // the session's batch size at once (growing step by step would pay cudaMalloc again and again)
        const size_t want = std::max(s->max_batch_bytes, total);
        const double tg = now_ms();
        int rc = ensure_staging(P, std::max(P.stage_cap, want), std::max(P.out_cap, std::max<size_t>(nsegs * 64, 1u << 20)),
                                kStageBufs);
        if (rc) return rc;
        if (trace_on())
            fprintf(stderr, "[snapgpu] session: staging grown to %d x %zu MiB (+ %zu KiB of digests) in %.2f ms\n", kStageBufs,
                    P.stage_cap >> 20, P.out_cap >> 10, now_ms() - tg);
    }
    const bool tr = trace_on();
    if (tr) {
        if (!P.tr_base) {
            SG_CUDA(cudaEventCreate(&P.tr_base));
            for (auto &row : P.tr_ev)
                for (auto &e : row) SG_CUDA(cudaEventCreate(&e));
        }
        if (s->in_flight == 0 && s->next_ticket == 1) {
            SG_CUDA(cudaEventRecord(P.tr_base, P.copy_stream));
            P.tr_base_host = now_ms() - s->t_open;
        }
        SG_CUDA(cudaEventRecord(P.tr_ev[b][0], P.copy_stream));
    }
    for (size_t k = 0; k < nspans; k++)
        if (spans[k].bytes)
            SG_CUDA(cudaMemcpyAsync(P.d_stage[b] + s->span_base[k], spans[k].ptr, spans[k].bytes, cudaMemcpyHostToDevice,
                                    P.copy_stream));
    R.h2d_bytes += total;
    SG_CUDA(cudaEventRecord(P.ev_copied[b], P.copy_stream));
    if (tr) SG_CUDA(cudaEventRecord(P.tr_ev[b][1], P.copy_stream));
    cudaStream_t cs = P.slot_stream[b];
    SG_CUDA(cudaStreamWaitEvent(cs, P.ev_copied[b], 0));
    const uint64_t *base = s->span_base.data();
    auto get = [segs, base](size_t i) { return SegDesc{base[segs[i].span] + segs[i].off, segs[i].len, 0, (u32)i, 0}; };
    int rc = launch_sha512(P, cs, P.d_stage[b], get, nsegs, P.d_out[b], P.slot_long_stream[b]);
    if (rc) return rc;
    if (tr) SG_CUDA(cudaEventRecord(P.tr_ev[b][2], cs));
    SG_CUDA(cudaMemcpyAsync(P.h_out[b], P.d_out[b], nsegs * 64, cudaMemcpyDeviceToHost, cs));
    R.d2h_bytes += nsegs * 64;
    SG_CUDA(cudaEventRecord(P.ev_done[b], cs));
    if (tr) SG_CUDA(cudaEventRecord(P.tr_ev[b][3], cs));
    BatchSession::Slot &S = lane->slot[b];
    S.busy = true;
    S.copy_reported = false;
    S.ticket = s->next_ticket++;
    S.dst.assign(digest_dst, digest_dst + nsegs);
    s->in_flight++;
    if (ticket) *ticket = S.ticket;
    return 0;
}

// ------------------------------------------------------------------------------------------
// sharding across devices: no collective, results gathered by index on the host
// ------------------------------------------------------------------------------------------

// Items heavier than 1/(4*ndev) of the job are placed longest-first on the least loaded
// device; the rest are dealt out as contiguous runs (keeps each device's H2D spans dense).
static void shard_items(const std::vector<WorkItem> &all, const std::vector<uint64_t> &weight, int ndev,
                        std::vector<std::vector<WorkItem>> &out) {
    out.assign((size_t)ndev, {});
    if (ndev == 1) {
        out[0] = all;
        return;
    }
    for (auto &o : out) o.reserve(all.size() / (size_t)ndev + 16);
    uint64_t total = 0;
    for (uint64_t w : weight) total += w;
    const uint64_t big = std::max<uint64_t>(total / (4 * (uint64_t)ndev), 1);
    std::vector<uint64_t> load((size_t)ndev, 0);
    std::vector<size_t> bigs;
    for (size_t i = 0; i < all.size(); i++)
        if (weight[i] > big) bigs.push_back(i);
    std::sort(bigs.begin(), bigs.end(), [&](size_t x, size_t y) { return weight[x] > weight[y]; });
    std::vector<char> placed(all.size(), 0);
    for (size_t i : bigs) {
        size_t d = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        out[d].push_back(all[i]);
        load[d] += weight[i];
        placed[i] = 1;
    }
    const uint64_t target = (total + ndev - 1) / ndev;
    size_t d = 0;
    for (size_t i = 0; i < all.size(); i++) {
        if (placed[i]) continue;
        while (d + 1 < (size_t)ndev && load[d] >= target) d++;
        out[d].push_back(all[i]);
        load[d] += weight[i];
    }
}

static int run_on_devices(const std::vector<std::vector<WorkItem>> &shards,
                          const std::function<int(Device &, const std::vector<WorkItem> &)> &fn) {
    auto &R = rt();
    const size_t ndev = R.devs.size();
    if (ndev == 1) return fn(*R.devs[0], shards[0]);
    std::vector<int> rcs(ndev, 0);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> threads;
    for (size_t d = 0; d < ndev; d++)
        threads.emplace_back([&, d]() {
            rcs[d] = fn(*R.devs[d], shards[d]);
            if (rcs[d]) errs[d] = g_last_error;
        });
    for (auto &t : threads) t.join();
    for (size_t d = 0; d < ndev; d++)
        if (rcs[d]) return fail(rcs[d], "device %zu: %s", d, errs[d].c_str());
    return 0;
}

// The common multi-device call -- plain files, none heavy enough to need placing on its own --
// needs no item copies at all: the caller's arrays are cut into ndev contiguous index ranges of
// equal weight and every device works on its range in place.  Returns false when an item weighs
// more than 1/(4 ndev) of the job (shard_items then places those first).
template <class Weight>
static bool split_contiguous(const uint64_t *lengths, size_t n, size_t ndev, Weight weight_of, std::vector<size_t> &cut) {
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += weight_of(lengths[i]);
    const uint64_t big = std::max<uint64_t>(total / (4 * (uint64_t)ndev), 1);
    cut.assign(ndev + 1, n);
    cut[0] = 0;
    uint64_t acc = 0;
    size_t d = 1;
    for (size_t i = 0; i < n; i++) {
        const uint64_t w = weight_of(lengths[i]);
        if (w > big) return false;
        while (d < ndev && acc >= total * d / ndev) cut[d++] = i;
        acc += w;
    }
    return true;
}

static int run_on_ranges(const std::vector<size_t> &cut, const std::function<int(Device &, size_t lo, size_t hi)> &fn) {
    auto &R = rt();
    const size_t ndev = R.devs.size();
    std::vector<int> rcs(ndev, 0);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> threads;
    for (size_t d = 0; d < ndev; d++)
        threads.emplace_back([&, d]() {
            if (cut[d] == cut[d + 1]) return;
            rcs[d] = fn(*R.devs[d], cut[d], cut[d + 1]);
            if (rcs[d]) errs[d] = g_last_error;
        });
    for (auto &t : threads) t.join();
    for (size_t d = 0; d < ndev; d++)
        if (rcs[d]) return fail(rcs[d], "device %zu: %s", d, errs[d].c_str());
    return 0;
}

static int sha512_host_items(const uint8_t *data, const std::vector<WorkItem> &all, uint8_t *digests);

int sha512_host_segments(const uint8_t *data, const HostSeg *segs, size_t n, uint8_t *digests) {
    DevsInUse use(rt().devs_mu);
    if (rt().devs.empty()) return fail(SNAPGPU_ENOINIT, "snapgpu_init has not been called (or failed)");
    if (n == 0) return 0;
    if (!segs || !digests || (!data && n)) return fail(SNAPGPU_EINVAL, "null argument");
    std::vector<WorkItem> all(n);
    for (size_t i = 0; i < n; i++) {
        if (segs[i].len >= kMaxSegBytes) return fail(SNAPGPU_EINVAL, "file %zu too large", i);
        u32 f = 0;
        if (segs[i].flags & kHostSegContinue) f |= kSegContinue;
        if (segs[i].flags & kHostSegNoFinal) f |= kSegNoFinal;
        if ((f & kSegNoFinal) && (segs[i].len & 127))
            return fail(SNAPGPU_EINVAL, "non-final segment %zu is not a multiple of 128 bytes", i);
        all[i] = WorkItem{i, segs[i].off, segs[i].len, segs[i].prefix, f};
    }
    return sha512_host_items(data, all, digests);
}

// `all` carries user_index = position; one device: the list is used in place.
static int sha512_host_items(const uint8_t *data, const std::vector<WorkItem> &all, uint8_t *digests) {
    auto &R = rt();
    if (R.devs.size() == 1) return sha512_shard(*R.devs[0], data, all, digests);
    std::vector<uint64_t> weight(all.size());
    for (size_t i = 0; i < all.size(); i++) weight[i] = seg_blocks(all[i].len, all[i].flags) + 1;
    std::vector<std::vector<WorkItem>> shards;
    shard_items(all, weight, (int)R.devs.size(), shards);
    return run_on_devices(shards, [&](Device &D, const std::vector<WorkItem> &s) {
        return sha512_shard(D, data, s, digests);
    });
}

}  // namespace snapgpu

// ==========================================================================================
// extern "C"
// ==========================================================================================

using namespace snapgpu;

extern "C" {

const char *snapgpu_last_error(void) { return g_last_error.c_str(); }
const char *snapgpu_version(void) { return "snapgpu 0.1 (sm_100a)"; }

int snapgpu_init(const int *devices, int ndev) {
    auto &R = rt();
    std::lock_guard<std::mutex> lock(R.mu);
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0)
        return fail(SNAPGPU_ECUDA, "no CUDA device: %s (libsnapgpu has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    std::vector<int> want;
    if (ndev <= 0) {
        for (int i = 0; i < visible; i++) want.push_back(i);
    } else {
        for (int i = 0; i < ndev; i++) want.push_back(devices ? devices[i] : i);
    }
    for (int o : want)
        if (o < 0 || o >= visible) return fail(SNAPGPU_EINVAL, "device ordinal %d not visible (%d devices)", o, visible);
    {
        DevsInUse use(R.devs_mu);
        bool same = R.devs.size() == want.size();
        for (size_t i = 0; same && i < want.size(); i++) same = R.devs[i]->ordinal == want[i];
        if (same) return 0;
    }
    // the new device list is built aside and published in one step; the exclusive lock waits for
    // the calls in flight on the old list before it is torn down
    std::vector<std::unique_ptr<Device>> fresh;
    std::unique_lock<std::shared_mutex> excl(R.devs_mu);
    for (auto &d : R.devs) destroy_device(*d);
    R.devs.clear();
    for (int o : want) {
        std::unique_ptr<Device> d(new Device());
        int rc = init_device(*d, o);
        if (rc) {
            destroy_device(*d);
            for (auto &x : fresh) destroy_device(*x);
            return rc;
        }
        fresh.push_back(std::move(d));
    }
    R.devs = std::move(fresh);
    return 0;
}

void snapgpu_shutdown(void) {
    auto &R = rt();
    std::lock_guard<std::mutex> lock(R.mu);
    std::unique_lock<std::shared_mutex> excl(R.devs_mu);
    for (auto &d : R.devs) destroy_device(*d);
    R.devs.clear();
}

int snapgpu_num_devices(void) {
    DevsInUse use(rt().devs_mu);
    return (int)rt().devs.size();
}

#ifdef SNAPGPU_TRACE_WARPS
// experimental build only: the per-warp records of the last SHA-512 launch on device `dev`
int snapgpu_test_warp_trace(int dev, unsigned long long *out, size_t nwarps) {
    DevsInUse use(rt().devs_mu);
    Device *D = nullptr;
    int rc = get_device(dev, &D);
    if (rc) return rc;
    SG_CUDA(cudaSetDevice(D->ordinal));
    SG_CUDA(cudaDeviceSynchronize());
    SG_CUDA(cudaMemcpyFromSymbol(out, g_warp_trace, std::min<size_t>(nwarps, 8192) * 4 * sizeof(unsigned long long)));
    return 0;
}
#endif

int snapgpu_set_option(const char *key, long long value) {
    if (!key) return fail(SNAPGPU_EINVAL, "null key");
    auto &o = rt().opt;
    std::string k(key);
    if (k == "staging_bytes") {
        if (value < (1ll << 20)) return fail(SNAPGPU_EINVAL, "staging_bytes must be >= 1 MiB");
        o.staging_bytes = value & ~255ll;
    } else if (k == "sha_warps_per_sm") {
        if (value < 0 || value > kShaCtasPerSmMax) return fail(SNAPGPU_EINVAL, "sha_warps_per_sm out of range");
        o.sha_warps_per_sm = value;
    } else if (k == "sha_variant") {
        if (value < 0 || value >= kShaVariants) return fail(SNAPGPU_EINVAL, "sha_variant out of range");
        o.sha_variant = value;
    } else if (k == "cmp_ctas_per_sm") {
        if (value < 0 || value > 32) return fail(SNAPGPU_EINVAL, "cmp_ctas_per_sm out of range");
        o.cmp_ctas_per_sm = value;
    } else if (k == "time_kernels") {
        o.time_kernels = value ? 1 : 0;
    } else if (k == "feeders") {
        if (value < 0 || value > kFeeders) return fail(SNAPGPU_EINVAL, "feeders out of range");
        o.feeders = value;
    } else if (k == "two_ended") {
        if (value < 0 || value > 2) return fail(SNAPGPU_EINVAL, "two_ended: 0 off, 1 when the launch's lengths are spread wide (default), 2 always");
        o.two_ended = value;
    } else if (k == "pair_form") {
        if (value < 0 || value > 1) return fail(SNAPGPU_EINVAL, "pair_form: 0 shared-memory mailboxes, 1 shuffle exchange");
        o.pair_form = value;
    } else if (k == "taper") {
        o.taper = value ? 1 : 0;
    } else if (k == "long_min_blocks") {
        if (value < 0) return fail(SNAPGPU_EINVAL, "long_min_blocks: 0 = default, else the smallest file (in 128-byte blocks) the long-file bin takes");
        o.long_min_blocks = value;
    } else if (k == "pair_files_per_cta") {
        if (value < 0 || value > kPairFilesPerCta) return fail(SNAPGPU_EINVAL, "pair_files_per_cta: 0 auto, 1..16");
        o.pair_files_per_cta = value;
    } else if (k == "long_kernel") {
        if (value < 0 || value > 2) return fail(SNAPGPU_EINVAL, "long_kernel: 0 off, 1 one lane per file, 2 lane pair per file");
        o.long_kernel = value;
    } else {
        return fail(SNAPGPU_EINVAL, "unknown option %s", key);
    }
    return 0;
}

void *snapgpu_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    DevsInUse use(rt().devs_mu);
    if (rt().devs.empty()) {
        set_error("snapgpu_init has not been called (or failed)");
        return nullptr;
    }
    cudaSetDevice(rt().devs[0]->ordinal);
    cudaError_t e = host_pinned_alloc(&p, bytes);
    if (e != cudaSuccess) {
        set_error("pinning %zu bytes of host memory failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

void snapgpu_free_pinned(void *p) { host_pinned_free(p); }

void snapgpu_free(void *p) { free(p); }

int snapgpu_sha512_batch(const uint8_t *data, const uint64_t *offsets, const uint64_t *lengths, size_t nfiles,
                         uint8_t *digests) {
    DevsInUse use(rt().devs_mu);
    const bool ready = !rt().devs.empty();
    if (nfiles == 0) return ready ? 0 : fail(SNAPGPU_ENOINIT, "snapgpu_init has not been called (or failed)");
    if (!data || !offsets || !lengths || !digests) return fail(SNAPGPU_EINVAL, "null argument");
    if (!ready) return fail(SNAPGPU_ENOINIT, "snapgpu_init has not been called (or failed)");
    if (rt().devs.size() == 1)       // one device: the caller's arrays are the item list
        return sha512_shard(*rt().devs[0], data, ItemList(offsets, lengths, nfiles), digests);
    std::vector<size_t> cut;
    if (split_contiguous(lengths, nfiles, rt().devs.size(), [](uint64_t len) { return seg_blocks(len, 0) + 1; }, cut))
        return run_on_ranges(cut, [&](Device &D, size_t lo, size_t hi) {
            return sha512_shard(D, data, ItemList(offsets + lo, lengths + lo, hi - lo), digests + 64 * lo);
        });
    std::vector<WorkItem> all(nfiles);
    for (size_t i = 0; i < nfiles; i++) {
        if (lengths[i] >= kMaxSegBytes) return fail(SNAPGPU_EINVAL, "file %zu too large", i);
        all[i] = WorkItem{i, offsets[i], lengths[i], 0, 0};
    }
    return sha512_host_items(data, all, digests);
}

int snapgpu_sha512_stream(uint8_t state[64], int first, const uint8_t *data, uint64_t len, uint64_t prefix_bytes,
                          int final) {
    if (!state) return fail(SNAPGPU_EINVAL, "null state");
    HostSeg s{0, len, prefix_bytes, 0};
    if (!first) s.flags |= kHostSegContinue;
    if (!final) s.flags |= kHostSegNoFinal;
    static const uint8_t dummy[16] = {0};
    return sha512_host_segments(data ? data : dummy, &s, 1, state);
}

int snapgpu_cmp_batch(const uint8_t *a, const uint8_t *b, const uint64_t *offsets, const uint64_t *lengths,
                      size_t npairs, uint8_t *equal) {
    DevsInUse use(rt().devs_mu);
    if (rt().devs.empty()) return fail(SNAPGPU_ENOINIT, "snapgpu_init has not been called (or failed)");
    if (npairs == 0) return 0;
    if (!a || !b || !offsets || !lengths || !equal) return fail(SNAPGPU_EINVAL, "null argument");
    memset(equal, 1, npairs);
    if (rt().devs.size() == 1) return cmp_shard(*rt().devs[0], a, b, ItemList(offsets, lengths, npairs), equal);
    std::vector<size_t> cut;
    if (split_contiguous(lengths, npairs, rt().devs.size(), [](uint64_t len) { return len + 64; }, cut))
        return run_on_ranges(cut, [&](Device &D, size_t lo, size_t hi) {
            return cmp_shard(D, a, b, ItemList(offsets + lo, lengths + lo, hi - lo), equal + lo);
        });
    std::vector<WorkItem> all(npairs);
    std::vector<uint64_t> weight(npairs);
    for (size_t i = 0; i < npairs; i++) {
        all[i] = WorkItem{i, offsets[i], lengths[i], 0, 0};
        weight[i] = lengths[i] + 64;
    }
    std::vector<std::vector<WorkItem>> shards;
    shard_items(all, weight, (int)rt().devs.size(), shards);
    return run_on_devices(shards, [&](Device &D, const std::vector<WorkItem> &s) { return cmp_shard(D, a, b, s, equal); });
}

int snapgpu_sha512_batch_device(int dev, const void *d_data, const uint64_t *offsets, const uint64_t *lengths,
                                size_t nfiles, void *d_digests, void *stream) {
    DevsInUse use(rt().devs_mu);
    Device *dv = nullptr;
    int rc = get_device(dev, &dv);
    if (rc) return rc;
    if (nfiles == 0) return 0;
    if (!d_data || !offsets || !lengths || !d_digests) return fail(SNAPGPU_EINVAL, "null argument");
    PipeLease lease(*dv);
    Pipe *D = lease.pipe;
    SG_CUDA(cudaSetDevice(D->ordinal));
    if (nfiles > 0xfffffff0ull) return fail(SNAPGPU_EINVAL, "too many files in one launch (%zu)", nfiles);
    auto get = [offsets, lengths](size_t i) { return SegDesc{offsets[i], lengths[i], 0, (u32)i, 0}; };
    return launch_sha512(*D, (cudaStream_t)stream, static_cast<const uint8_t *>(d_data), get, nfiles,
                         static_cast<uint8_t *>(d_digests));
}

int snapgpu_cmp_batch_device(int dev, const void *d_a, const void *d_b, const uint64_t *offsets,
                             const uint64_t *lengths, size_t npairs, void *d_equal, void *stream) {
    DevsInUse use(rt().devs_mu);
    Device *dv = nullptr;
    int rc = get_device(dev, &dv);
    if (rc) return rc;
    if (npairs == 0) return 0;
    if (!d_a || !d_b || !offsets || !lengths || !d_equal) return fail(SNAPGPU_EINVAL, "null argument");
    PipeLease lease(*dv);
    Pipe *D = lease.pipe;
    SG_CUDA(cudaSetDevice(D->ordinal));
    std::vector<CmpItem> items(npairs);
    for (size_t i = 0; i < npairs; i++) items[i] = CmpItem{offsets[i], lengths[i]};
    return launch_cmp(*D, (cudaStream_t)stream, static_cast<const uint8_t *>(d_a), static_cast<const uint8_t *>(d_b),
                      items.data(), npairs, static_cast<uint8_t *>(d_equal));
}

int snapgpu_synth_fill_device(int dev, void *d_data, const uint64_t *offsets, const uint64_t *lengths, size_t nfiles,
                              uint64_t first_index, uint64_t seed, void *stream) {
    DevsInUse use(rt().devs_mu);
    Device *dv = nullptr;
    int rc = get_device(dev, &dv);
    if (rc) return rc;
    if (nfiles == 0) return 0;
    if (!d_data || !offsets || !lengths) return fail(SNAPGPU_EINVAL, "null argument");
    if (nfiles > 0xfffffff0ull) return fail(SNAPGPU_EINVAL, "too many files");
    PipeLease lease(*dv);
    Pipe *D = lease.pipe;
    SG_CUDA(cudaSetDevice(D->ordinal));
    PlanSlot *slot;
    if ((rc = acquire_slot(*D, nfiles * sizeof(SynthFile), &slot))) return rc;
    SynthFile *h = static_cast<SynthFile *>(slot->h_buf);
    for (size_t i = 0; i < nfiles; i++) h[i] = SynthFile{offsets[i], lengths[i]};
    cudaStream_t s = (cudaStream_t)stream;
    SG_CUDA(cudaMemcpyAsync(slot->d_buf, h, nfiles * sizeof(SynthFile), cudaMemcpyHostToDevice, s));
    const u32 grid = (u32)std::min<size_t>(nfiles, (size_t)D->sm_count * 16);
    synth_fill_kernel<<<grid, 256, 0, s>>>(static_cast<uint8_t *>(d_data), static_cast<const SynthFile *>(slot->d_buf),
                                           (u32)nfiles, first_index, seed);
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaEventRecord(slot->done, s));
    slot->in_flight = true;
    rt().kernel_launches++;
    return 0;
}

int snapgpu_get_stats(snapgpu_stats *out) {
    if (!out) return fail(SNAPGPU_EINVAL, "null argument");
    auto &R = rt();
    memset(out, 0, sizeof *out);
    out->kernel_launches = R.kernel_launches;
    out->sha512_launches = R.sha_launches;
    out->cmp_launches = R.cmp_launches;
    out->sha512_long_launches = R.sha_long_launches;
    out->h2d_bytes = R.h2d_bytes;
    out->d2h_bytes = R.d2h_bytes;
    double sha_sum = 0, cmp_sum = 0;
    uint64_t sha_n = 0, cmp_n = 0;
    DevsInUse use(R.devs_mu);
    for (auto &dp : R.devs)
        for (auto &pp : dp->pipes) {
            Pipe &D = *pp;
            std::lock_guard<std::mutex> lock(D.mu);
            cudaSetDevice(D.ordinal);
            harvest_timings(D.sha_t, 8, D.sha_ms_sum, D.sha_ms_n, D.sha_ms_last, true);
            harvest_timings(D.cmp_t, 8, D.cmp_ms_sum, D.cmp_ms_n, D.cmp_ms_last, true);
            sha_sum += D.sha_ms_sum;
            sha_n += D.sha_ms_n;
            cmp_sum += D.cmp_ms_sum;
            cmp_n += D.cmp_ms_n;
            if (D.sha_ms_n) out->last_sha512_kernel_ms = D.sha_ms_last;
            if (D.cmp_ms_n) out->last_cmp_kernel_ms = D.cmp_ms_last;
        }
    out->sha512_kernel_ms_sum = sha_sum;
    out->sha512_kernel_timed = sha_n;
    out->cmp_kernel_ms_sum = cmp_sum;
    out->cmp_kernel_timed = cmp_n;
    return 0;
}

void snapgpu_reset_stats(void) {
    auto &R = rt();
    R.kernel_launches = 0;
    R.sha_launches = 0;
    R.sha_long_launches = 0;
    R.cmp_launches = 0;
    R.h2d_bytes = 0;
    R.d2h_bytes = 0;
    DevsInUse use(R.devs_mu);
    for (auto &dp : R.devs)
        for (auto &pp : dp->pipes) {
            Pipe &D = *pp;
            std::lock_guard<std::mutex> lock(D.mu);
            cudaSetDevice(D.ordinal);
            harvest_timings(D.sha_t, 8, D.sha_ms_sum, D.sha_ms_n, D.sha_ms_last, true);
            harvest_timings(D.cmp_t, 8, D.cmp_ms_sum, D.cmp_ms_n, D.cmp_ms_last, true);
            D.sha_ms_sum = D.cmp_ms_sum = 0;
            D.sha_ms_n = D.cmp_ms_n = 0;
        }
}

// ---- test hooks: host logic only, callable without a GPU ---------------------------------

// Launch order the device-side length binning produces for files of these lengths:
// order[k] = index of the k-th file of the plan.  Needs a GPU (runs the plan kernels on device 0).
int snapgpu_test_plan_order(const uint64_t *lengths, size_t n, uint32_t *order) {
    if (!lengths || !order) return fail(SNAPGPU_EINVAL, "null argument");
    DevsInUse use(rt().devs_mu);
    Device *dv = nullptr;
    int rc = get_device(0, &dv);
    if (rc) return rc;
    if (n == 0) return 0;
    PipeLease lease(*dv);
    Pipe *D = lease.pipe;
    SG_CUDA(cudaSetDevice(D->ordinal));
    PlanSlot *slot;
    if ((rc = acquire_slot(*D, plan_bytes(n, kPlanTopMax + 1), &slot))) return rc;
    PlanInfo info;
    write_descriptors([lengths](size_t i) { return SegDesc{0, std::min<uint64_t>(lengths[i], kMaxSegBytes - 1), 0, (u32)i, 0}; },
                      n, static_cast<SegDesc *>(slot->h_buf), &info);
    const DevicePlan plan = plan_layout(slot->d_buf, n);
    cudaStream_t s = D->compute_stream;
    SG_CUDA(cudaMemcpyAsync(slot->d_buf, slot->h_buf, n * sizeof(SegDesc), cudaMemcpyHostToDevice, s));
    if ((rc = enqueue_length_binning(*D, s, plan, n, info.max_blocks))) return rc;
    SG_CUDA(cudaMemcpyAsync(order, plan.order, n * sizeof(u32), cudaMemcpyDeviceToHost, s));
    SG_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// The long-file bin's choice for files of these lengths on a device of sm_count SMs (long_mode as option
// long_kernel, min_blocks as option long_min_blocks): in_bin[i] = 1 for the files that leave the batched kernel,
// *per_cta = files per CTA of the lane-pair kernel (0 when nothing is binned).  Host logic only, no GPU.
int snapgpu_test_long_bin(const uint64_t *lengths, size_t n, int sm_count, int long_mode, long long min_blocks,
                          uint8_t *in_bin, uint32_t *per_cta) {
    if (!lengths || !in_bin || !per_cta || sm_count < 1 || n > 0xfffffff0ull) return fail(SNAPGPU_EINVAL, "bad argument");
    std::vector<SegDesc> descs(n + kLongMaxFiles);
    PlanInfo info;
    info.min_long_blocks = min_blocks > 0 ? (uint64_t)min_blocks : long_mode >= 2 ? kLongMinBlocksPair : kLongMinBlocksLane;
    write_descriptors([lengths](size_t i) { return SegDesc{0, std::min<uint64_t>(lengths[i], kMaxSegBytes - 1), 0, (u32)i, 0}; },
                      n, descs.data(), &info);
    uint64_t total_blocks = info.total_blocks, max_blocks = info.max_blocks;
    const size_t n_long = long_mode ? select_long_bin(info, descs.data(), n, sm_count, long_mode, &total_blocks, &max_blocks) : 0;
    for (size_t i = 0; i < n; i++) in_bin[i] = (descs[i].flags & kSegSkip) ? 1 : 0;
    *per_cta = n_long && long_mode >= 2 ? pair_cta_shape(n_long, sm_count, 0) : 0;
    return 0;
}

// Device each item of a batch is sharded to (weights as used by the sharder).
int snapgpu_test_shard(const uint64_t *weights, size_t n, int ndev, int *device_of) {
    if (!weights || !device_of || ndev < 1) return fail(SNAPGPU_EINVAL, "bad argument");
    std::vector<WorkItem> all(n);
    std::vector<uint64_t> w(weights, weights + n);
    for (size_t i = 0; i < n; i++) all[i] = WorkItem{i, 0, 0, 0, 0};
    std::vector<std::vector<WorkItem>> shards;
    shard_items(all, w, ndev, shards);
    for (int d = 0; d < ndev; d++)
        for (const WorkItem &it : shards[(size_t)d]) device_of[it.user_index] = d;
    return 0;
}

// The in-place multi-device split of a plain file list (split_contiguous): cut[0..ndev] are the
// index boundaries.  Returns 1 when the list can be cut this way, 0 when an item is too heavy
// (shard_items then places the heavy ones first), negative on a bad argument.
int snapgpu_test_split(const uint64_t *lengths, size_t n, int ndev, size_t *cut) {
    if (!lengths || !cut || ndev < 1) return fail(SNAPGPU_EINVAL, "bad argument");
    std::vector<size_t> c;
    const bool ok = split_contiguous(lengths, n, (size_t)ndev, [](uint64_t len) { return seg_blocks(len, 0) + 1; }, c);
    if (!ok) return 0;
    for (int d = 0; d <= ndev; d++) cut[d] = c[(size_t)d];
    return 1;
}

// Chunking of a host batch for a staging buffer of `cap` bytes.  Writes one row of six u64
// per produced item: user index, off, len, prefix, flags, chunk number.  Returns the number
// of rows (or a negative error); rows beyond max_rows are counted but not written.
long long snapgpu_test_chunks(const uint64_t *offsets, const uint64_t *lengths, size_t n, uint64_t cap, int is_sha,
                              uint64_t *rows, size_t max_rows) {
    if (!offsets || !lengths || cap < 4096) return fail(SNAPGPU_EINVAL, "bad argument");
    std::vector<WorkItem> in(n);
    ChunkPlan plan;
    for (size_t i = 0; i < n; i++) in[i] = WorkItem{i, offsets[i], lengths[i], 0, 0};
    if (is_sha == 2) {                       // as sha512_shard cuts a shard: chunk sizes ramp up and taper off
        const ItemList list(in);
        ChunkStream cs(list, true, plan);
        Chunk c;
        for (size_t ci = 0; cs.next(taper_cap(ramp_cap(ci, (size_t)cap), list, cs.k), (size_t)cap, &c); ci++) {}
    } else {
        build_chunks(in, (size_t)cap, is_sha != 0, plan);
    }
    const std::vector<Chunk> &chunks = plan.chunks;
    size_t row = 0;
    for (size_t c = 0; c < chunks.size(); c++)
        for (size_t i = 0; i < chunks[c].count; i++, row++) {
            if (!rows || row >= max_rows) continue;
            const WorkItem &w = plan.item(chunks[c], i);
            uint64_t *r = rows + 6 * row;
            r[0] = w.user_index; r[1] = w.off; r[2] = w.len; r[3] = w.prefix; r[4] = w.flags; r[5] = c;
        }
    return (long long)row;
}

// Raw host-to-device copy rate (bench support): every bound device copies `bytes_per_dev` bytes
// `reps` times from its own slice of `host` (device d reads host + d * bytes_per_dev) with the call
// the pipeline itself uses -- cudaMemcpyAsync on a pipe's copy stream -- all devices at once, one
// thread each.  *seconds = the longest device's time for its reps (CUDA events), the ceiling that
// the end-to-end figures are fractions of.
int snapgpu_h2d_probe(const void *host, size_t bytes_per_dev, int reps, double *seconds) {
    if (!host || !seconds || bytes_per_dev == 0 || reps < 1) return fail(SNAPGPU_EINVAL, "bad argument");
    DevsInUse use(rt().devs_mu);
    auto &R = rt();
    if (R.devs.empty()) return fail(SNAPGPU_ENOINIT, "snapgpu_init has not been called (or failed)");
    const size_t ndev = R.devs.size();
    std::vector<double> secs(ndev, 0.0);
    std::vector<int> rcs(ndev, 0);
    std::vector<std::string> errs(ndev);
    std::atomic<size_t> ready{0};
    auto run = [&](size_t d) {
        PipeLease lease(*R.devs[d]);
        Pipe &P = *lease.pipe;
        auto cuda_fail = [&](cudaError_t e, const char *what) {
            rcs[d] = SNAPGPU_ECUDA;
            errs[d] = std::string(what) + ": " + cudaGetErrorString(e);
        };
        cudaError_t e;
        if ((e = cudaSetDevice(P.ordinal)) != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
        void *dst = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if ((e = cudaMalloc(&dst, bytes_per_dev)) != cudaSuccess) return cuda_fail(e, "cudaMalloc");
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const uint8_t *src = static_cast<const uint8_t *>(host) + d * bytes_per_dev;
        e = cudaMemcpyAsync(dst, src, bytes_per_dev, cudaMemcpyHostToDevice, P.copy_stream);      // warm-up
        if (e == cudaSuccess) e = cudaStreamSynchronize(P.copy_stream);
        ready++;
        while (ready.load() < ndev) std::this_thread::yield();       // all devices start together
        if (e == cudaSuccess) e = cudaEventRecord(e0, P.copy_stream);
        for (int r = 0; r < reps && e == cudaSuccess; r++)
            e = cudaMemcpyAsync(dst, src, bytes_per_dev, cudaMemcpyHostToDevice, P.copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(e1, P.copy_stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        secs[d] = ms * 1e-3;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(dst);
        if (e != cudaSuccess) cuda_fail(e, "host-to-device copy");
    };
    std::vector<std::thread> th;
    for (size_t d = 1; d < ndev; d++) th.emplace_back(run, d);
    run(0);
    for (auto &t : th) t.join();
    for (size_t d = 0; d < ndev; d++)
        if (rcs[d]) return fail(rcs[d], "device %zu: %s", d, errs[d].c_str());
    *seconds = *std::max_element(secs.begin(), secs.end());
    return 0;
}

typedef void (*ProbeKernel)(uint32_t *, int, uint32_t, uint32_t, unsigned long long *);

int snapgpu_pipe_microbench(int dev, int kind, int warps_per_sm, double *inst_per_clk_per_sm, double *elapsed_ms,
                            double *sm_clock_mhz) {
    DevsInUse use(rt().devs_mu);
    Device *dv = nullptr;
    int rc = get_device(dev, &dv);
    if (rc) return rc;
    static const ProbeKernel table[kProbeCount] = {
        pipe_probe_kernel<0>, pipe_probe_kernel<1>, pipe_probe_kernel<2>, pipe_probe_kernel<3>,
        pipe_probe_kernel<4>, pipe_probe_kernel<5>, pipe_probe_kernel<6>, pipe_probe_kernel<7>,
        pipe_probe_kernel<8>, pipe_probe_kernel<9>, pipe_probe_kernel<10>, pipe_probe_kernel<11>,
    };
    if (kind < 0 || kind >= kProbeCount) return fail(SNAPGPU_EINVAL, "unknown probe kind %d", kind);
    if (warps_per_sm < 4 || warps_per_sm > 64 || warps_per_sm % 4) return fail(SNAPGPU_EINVAL, "warps_per_sm must be 4..64, multiple of 4");
    PipeLease lease(*dv);
    Pipe *D = lease.pipe;
    SG_CUDA(cudaSetDevice(D->ordinal));
    const int ctas_per_sm = warps_per_sm / (kProbeThreads / 32);
    const int grid = D->sm_count * ctas_per_sm;
    const int iters = 4096;
    uint32_t *d_out = nullptr;
    unsigned long long *d_clk = nullptr;
    SG_CUDA(cudaMalloc(&d_out, (size_t)grid * kProbeThreads * sizeof(uint32_t)));
    SG_CUDA(cudaMalloc(&d_clk, 2 * sizeof(unsigned long long)));
    cudaEvent_t e0, e1;
    SG_CUDA(cudaEventCreate(&e0));
    SG_CUDA(cudaEventCreate(&e1));
    cudaStream_t s = D->compute_stream;
    float best = 1e30f;
    unsigned long long clk[2] = {0, 0};
    for (int rep = 0; rep < 4; rep++) {   // first repetition is the warm-up
        SG_CUDA(cudaEventRecord(e0, s));
        table[kind]<<<grid, kProbeThreads, 0, s>>>(d_out, iters, 1u, 0x9e3779b9u, d_clk);
        SG_CUDA(cudaGetLastError());
        SG_CUDA(cudaEventRecord(e1, s));
        SG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        SG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) {
            best = ms;
            SG_CUDA(cudaMemcpy(clk, d_clk, sizeof clk, cudaMemcpyDeviceToHost));
        }
    }
    rt().kernel_launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    cudaFree(d_clk);
    const double mhz = clk[1] ? (double)clk[0] / (double)clk[1] * 1e3 : 0.0;
    // Rate from the whole launch (CUDA events x measured SM clock): one CTA's own clock64 span
    // is misleading because the warp scheduler lets the oldest CTA of an SM run ahead.
    const double warp_insts = (double)iters * probe_ops_per_iter(kind) * warps_per_sm;   // per SM
    const double cycles = best * 1e-3 * mhz * 1e6;
    if (inst_per_clk_per_sm) *inst_per_clk_per_sm = cycles > 0 ? warp_insts / cycles : 0.0;
    if (elapsed_ms) *elapsed_ms = best;
    if (sm_clock_mhz) *sm_clock_mhz = mhz;
    return 0;
}

}  // extern "C"
