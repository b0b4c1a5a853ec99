// fake_runtime.cpp -- TEST INFRASTRUCTURE, never shipped: stands in for snapgpu.cu so that the
// host side of libsnapgpu (snappy_b200/csrc/host_path.cpp: tree walk, packer pool, chunk
// recycling, chain streamer, YAML writer, copyToBuildDir, DirUpdated) can run on a machine
// without a GPU, under AddressSanitizer / UndefinedBehaviorSanitizer / ThreadSanitizer.
//
// The arithmetic comes from the CPU oracle (oracle/sha512_oracle.c), which only tests may
// link.  The product library has no such path: without CUDA its entry points fail.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../snappy_b200/csrc/runtime.hpp"

extern "C" {
// layout of the oracle's streaming context (oracle/sha512_oracle.c)
typedef struct {
    uint64_t h[8];
    uint8_t x[128];
    size_t nx;
    uint64_t len;
} oracle_sha512_ctx;
void oracle_sha512_init(oracle_sha512_ctx *c);
void oracle_sha512_update(oracle_sha512_ctx *c, const uint8_t *p, size_t n);
void oracle_sha512_final(const oracle_sha512_ctx *c, uint8_t out[64]);
}

namespace snapgpu {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
std::string hex_lower(const uint8_t *p, size_t n) {
    static const char d[] = "0123456789abcdef";
    std::string s(2 * n, '0');
    for (size_t i = 0; i < n; i++) {
        s[2 * i] = d[p[i] >> 4];
        s[2 * i + 1] = d[p[i] & 15];
    }
    return s;
}
bool runtime_ready() { return true; }
int ensure_init() { return 0; }
size_t staging_bytes() {
    const char *e = getenv("HOSTSIM_STAGING");
    return e ? (size_t)atoll(e) : (size_t)64 << 20;
}

static void one_segment(const uint8_t *data, uint64_t len, uint64_t prefix, bool cont, bool no_final, uint8_t *state) {
    static const bool nohash = getenv("HOSTSIM_NOHASH") != nullptr;      // host-pipeline timing runs: digests stay zero
    if (nohash) {
        memset(state, 0xab, 64);
        return;
    }
    oracle_sha512_ctx c;
    oracle_sha512_init(&c);
    if (cont) {
        for (int i = 0; i < 8; i++) {
            uint64_t v = 0;
            for (int k = 0; k < 8; k++) v = (v << 8) | state[8 * i + k];
            c.h[i] = v;
        }
        c.len = prefix;
    }
    oracle_sha512_update(&c, data, (size_t)len);
    if (no_final) {
        for (int i = 0; i < 8; i++)
            for (int k = 0; k < 8; k++) state[8 * i + k] = (uint8_t)(c.h[i] >> (56 - 8 * k));
    } else {
        oracle_sha512_final(&c, state);
    }
}

int sha512_host_segments(const uint8_t *data, const HostSeg *segs, size_t n, uint8_t *digests) {
    for (size_t i = 0; i < n; i++) {
        const bool no_final = segs[i].flags & kHostSegNoFinal;
        if (no_final && (segs[i].len & 127)) return fail(SNAPGPU_EINVAL, "non-final segment %zu is not a multiple of 128 bytes", i);
        one_segment(data + segs[i].off, segs[i].len, segs[i].prefix, segs[i].flags & kHostSegContinue, no_final, digests + 64 * i);
    }
    return 0;
}

// a session whose batches complete at once, but whose "copied" reports arrive one poll late and
// whose slots fill up, so that the caller's recycling and back-pressure paths run
class BatchSession {
public:
    std::vector<uint64_t> unreported;
    uint64_t next = 1;
    size_t max_bytes = 0;
};
int session_open(BatchSession **out, size_t max_batch_bytes) {
    *out = new BatchSession();
    (*out)->max_bytes = max_batch_bytes;
    return 0;
}
void session_close(BatchSession *s) { delete s; }
size_t session_in_flight(const BatchSession *s) { return s->unreported.size(); }
size_t session_capacity(const BatchSession *) { return 3; }
int session_poll(BatchSession *s, std::vector<uint64_t> *copied, bool wait_all) {
    if (s->unreported.empty()) return 0;
    if (wait_all) {
        if (copied) copied->insert(copied->end(), s->unreported.begin(), s->unreported.end());
        s->unreported.clear();
    } else {
        if (copied) copied->push_back(s->unreported.front());
        s->unreported.erase(s->unreported.begin());
    }
    return 0;
}
int session_submit(BatchSession *s, const HostSpan *spans, size_t nspans, const SpanSeg *segs, uint8_t *const *digest_dst,
                   size_t nsegs, uint64_t *ticket, std::vector<uint64_t> *copied) {
    size_t total = 0;
    for (size_t k = 0; k < nspans; k++) total += (spans[k].bytes + 255) & ~(size_t)255;
    if (total > s->max_bytes) return fail(SNAPGPU_EINVAL, "batch of %zu bytes exceeds the session's %zu", total, s->max_bytes);
    if (getenv("HOSTSIM_FAIL_SUBMIT") && s->next == (uint64_t)atoll(getenv("HOSTSIM_FAIL_SUBMIT")))
        return fail(SNAPGPU_ECUDA, "injected batch failure");
    while (s->unreported.size() >= 3) session_poll(s, copied, false);
    for (size_t i = 0; i < nsegs; i++) {
        if (segs[i].span >= nspans || segs[i].off + segs[i].len > spans[segs[i].span].bytes)
            return fail(SNAPGPU_EINVAL, "segment %zu outside its span", i);
        one_segment(spans[segs[i].span].ptr + segs[i].off, segs[i].len, 0, false, false, digest_dst[i]);
    }
    *ticket = s->next++;
    s->unreported.push_back(*ticket);
    return 0;
}

}  // namespace snapgpu

extern "C" {
const char *snapgpu_last_error(void) { return snapgpu::g_err.c_str(); }
int snapgpu_init(const int *, int) { return 0; }
void *snapgpu_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (posix_memalign(&p, 4096, bytes ? bytes : 1)) return nullptr;
    return p;
}
void snapgpu_free_pinned(void *p) { free(p); }
void snapgpu_free(void *p) { free(p); }
int snapgpu_cmp_batch(const uint8_t *a, const uint8_t *b, const uint64_t *offsets, const uint64_t *lengths, size_t npairs,
                      uint8_t *equal) {
    for (size_t i = 0; i < npairs; i++) equal[i] = memcmp(a + offsets[i], b + offsets[i], (size_t)lengths[i]) == 0;
    return 0;
}
}
