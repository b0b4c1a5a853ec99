// runtime.hpp -- internal interfaces of libsnapgpu shared by the CUDA runtime layer
// (snapgpu.cu) and the host-side mirror of the reference's Go functions (host_path.cpp).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/snapgpu.h"

namespace snapgpu {

// thread-local error text behind snapgpu_last_error()
void set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

// A piece of a message inside a packed host buffer.  Segment i uses digest slot i; with
// kHostSegContinue the slot holds the chaining value to start from, with kHostSegNoFinal the
// chaining value (not the padded digest) is written back.
struct HostSeg {
    uint64_t off, len, prefix;
    uint32_t flags;
};
enum : uint32_t { kHostSegContinue = 1u, kHostSegNoFinal = 2u };

// Hash `n` segments of `data` on the bound devices; `digests` is n*64 bytes, in/out.
int sha512_host_segments(const uint8_t *data, const HostSeg *segs, size_t n, uint8_t *digests);

// ---- batch session ------------------------------------------------------------------------
// The streaming form of the host-buffer pipeline, for a caller that produces its input piece by
// piece (the tree hasher of host_path.cpp): batches of whole messages lying in pinned host
// memory are enqueued as they become ready -- host-to-device copies, length binning, the SHA-512
// kernel and the copy of the digests back all run behind the call -- on a pipe of each bound
// device held for the session's lifetime (batches go to the device with a free slot, two per
// device).  One thread drives a session.
struct HostSpan {
    const uint8_t *ptr;     // pinned (cudaHostAlloc) memory
    size_t bytes;
};
struct SpanSeg {
    uint32_t span;          // index into the batch's spans
    uint64_t off, len;      // the message is spans[span].ptr[off .. off+len)
};
class BatchSession;
int session_open(BatchSession **out, size_t max_batch_bytes);
void session_close(BatchSession *s);
// Enqueue one batch; *ticket names it.  Blocks only while every slot is taken (then it waits for
// the oldest batch).  digest_dst[i] receives the 64-byte digest of segs[i] once the batch is done.
// Tickets retired while waiting are appended to *copied like session_poll does.
int session_submit(BatchSession *s, const HostSpan *spans, size_t nspans, const SpanSeg *segs,
                   uint8_t *const *digest_dst, size_t nsegs, uint64_t *ticket, std::vector<uint64_t> *copied);
// Progress without blocking: tickets whose host-to-device copies have finished since the last
// call are appended to *copied (their spans may be overwritten from then on); finished batches
// have their digests stored.  wait_all: block until every batch is done.
int session_poll(BatchSession *s, std::vector<uint64_t> *copied, bool wait_all);
size_t session_in_flight(const BatchSession *s);
size_t session_capacity(const BatchSession *s);     // batches that can be in flight at once

bool runtime_ready();
int ensure_init();              // lazy snapgpu_init(NULL, 0) for the whole-function drop-ins
size_t staging_bytes();

std::string hex_lower(const uint8_t *p, size_t n);

}  // namespace snapgpu
