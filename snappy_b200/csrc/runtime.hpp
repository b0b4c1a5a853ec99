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

bool runtime_ready();
int ensure_init();              // lazy snapgpu_init(NULL, 0) for the whole-function drop-ins
size_t staging_bytes();

std::string hex_lower(const uint8_t *p, size_t n);

}  // namespace snapgpu
