/*
 * snapgpu.h -- C ABI of libsnapgpu, the B200 (sm_100a) drop-in for the one data-parallel
 * hot path of Ubuntu's `snappy` package manager: per-file SHA-512 for hashes.yaml and the
 * byte-wise file compare.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Every entry point below names the reference interface it replaces (paths relative to the
 * reference tree).  INTEGRATION.md shows the cgo binding for each.
 *
 * Conventions
 *   - int-returning functions: 0 on success, negative on failure (SNAPGPU_E*); the text of
 *     the failure is in snapgpu_last_error() (thread-local, owned by the library).
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails.
 *   - every call is synchronous on return (the caller's buffers may be reused or freed, as
 *     cgo requires), except the *_device calls, which enqueue on the stream they are given.
 *   - thread-safe: calls may come from any OS thread; per-device work is serialised inside.
 */
#ifndef SNAPGPU_H
#define SNAPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNAPGPU_OK        0
#define SNAPGPU_ECUDA    -1   /* CUDA runtime/driver failure (no device, OOM, launch error) */
#define SNAPGPU_EINVAL   -2   /* bad argument */
#define SNAPGPU_EIO      -3   /* host file I/O failed; errno text in last_error */
#define SNAPGPU_EMODE    -4   /* "Unknown file mode" (snappy/hashes.go:44) */
#define SNAPGPU_ENAME    -5   /* file name outside what the YAML writer supports */
#define SNAPGPU_ENOINIT  -6   /* snapgpu_init not called / failed */

/* ---- lifetime ---------------------------------------------------------------------- */

/* Bind the library to `ndev` CUDA devices (ordinals in `devices`; NULL = 0..ndev-1;
 * ndev <= 0 = every visible device).  Creates per-device streams and staging buffers.
 * Idempotent; a second call with a different device list re-initialises. */
int snapgpu_init(const int *devices, int ndev);
void snapgpu_shutdown(void);
int snapgpu_num_devices(void);                 /* devices bound by snapgpu_init */
const char *snapgpu_last_error(void);
const char *snapgpu_version(void);

/* Tunables (all optional).  Known keys: "staging_bytes" (per-device H2D chunk, default
 * 1 GiB, two buffers), "sha_warps_per_sm" (0 = auto), "sha_variant" (0 = default kernel),
 * "cmp_ctas_per_sm", "time_kernels", "feeders" (host threads per device that bounce pageable
 * input into pinned memory, 0 = auto), "long_kernel" (the long-file bin: 0 off, 1 one lane per
 * file, 2 a lane pair per file = default), "pair_form" (how the two lanes of a pair exchange round results:
 * 0 shared-memory mailboxes = default, 1 warp shuffle; sha512_pair.cuh), "long_min_blocks" (smallest file, in
 * 128-byte blocks, the long-file bin considers; 0 = default: 256 for the lane-pair form), "pair_files_per_cta" (0 = default: one
 * long file per CTA while a quarter of the SMs last, then 2, then 16; 1..16 = exactly this many per CTA), "taper"
 * (1 = default: a host-buffer call ends on a 128 MiB and a 64 MiB chunk), "two_ended" (claims from both ends of the
 * length-sorted plan: 0 never, 1 = default: when the longest file is at least three times the mean, 2 always). */
int snapgpu_set_option(const char *key, long long value);

/* C-owned pinned host memory for the Go side to pack file contents into
 * (cgo: C must not keep Go pointers; pinned memory makes the H2D copy DMA-direct). */
void *snapgpu_alloc_pinned(size_t bytes);
void snapgpu_free_pinned(void *p);

/* ---- batch kernels, host buffers (the end-to-end path) -------------------------------
 *
 * Replaces the arithmetic of helpers.Sha512sum (helpers/helpers.go:188-201; Go stdlib
 * crypto/sha512 behind it) for `nfiles` files at once.  File i is
 * data[offsets[i] .. offsets[i]+lengths[i]).  digests receives nfiles*64 bytes: the
 * big-endian H0..H7 of FIPS 180-4, i.e. exactly hasher.Sum(nil).  Fast path when every
 * offset is a multiple of 16 (and `data` is 16-byte aligned); any alignment is accepted.
 * With several devices bound the file list is sharded across them; no collective.
 */
int snapgpu_sha512_batch(const uint8_t *data, const uint64_t *offsets, const uint64_t *lengths,
                         size_t nfiles, uint8_t *digests);

/* Streaming form for one long message, the shape of Go's hash.Hash Write/Sum that
 * helpers.Sha512sum drives through io.Copy (helpers/helpers.go:195-200): `state` is 64
 * caller-owned bytes carrying the chaining value between calls.  first != 0 starts from
 * the SHA-512 IV; every call but the final one must pass a multiple of 128 bytes;
 * prefix_bytes = bytes already hashed; final != 0 pads and leaves the digest in `state`. */
int snapgpu_sha512_stream(uint8_t state[64], int first, const uint8_t *data, uint64_t len,
                          uint64_t prefix_bytes, int final);

/* The same as an object with the method set of Go's hash.Hash, for io.Copy(hasher, r)
 * (helpers/helpers.go:195-196) or io.MultiWriter(tarball, hasher) while data.tar.gz is being
 * written (clickdeb/deb.go:360-366 -> snappy/build.go:222): Write gathers bytes in pinned
 * memory and hashes each full 512 KiB piece on a worker thread while the caller goes on writing;
 * Sum returns the digest of everything written so far without disturbing the state. */
typedef struct snapgpu_hasher snapgpu_hasher;
snapgpu_hasher *snapgpu_hasher_new(void);                                     /* sha512.New()  */
int snapgpu_hasher_write(snapgpu_hasher *h, const uint8_t *p, size_t n);      /* Write         */
int snapgpu_hasher_sum(snapgpu_hasher *h, uint8_t digest[64]);                /* Sum(nil)      */
void snapgpu_hasher_free(snapgpu_hasher *h);

/* Replaces streamsEqual / bytes.Equal (helpers/cmp.go:61-86) for `npairs` pairs whose sizes
 * were already found equal on the host (helpers/cmp.go:54).  Pair i is a[off..off+len) vs
 * b[off..off+len).  equal[i] = 1 if identical, else 0. */
int snapgpu_cmp_batch(const uint8_t *a, const uint8_t *b, const uint64_t *offsets,
                      const uint64_t *lengths, size_t npairs, uint8_t *equal);

/* ---- batch kernels, device-resident data (kernel-only path) --------------------------
 *
 * Same contracts, but `d_data`/`d_a`/`d_b`/`d_digests`/`d_equal` are device pointers on
 * bound device number `dev` (index into the snapgpu_init list) and the work is enqueued on
 * `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream).  offsets and
 * lengths stay host arrays (they come from stat(2)).  The allocations behind d_data, d_a
 * and d_b must extend to the next 16-byte boundary past the last file or pair: the aligned
 * kernels read whole 16-byte words (bytes past an item's end never reach a digest or a
 * verdict).  cudaMalloc'd buffers always do.  Returns after enqueueing.
 */
int snapgpu_sha512_batch_device(int dev, const void *d_data, const uint64_t *offsets,
                                const uint64_t *lengths, size_t nfiles, void *d_digests,
                                void *stream);
int snapgpu_cmp_batch_device(int dev, const void *d_a, const void *d_b, const uint64_t *offsets,
                             const uint64_t *lengths, size_t npairs, void *d_equal, void *stream);

/* ---- whole-function drop-ins (host side in C++, same semantics as the Go functions) --- */

/* helpers.Sha512sum (helpers/helpers.go:188): lowercase hex digest of one file.
 * SNAPGPU_EIO mirrors the os.Open / io.Copy error return. */
int snapgpu_sha512sum_file(const char *infile, char hexdigest[129]);

/* writeHashes (snappy/build.go:216-270): hash data_tar, walk build_dir in filepath.Walk
 * order, skip "/DEBIAN*", emit DEBIAN/hashes.yaml byte-identical to yaml.v2's output. */
int snapgpu_write_hashes(const char *build_dir, const char *data_tar);

/* Same walk, but returns the YAML in a malloc'd buffer (*out, *out_len; free with
 * snapgpu_free) instead of only writing the file; used by tests and the Go shim's
 * verification mode. */
int snapgpu_hashes_yaml(const char *build_dir, const char *data_tar, char **out, size_t *out_len);

/* writeHashes when archive-sha512 is known already: clickdeb.Build writes data.tar.gz through gzip -9
 * (clickdeb/deb.go:261-285,360-366) far slower than one SHA-512 chain runs on the GPU, so a
 * snapgpu_hasher fed by an io.MultiWriter beside the file has the digest the moment the archive is
 * closed, and the callback (snappy/build.go:517-520) passes it here instead of the file name.  The
 * finished archive is not read again; everything else is snapgpu_write_hashes / snapgpu_hashes_yaml. */
int snapgpu_write_hashes_digest(const char *build_dir, const uint8_t archive_sha512[64]);
int snapgpu_hashes_yaml_digest(const char *build_dir, const uint8_t archive_sha512[64], char **out,
                               size_t *out_len);

/* Phases of the most recent snapgpu_write_hashes / snapgpu_hashes_yaml / snapgpu_verify_hashes
 * made by the calling thread (bench support).  Directory scan, file reads, host-to-device
 * copies and kernels overlap inside pack_ms; the tails are what was left to wait for after the
 * last file had been read. */
typedef struct snapgpu_tree_stats_t {
    double total_ms;
    double pack_ms;         /* scan + open/read/close of every file (GPU batches run underneath) */
    double gpu_tail_ms;     /* waiting for the last batches after the last file was packed */
    double chain_tail_ms;   /* waiting for the archive's (and other long files') chain after that */
    double yaml_ms;         /* walk-order assembly + document */
    uint64_t entries, files_hashed, files_cached, batches, yaml_bytes;
    unsigned pack_threads;
} snapgpu_tree_stats_t;
int snapgpu_tree_stats(snapgpu_tree_stats_t *out);

/* helpers.FilesAreEqual (helpers/cmp.go:31): 1 equal, 0 not equal OR any error. */
int snapgpu_files_are_equal(const char *a, const char *b);

/* helpers.DirUpdated (helpers/cmp.go:97): names (pfx+basename) of files present in both
 * directories whose contents differ.  *names is a malloc'd block of NUL-terminated strings
 * laid end to end (sorted), *count their number; free with snapgpu_free.  All pairs of one
 * call are compared in a single batched launch. */
int snapgpu_dir_updated(const char *dir_a, const char *dir_b, const char *pfx, char **names,
                        size_t *count);

/* policy.AppArmorDelta (policy/policy.go:162): DirUpdated over policygroups and templates. */
int snapgpu_apparmor_delta(const char *old_path, const char *new_path, const char *prefix,
                           char **policies, size_t *npolicies, char **templates,
                           size_t *ntemplates);

/* copyToBuildDir (snappy/build.go:362-418), the step of snappy.Build that stages the source
 * tree before writeHashes runs: removes an empty build_dir, walks source_dir in filepath.Walk
 * order skipping shouldExclude names (snappy/build.go:52-83; directories with their subtree),
 * creates directories with the source's mode, hard-links files and copies what cannot be linked.
 * Copied files are read ONCE: the bytes go to the build dir and through the SHA-512 kernel in the
 * same pass, and the digest is kept (keyed by device and inode of the written file, valid while its
 * size, mtime and ctime stay what they were) for the ONE snapgpu_write_hashes that follows, which then
 * does not read them again.  On error the copy may be
 * partial, as in the reference.  flags: SNAPGPU_COPY_NO_LINK = never hard-link. */
#define SNAPGPU_COPY_NO_LINK 1
int snapgpu_copy_to_build_dir(const char *source_dir, const char *build_dir, int flags);

/* Optional warm-up: does ahead of time what the first snapgpu_write_hashes of a process otherwise waits for (pinning
 * the file packer's chunk pool, 256 MiB; staging, digest and plan buffers; the first launch of each kernel).  A build
 * calls it from a goroutine as soon as it starts: by the time the tree has been copied and
 * the archive compressed, writeHashes runs warm (config 2: 42 ms instead of 80-250).  Replaces nothing in the reference. */
int snapgpu_warm(void);
int snapgpu_should_exclude(const char *base_name);          /* shouldExclude, 1 = excluded */
void snapgpu_digest_cache_clear(void);
void snapgpu_digest_cache_stats(size_t *entries, uint64_t *hits);

/* hashes.yaml verification -- a consumer the reference does not have: it writes the per-file
 * list (snappy/build.go:249-256) but reads back only archive-sha512 (snappy/snapp.go:466-478).
 * Re-hashes the tree under `root` exactly as snapgpu_write_hashes would (DEBIAN/ is neither
 * created nor listed) and compares entry by entry with the document at yaml_path.  *report gets
 * NUL-terminated strings laid end to end (free with snapgpu_free), *count their number:
 *   "missing: NAME"   listed but not in the tree      "extra: NAME"   in the tree but not listed
 *   "changed: NAME (size sha512 mode)"  with the fields that differ
 *   "archive-sha512 differs"  only when data_tar is given (NULL = do not check the archive)
 * NAME is the name scalar as the document spells it.  count == 0 means the tree verifies.  A
 * yaml_path inside the tree (meta/hashes.yaml at install time) is not reported as extra. */
int snapgpu_verify_hashes(const char *root, const char *yaml_path, const char *data_tar, char **report,
                          size_t *count);

/* The reader side, as the reference uses it: NewSnapPartFromYaml reads meta/hashes.yaml and keeps
 * archive-sha512 as the part's hash (snappy/snapp.go:466-478).  Like yaml.Unmarshal into hashesYaml
 * this decodes every entry's mode (yamlFileMode.UnmarshalYAML, snappy/hashes.go:59-88): a mode that
 * does not start with d, f or l fails the read with SNAPGPU_EMODE "Unknown file mode ...".  Host
 * logic only (no GPU).  hexdigest gets the NUL-terminated value (129 bytes hold a SHA-512). */
int snapgpu_read_archive_sha512(const char *yaml_path, char *hexdigest, size_t cap);

void snapgpu_free(void *p);

/* ---- synthetic inputs and instrumentation (bench/test support, not product API) ------ */

/* Fill device memory with the deterministic benchmark content of SURVEY.md section 8(d):
 * little-endian u64 word j of file i = splitmix64(seed + i*0x9E3779B97F4A7C15 + j). */
int snapgpu_synth_fill_device(int dev, void *d_data, const uint64_t *offsets,
                              const uint64_t *lengths, size_t nfiles, uint64_t first_index,
                              uint64_t seed, void *stream);

/* Counters since init (or the last reset): kernel launches made by this library, and the
 * device time of the last sha512/cmp launch group as measured with CUDA events. */
typedef struct snapgpu_stats {
    uint64_t kernel_launches;
    uint64_t sha512_launches;
    uint64_t cmp_launches;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    double last_sha512_kernel_ms;   /* CUDA-event time of the most recent launch */
    double last_cmp_kernel_ms;
    double sha512_kernel_ms_sum;    /* summed over the launches timed since the last reset */
    uint64_t sha512_kernel_timed;
    double cmp_kernel_ms_sum;
    uint64_t cmp_kernel_timed;
    uint64_t sha512_long_launches;  /* launches of the long-file bin's kernel (counted in kernel_launches too) */
} snapgpu_stats;
int snapgpu_get_stats(snapgpu_stats *out);
void snapgpu_reset_stats(void);

/* Integer-pipe micro-benchmark that defines the measured SHA-512 roofline: runs a
 * register-only loop of the named instruction mix on device `dev` and returns warp
 * instructions per clock per SM.  kind: 0 IADD3, 1 LOP3, 2 SHF, 3 IMAD, 4 IMAD.WIDE,
 * 5 ALU+IMAD interleaved, 6 ALU+IMAD.WIDE interleaved. */
int snapgpu_pipe_microbench(int dev, int kind, int warps_per_sm, double *inst_per_clk_per_sm,
                            double *elapsed_ms, double *sm_clock_mhz);

/* Raw pinned host-to-device copy rate on every bound device at once (device d copies
 * bytes_per_dev bytes from host + d * bytes_per_dev, `reps` times, through the same
 * cudaMemcpyAsync path the pipeline uses): *seconds = the slowest device's time.  The ceiling
 * the end-to-end figures are stated against. */
int snapgpu_h2d_probe(const void *host, size_t bytes_per_dev, int reps, double *seconds);

/* ---- test hooks (see tests/) ------------------------------------------------------------ */
/* host logic only, usable without a GPU */
int snapgpu_test_yaml_from_digests(const char *build_dir, const uint8_t *digests, size_t ndigests,
                                   char **out, size_t *out_len);
/* one fileHash as yaml.v2 renders it (snappy/hashes_test.go:30-33); size < 0 / sha512_hex NULL = omitted */
int snapgpu_test_filehash_yaml(const char *name, long long size, const char *sha512_hex, unsigned mode,
                               char **out, size_t *out_len);
int snapgpu_test_shard(const uint64_t *weights, size_t n, int ndev, int *device_of);
int snapgpu_test_split(const uint64_t *lengths, size_t n, int ndev, size_t *cut);   /* 1 = cut in place, 0 = too heavy an item */
/* rows of (user index, offset, length, prefix, flags, chunk) as the host-buffer pipeline cuts a shard into chunks of at
 * most `cap` bytes: is_sha 0 compare, 1 SHA-512 with one fixed chunk size, 2 SHA-512 with the pipeline's sizes (64 MiB,
 * 256 MiB, then `cap`, ending on 128 MiB and 64 MiB) */
long long snapgpu_test_chunks(const uint64_t *offsets, const uint64_t *lengths, size_t n,
                              uint64_t cap, int is_sha, uint64_t *rows, size_t max_rows);
/* which files the long-file bin takes out of a launch of files of these lengths on a device of sm_count SMs (long_mode,
 * min_blocks: options long_kernel and long_min_blocks), and how many files a lane-pair CTA gets; host logic, no GPU */
int snapgpu_test_long_bin(const uint64_t *lengths, size_t n, int sm_count, int long_mode, long long min_blocks,
                          uint8_t *in_bin, uint32_t *per_cta);
/* the order the device-side length binning gives files of these lengths (needs a GPU) */
int snapgpu_test_plan_order(const uint64_t *lengths, size_t n, uint32_t *order);

#ifdef __cplusplus
}
#endif
#endif /* SNAPGPU_H */
