/*
 * oracle/sha512_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the reference's hashing and compare path:
 *
 *   helpers.Sha512sum        /root/reference/helpers/helpers.go:188-201
 *   SystemImagePart.Hash     /root/reference/snappy/systemimage.go:122-128
 *   helpers.FilesAreEqual    /root/reference/helpers/cmp.go:31-59
 *   streamsEqual             /root/reference/helpers/cmp.go:61-86   (bufsz: cmp.go:27)
 *
 * The arithmetic of Sha512sum lives in Go's standard library `crypto/sha512`
 * (imported at helpers/helpers.go:22; Go version unpinned, debian/control:11),
 * which is not under /root/reference.  It is restated here from the published
 * algorithm, FIPS 180-4 section 6.4 (SHA-512), with the same streaming shape
 * as Go's digest: 128-byte chunk buffer, running byte length, padding emitted
 * at Sum() time.
 *
 * Parity pinning: tests/test_oracle.py checks this file against all four
 * known-answer digests the reference's own tests hold (SURVEY.md section 8c,
 * items 1-4) and against hashlib / OpenSSL on every length 0..300 and random
 * multi-block lengths.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  Nothing under snappy_b200/ does.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include <pthread.h>
#include <dlfcn.h>
#include <time.h>
#include <sys/stat.h>

/* ------------------------------------------------------------------ */
/* SHA-512 (FIPS 180-4 section 4.2.3 constants, 5.3.5 initial value)   */
/* ------------------------------------------------------------------ */

static const uint64_t K512[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL,
    0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL,
    0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL,
    0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL,
    0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL,
    0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, 0x06ca6351e003826fULL, 0x142929670a0e6e70ULL,
    0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL,
    0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL,
    0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL,
    0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL,
    0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL,
    0xca273eceea26619cULL, 0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL,
    0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL,
    0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL,
};

typedef struct {
    uint64_t h[8];
    uint8_t  x[128];   /* pending partial chunk */
    size_t   nx;       /* bytes in x */
    uint64_t len;      /* total bytes written */
} oracle_sha512_ctx;

static inline uint64_t rotr64(uint64_t v, unsigned n) { return (v >> n) | (v << (64 - n)); }

static inline uint64_t load_be64(const uint8_t *p)
{
    return ((uint64_t)p[0] << 56) | ((uint64_t)p[1] << 48) | ((uint64_t)p[2] << 40) |
           ((uint64_t)p[3] << 32) | ((uint64_t)p[4] << 24) | ((uint64_t)p[5] << 16) |
           ((uint64_t)p[6] << 8) | (uint64_t)p[7];
}

static inline void store_be64(uint8_t *p, uint64_t v)
{
    for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i));
}

/* FIPS 180-4 section 6.4.2, steps 1-4, for `nchunks` consecutive 128-byte chunks. */
static void sha512_chunks(uint64_t h[8], const uint8_t *p, size_t nchunks)
{
    uint64_t w[80];
    while (nchunks--) {
        for (int t = 0; t < 16; t++) w[t] = load_be64(p + 8 * t);
        for (int t = 16; t < 80; t++) {
            uint64_t s0 = rotr64(w[t - 15], 1) ^ rotr64(w[t - 15], 8) ^ (w[t - 15] >> 7);
            uint64_t s1 = rotr64(w[t - 2], 19) ^ rotr64(w[t - 2], 61) ^ (w[t - 2] >> 6);
            w[t] = s1 + w[t - 7] + s0 + w[t - 16];
        }
        uint64_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int t = 0; t < 80; t++) {
            uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
            uint64_t ch = (e & f) ^ (~e & g);
            uint64_t t1 = hh + S1 + ch + K512[t] + w[t];
            uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
            uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint64_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        p += 128;
    }
}

void oracle_sha512_init(oracle_sha512_ctx *c)
{
    static const uint64_t iv[8] = {
        0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
        0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL,
    };
    memcpy(c->h, iv, sizeof iv);
    c->nx = 0;
    c->len = 0;
}

/* Same shape as a Go hash.Hash Write: top up the pending chunk, run whole chunks, keep the rest. */
void oracle_sha512_update(oracle_sha512_ctx *c, const uint8_t *p, size_t n)
{
    c->len += n;
    if (c->nx > 0) {
        size_t take = 128 - c->nx;
        if (take > n) take = n;
        memcpy(c->x + c->nx, p, take);
        c->nx += take;
        p += take;
        n -= take;
        if (c->nx == 128) {
            sha512_chunks(c->h, c->x, 1);
            c->nx = 0;
        }
    }
    if (n >= 128) {
        size_t whole = n / 128;
        sha512_chunks(c->h, p, whole);
        p += whole * 128;
        n -= whole * 128;
    }
    if (n > 0) {
        memcpy(c->x, p, n);
        c->nx = n;
    }
}

/* Padding: 0x80, zeros up to 112 mod 128, then the 128-bit big-endian bit count. */
void oracle_sha512_final(oracle_sha512_ctx *c, uint8_t out[64])
{
    uint64_t len = c->len;
    uint8_t pad[128 + 16];
    memset(pad, 0, sizeof pad);
    pad[0] = 0x80;
    size_t r = (size_t)(len % 128);
    size_t npad = (r < 112) ? (112 - r) : (128 + 112 - r);
    store_be64(pad + npad, len >> 61);
    store_be64(pad + npad + 8, len << 3);
    oracle_sha512_update(c, pad, npad + 16);
    for (int i = 0; i < 8; i++) store_be64(out + 8 * i, c->h[i]);
}

void oracle_sha512(const uint8_t *data, size_t n, uint8_t out[64])
{
    oracle_sha512_ctx c;
    oracle_sha512_init(&c);
    /* helpers.Sha512sum feeds the hasher through io.Copy, i.e. in 32 KiB reads. */
    while (n > 0) {
        size_t take = n > 32768 ? 32768 : n;
        oracle_sha512_update(&c, data, take);
        data += take;
        n -= take;
    }
    oracle_sha512_final(&c, out);
}

static void hex_lower(const uint8_t *in, size_t n, char *out)
{
    static const char d[] = "0123456789abcdef";
    for (size_t i = 0; i < n; i++) {
        out[2 * i] = d[in[i] >> 4];
        out[2 * i + 1] = d[in[i] & 15];
    }
    out[2 * n] = 0;
}

/* helpers.Sha512sum: open, stream, hex.  Returns 0 or -errno. */
int oracle_sha512sum_file(const char *path, char hex_out[129])
{
    int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return -errno;
    oracle_sha512_ctx c;
    oracle_sha512_init(&c);
    static __thread uint8_t buf[32768];
    for (;;) {
        ssize_t r = read(fd, buf, sizeof buf);
        if (r < 0) {
            if (errno == EINTR) continue;
            int e = errno;
            close(fd);
            return -e;
        }
        if (r == 0) break;
        oracle_sha512_update(&c, buf, (size_t)r);
    }
    close(fd);
    uint8_t dg[64];
    oracle_sha512_final(&c, dg);
    hex_lower(dg, 64, hex_out);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Optional OpenSSL block function for the timed CPU baseline          */
/* ------------------------------------------------------------------ */
/* Go's crypto/sha512 uses hand-written amd64 assembly; the closest    */
/* thing in this image is libcrypto's.  Loaded with dlopen so that the */
/* oracle itself has no link-time dependency.                          */

typedef int (*ossl_init_fn)(void *);
typedef int (*ossl_update_fn)(void *, const void *, size_t);
typedef int (*ossl_final_fn)(unsigned char *, void *);
static ossl_init_fn ossl_init;
static ossl_update_fn ossl_update;
static ossl_final_fn ossl_final;
static int ossl_state; /* 0 unknown, 1 ok, -1 unavailable */
static pthread_once_t ossl_once = PTHREAD_ONCE_INIT;

static void ossl_load(void)
{
    const char *names[] = {"libcrypto.so.3", "libcrypto.so", "libcrypto.so.1.1", NULL};
    void *h = NULL;
    for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
    if (h) {
        ossl_init = (ossl_init_fn)dlsym(h, "SHA512_Init");
        ossl_update = (ossl_update_fn)dlsym(h, "SHA512_Update");
        ossl_final = (ossl_final_fn)dlsym(h, "SHA512_Final");
    }
    ossl_state = (ossl_init && ossl_update && ossl_final) ? 1 : -1;
}

int oracle_have_openssl(void)
{
    pthread_once(&ossl_once, ossl_load);
    return ossl_state == 1;
}

static void sha512_one(const uint8_t *data, size_t n, uint8_t out[64], int use_openssl)
{
    if (use_openssl && oracle_have_openssl()) {
        /* SHA512_CTX is 216 bytes in OpenSSL 1.1/3.x; leave generous room. */
        uint64_t ctx[64];
        ossl_init(ctx);
        while (n > 0) {
            size_t take = n > 32768 ? 32768 : n;
            ossl_update(ctx, data, take);
            data += take;
            n -= take;
        }
        ossl_final(out, ctx);
    } else {
        oracle_sha512(data, n, out);
    }
}

/* ------------------------------------------------------------------ */
/* streamsEqual / FilesAreEqual                                        */
/* ------------------------------------------------------------------ */

#define ORACLE_BUFSZ (16 * 1024) /* helpers/cmp.go:27 */

/*
 * streamsEqual over two in-memory streams.  Each round reads up to 16 KiB
 * from both (io.ReadAtLeast semantics: a short final read is
 * ErrUnexpectedEOF, a zero-byte read is EOF); both EOF -> equal; exactly one
 * EOF -> differ; otherwise bytes.Equal on what was read (different lengths
 * compare unequal).
 */
int oracle_streams_equal(const uint8_t *a, size_t la, const uint8_t *b, size_t lb)
{
    size_t pa = 0, pb = 0;
    for (;;) {
        size_t ra = la - pa < ORACLE_BUFSZ ? la - pa : ORACLE_BUFSZ;
        size_t rb = lb - pb < ORACLE_BUFSZ ? lb - pb : ORACLE_BUFSZ;
        int eofa = (ra == 0), eofb = (rb == 0);
        if (eofa && eofb) return 1;
        if (eofa || eofb) return 0;              /* EOF vs data: not "tailMightBeEqual" */
        int shorta = ra < ORACLE_BUFSZ, shortb = rb < ORACLE_BUFSZ;
        if (shorta != shortb) return 0;          /* nil vs ErrUnexpectedEOF */
        if (ra != rb || memcmp(a + pa, b + pb, ra) != 0) return 0;
        pa += ra;
        pb += rb;
    }
}

static ssize_t read_at_least(int fd, uint8_t *buf, size_t want)
{
    size_t got = 0;
    while (got < want) {
        ssize_t r = read(fd, buf + got, want - got);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        if (r == 0) break;
        got += (size_t)r;
    }
    return (ssize_t)got;
}

/* helpers.FilesAreEqual: every failure is "false". */
int oracle_files_are_equal(const char *pa, const char *pb)
{
    int fa = open(pa, O_RDONLY | O_CLOEXEC);
    if (fa < 0) return 0;
    int fb = open(pb, O_RDONLY | O_CLOEXEC);
    if (fb < 0) {
        close(fa);
        return 0;
    }
    struct stat sa, sb;
    int eq = 0;
    if (fstat(fa, &sa) != 0 || fstat(fb, &sb) != 0) goto out;
    if (sa.st_size != sb.st_size) goto out;
    {
        static __thread uint8_t bufa[ORACLE_BUFSZ], bufb[ORACLE_BUFSZ];
        for (;;) {
            ssize_t ra = read_at_least(fa, bufa, ORACLE_BUFSZ);
            ssize_t rb = read_at_least(fb, bufb, ORACLE_BUFSZ);
            if (ra < 0 || rb < 0) goto out;
            if (ra == 0 && rb == 0) {
                eq = 1;
                goto out;
            }
            if (ra == 0 || rb == 0) goto out;
            if ((ra < ORACLE_BUFSZ) != (rb < ORACLE_BUFSZ)) goto out;
            if (ra != rb || memcmp(bufa, bufb, (size_t)ra) != 0) goto out;
        }
    }
out:
    close(fa);
    close(fb);
    return eq;
}

/* ------------------------------------------------------------------ */
/* Batch drivers (serial = what the reference does; threaded = best CPU)*/
/* ------------------------------------------------------------------ */

typedef struct {
    const uint8_t *a, *b;
    const uint64_t *offsets, *lengths;
    size_t lo, hi;
    uint8_t *out;
    int use_openssl;
    int is_cmp;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        if (j->is_cmp)
            j->out[i] = (uint8_t)oracle_streams_equal(j->a + j->offsets[i], j->lengths[i],
                                                      j->b + j->offsets[i], j->lengths[i]);
        else
            sha512_one(j->a + j->offsets[i], j->lengths[i], j->out + 64 * i, j->use_openssl);
    }
    return NULL;
}

static int run_batch(batch_job proto, size_t n, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    if (nthreads == 1) {
        proto.lo = 0;
        proto.hi = n;
        batch_worker(&proto);
        return 0;
    }
    /* contiguous shards balanced by bytes */
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += proto.lengths[i] + 144;
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    batch_job *jobs = (batch_job *)calloc((size_t)nthreads, sizeof *jobs);
    if (!th || !jobs) {
        free(th);
        free(jobs);
        return -ENOMEM;
    }
    size_t pos = 0;
    uint64_t acc = 0;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = proto;
        jobs[t].lo = pos;
        uint64_t goal = total / (uint64_t)nthreads * (uint64_t)(t + 1);
        while (pos < n && (t == nthreads - 1 || acc < goal)) acc += proto.lengths[pos++] + 144;
        jobs[t].hi = pos;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
    return 0;
}

int oracle_sha512_batch(const uint8_t *data, const uint64_t *offsets, const uint64_t *lengths,
                        size_t nfiles, uint8_t *digests, int nthreads, int use_openssl)
{
    batch_job j;
    memset(&j, 0, sizeof j);
    j.a = data;
    j.offsets = offsets;
    j.lengths = lengths;
    j.out = digests;
    j.use_openssl = use_openssl;
    return run_batch(j, nfiles, nthreads);
}

int oracle_cmp_batch(const uint8_t *a, const uint8_t *b, const uint64_t *offsets,
                     const uint64_t *lengths, size_t npairs, uint8_t *equal, int nthreads)
{
    batch_job j;
    memset(&j, 0, sizeof j);
    j.a = a;
    j.b = b;
    j.offsets = offsets;
    j.lengths = lengths;
    j.out = equal;
    j.is_cmp = 1;
    return run_batch(j, npairs, nthreads);
}

/* ------------------------------------------------------------------ */
/* Sha512sum over a file list: the loop of writeHashes on a real tree  */
/* ------------------------------------------------------------------ */
/* snappy/build.go:228-259 calls helpers.Sha512sum for every regular    */
/* file, one goroutine, one file at a time (nthreads = 1).  nthreads >  */
/* 1 is the "best CPU" comparator of SURVEY.md 8(d): the file list      */
/* statically sharded over threads, contiguous shards balanced by       */
/* bytes, each thread running the same open / 32 KiB read / update /    */
/* close loop.  `paths` holds n NUL-terminated strings back to back,    */
/* `sizes` their lengths in bytes (for the balancing only).             */

typedef struct {
    const char **paths;
    size_t lo, hi;
    uint8_t *out;
    int use_openssl;
    int err;           /* -errno of the first failure of this shard */
} files_job;

static int sha512sum_fd_loop(const char *path, uint8_t out[64], int use_openssl)
{
    int fd = open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return -errno;
    static __thread uint8_t buf[32768];      /* io.Copy's buffer */
    oracle_sha512_ctx c;
    uint64_t octx[64];
    const int ossl = use_openssl && oracle_have_openssl();
    if (ossl) ossl_init(octx);
    else oracle_sha512_init(&c);
    for (;;) {
        ssize_t r = read(fd, buf, sizeof buf);
        if (r < 0) {
            if (errno == EINTR) continue;
            int e = errno;
            close(fd);
            return -e;
        }
        if (r == 0) break;
        if (ossl) ossl_update(octx, buf, (size_t)r);
        else oracle_sha512_update(&c, buf, (size_t)r);
    }
    close(fd);
    if (ossl) ossl_final(out, octx);
    else oracle_sha512_final(&c, out);
    return 0;
}

static void *files_worker(void *arg)
{
    files_job *j = (files_job *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        int rc = sha512sum_fd_loop(j->paths[i], j->out + 64 * i, j->use_openssl);
        if (rc && !j->err) j->err = rc;
    }
    return NULL;
}

int oracle_sha512sum_files(const char *paths, const uint64_t *sizes, size_t n, uint8_t *digests,
                           int nthreads, int use_openssl)
{
    if (n == 0) return 0;
    const char **list = (const char **)calloc(n, sizeof *list);
    if (!list) return -ENOMEM;
    const char *p = paths;
    for (size_t i = 0; i < n; i++) {
        list[i] = p;
        p += strlen(p) + 1;
    }
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = (int)n;
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) total += sizes[i] + 4096;     /* per-file cost: open/close */
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    files_job *jobs = (files_job *)calloc((size_t)nthreads, sizeof *jobs);
    if (!th || !jobs) {
        free(list);
        free(th);
        free(jobs);
        return -ENOMEM;
    }
    size_t pos = 0;
    uint64_t acc = 0;
    for (int t = 0; t < nthreads; t++) {
        jobs[t].paths = list;
        jobs[t].out = digests;
        jobs[t].use_openssl = use_openssl;
        jobs[t].lo = pos;
        uint64_t goal = total / (uint64_t)nthreads * (uint64_t)(t + 1);
        while (pos < n && (t == nthreads - 1 || acc < goal)) acc += sizes[pos++] + 4096;
        jobs[t].hi = pos;
        if (nthreads == 1) files_worker(&jobs[t]);
        else pthread_create(&th[t], NULL, files_worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        if (jobs[t].err && !rc) rc = jobs[t].err;
    }
    free(list);
    free(th);
    free(jobs);
    return rc;
}

/* bytes the reference's streamsEqual touches for one pair:
 * 2*L for an equal pair, 2*16384*(floor(first_diff/16384)+1) (capped at 2*L) otherwise. */
uint64_t oracle_cmp_algorithmic_bytes(const uint8_t *a, const uint8_t *b, uint64_t len)
{
    for (uint64_t pos = 0; pos < len; pos += ORACLE_BUFSZ) {
        uint64_t n = len - pos < ORACLE_BUFSZ ? len - pos : ORACLE_BUFSZ;
        if (memcmp(a + pos, b + pos, n) != 0) return 2 * (pos + n);
    }
    return 2 * len;
}

double oracle_now_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
