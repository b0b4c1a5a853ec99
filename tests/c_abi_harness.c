#define _GNU_SOURCE
/*
 * c_abi_harness.c -- a plain C caller of libsnapgpu, linked against the shared library the way
 * the cgo shim of INTEGRATION.md is (idiom: helpers/touch.go:20-57 -- C strings in, int status
 * out, errno-style text on failure).  It rebuilds the tree of TestBuildCreateDebianHashesSimple
 * (snappy/hashes_test.go:57-87) with libc calls, runs writeHashes through the C ABI and compares
 * DEBIAN/hashes.yaml with the reference's golden document (snappy/hashes_test.go:89-103), then
 * checks the reference's other known answers through the one-file entry points.
 *
 *   c_abi_harness WORKDIR GOLDEN_YAML      exit 0 = all checks passed
 *   c_abi_harness --no-gpu                 exit 0 = the library refuses to compute without CUDA
 */
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../include/snapgpu.h"

static int failures = 0;
#define CHECK(cond, ...)                          \
    do {                                          \
        if (!(cond)) {                            \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);         \
            fprintf(stderr, "\n");                \
            failures++;                           \
        }                                         \
    } while (0)

static void put_file(const char *dir, const char *name, const char *body, mode_t mode) {
    char path[4096];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, mode);
    if (fd < 0 || write(fd, body, strlen(body)) != (ssize_t)strlen(body)) {
        perror(path);
        exit(2);
    }
    fchmod(fd, mode);
    close(fd);
}

static char *slurp(const char *path, size_t *len) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    rewind(f);
    char *buf = malloc((size_t)n + 1);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) n = 0;
    buf[n] = 0;
    fclose(f);
    *len = (size_t)n;
    return buf;
}

int main(int argc, char **argv) {
    if (argc == 2 && !strcmp(argv[1], "--no-gpu")) {
        /* no CPU fallback: without a device init fails and so does every compute call */
        int rc = snapgpu_init(NULL, 0);
        if (rc == 0) {
            printf("a CUDA device is present; nothing to check\n");
            return 0;
        }
        uint8_t d[64], x = 'x';
        uint64_t off = 0, len = 1;
        CHECK(rc == SNAPGPU_ECUDA, "init returned %d", rc);
        CHECK(strstr(snapgpu_last_error(), "no CPU fallback") != NULL, "message: %s", snapgpu_last_error());
        CHECK(snapgpu_sha512_batch(&x, &off, &len, 1, d) < 0, "sha512_batch computed something without a GPU");
        CHECK(snapgpu_files_are_equal("/etc/hostname", "/etc/hostname") == 0, "files_are_equal answered without a GPU");
        printf("%s\n", failures ? "FAILED" : "ok: refused without CUDA");
        return failures ? 1 : 0;
    }
    if (argc != 3) {
        fprintf(stderr, "usage: %s WORKDIR GOLDEN_YAML | --no-gpu\n", argv[0]);
        return 2;
    }
    const char *work = argv[1];
    char tree[4000], path[4096], tar[4096];
    snprintf(tree, sizeof tree, "%s/tree", work);
    mkdir(work, 0755);
    mkdir(tree, 0755);
    snprintf(path, sizeof path, "%s/DEBIAN", tree);
    mkdir(path, 0755);
    put_file(path, "bar", "", 0644);
    put_file(tree, "foo", "", 0644);
    snprintf(path, sizeof path, "%s/bin", tree);
    mkdir(path, 0755);
    chmod(path, 0755);
    put_file(path, "bar", "bar\n", 0644);
    snprintf(path, sizeof path, "%s/broken-link", tree);
    unlink(path);
    if (symlink("/dsafdsafsadf", path) != 0) {
        perror("symlink");
        return 2;
    }
    snprintf(tar, sizeof tar, "%s/data.tar.gz", work);
    put_file(work, "data.tar.gz", "", 0644);

    CHECK(snapgpu_init(NULL, 1) == 0, "snapgpu_init: %s", snapgpu_last_error());
    CHECK(snapgpu_num_devices() == 1, "bound %d devices", snapgpu_num_devices());

    /* writeHashes(buildDir, dataTar) -- snappy/build.go:216 */
    int rc = snapgpu_write_hashes(tree, tar);
    CHECK(rc == 0, "snapgpu_write_hashes: %d %s", rc, snapgpu_last_error());
    size_t got_len = 0, want_len = 0;
    snprintf(path, sizeof path, "%s/DEBIAN/hashes.yaml", tree);
    char *got = slurp(path, &got_len), *want = slurp(argv[2], &want_len);
    CHECK(got && want, "cannot read %s or %s", path, argv[2]);
    if (got && want) CHECK(got_len == want_len && !memcmp(got, want, got_len), "hashes.yaml differs from the golden document:\n%s", got);
    struct stat st;
    CHECK(stat(path, &st) == 0 && (st.st_mode & 0777) == 0644, "hashes.yaml mode %o", (unsigned)(st.st_mode & 0777));

    /* the same document through the buffer-returning form */
    char *doc = NULL;
    size_t doc_len = 0;
    rc = snapgpu_hashes_yaml(tree, tar, &doc, &doc_len);
    CHECK(rc == 0 && doc && want && doc_len == want_len && !memcmp(doc, want, doc_len), "snapgpu_hashes_yaml: %d %s", rc, snapgpu_last_error());
    snapgpu_free(doc);

    /* helpers.Sha512sum -- helpers/helpers_test.go:171-175 */
    put_file(work, "x", "x", 0644);
    snprintf(path, sizeof path, "%s/x", work);
    char hex[129];
    rc = snapgpu_sha512sum_file(path, hex);
    CHECK(rc == 0 && !strcmp(hex, "a4abd4448c49562d828115d13a1fccea927f52b4d5459297f8b43e42da89238bc13626e43dcb38ddb082488927ec904fb42057443983e88585179d50551afe62"),
          "Sha512sum(\"x\") = %s (%s)", hex, snapgpu_last_error());
    snprintf(path, sizeof path, "%s/missing", work);
    rc = snapgpu_sha512sum_file(path, hex);
    CHECK(rc == SNAPGPU_EIO && hex[0] == 0 && strstr(snapgpu_last_error(), "no such file or directory") == NULL &&
              strstr(snapgpu_last_error(), "No such file or directory") != NULL,
          "missing file: %d '%s'", rc, snapgpu_last_error());

    /* helpers.FilesAreEqual -- helpers/cmp_test.go:29-63 */
    char a[4096], b[4096];
    snprintf(a, sizeof a, "%s/a", work);
    snprintf(b, sizeof b, "%s/b", work);
    put_file(work, "a", "same content", 0644);
    put_file(work, "b", "same content", 0644);
    CHECK(snapgpu_files_are_equal(a, b) == 1, "equal files reported different");
    put_file(work, "b", "same cOntent", 0644);
    CHECK(snapgpu_files_are_equal(a, b) == 0, "different files reported equal");
    CHECK(snapgpu_files_are_equal(a, path) == 0, "a missing file is 'not equal'");

    /* the batch entry point with C-owned pinned memory, as the Go packer uses it */
    uint8_t *pin = snapgpu_alloc_pinned(4096);
    CHECK(pin != NULL, "snapgpu_alloc_pinned: %s", snapgpu_last_error());
    if (pin) {
        memset(pin, 0, 4096);
        memcpy(pin + 16, "bar\n", 4);
        const uint64_t offs[2] = {0, 16}, lens[2] = {0, 4};
        uint8_t dg[128];
        rc = snapgpu_sha512_batch(pin, offs, lens, 2, dg);
        static const uint8_t empty8[8] = {0xcf, 0x83, 0xe1, 0x35, 0x7e, 0xef, 0xb8, 0xbd}, bar8[8] = {0xcc, 0x06, 0x80, 0x8c, 0xbb, 0xee, 0x05, 0x10};
        CHECK(rc == 0 && !memcmp(dg, empty8, 8) && !memcmp(dg + 64, bar8, 8), "sha512_batch: %d %s", rc, snapgpu_last_error());
        snapgpu_free_pinned(pin);
    }
    snapgpu_stats s;
    CHECK(snapgpu_get_stats(&s) == 0 && s.sha512_launches >= 3, "kernel launches: %llu", (unsigned long long)s.sha512_launches);
    snapgpu_shutdown();
    free(got);
    free(want);
    printf("%s\n", failures ? "FAILED" : "ok");
    return failures ? 1 : 0;
}
