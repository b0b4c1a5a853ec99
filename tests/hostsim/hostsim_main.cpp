// hostsim -- drives the host side of libsnapgpu against the fake runtime (see fake_runtime.cpp).
//   hostsim hashes_yaml DIR TAR            document to stdout
//   hostsim write_hashes DIR TAR
//   hostsim verify ROOT YAML [TAR]         report lines to stdout
//   hostsim copy SRC DST FLAGS [TAR]       copyToBuildDir, then (with TAR) writeHashes to stdout
//   hostsim dir_updated A B PFX            names to stdout
//   hostsim sha512sum FILE
//   hostsim repeat N DIR TAR               N documents must be identical (thread-order independence)
// exit code 0 on success; on a library error prints "ERR <code> <message>" and exits 3.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "../../include/snapgpu.h"

// HOSTSIM_WATCHDOG=seconds: when the run takes longer, every thread prints its backtrace and
// the process aborts (debugging aid for the threaded pipeline; needs SNAPGPU_SHARED_FDS=1 so
// that the pool threads still have a stderr)
#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <execinfo.h>
#include <signal.h>
#include <sys/syscall.h>
#include <unistd.h>
static void dump_handler(int) {
    void *frames[48];
    const int n = backtrace(frames, 48);
    char head[64];
    const int m = snprintf(head, sizeof head, "---- thread %ld\n", (long)syscall(SYS_gettid));
    if (write(2, head, (size_t)m) < 0) return;
    backtrace_symbols_fd(frames, n, 2);
}
static void alarm_handler(int) {
    DIR *d = opendir("/proc/self/task");
    const long self = (long)syscall(SYS_gettid);
    while (struct dirent *e = readdir(d)) {
        const long tid = atol(e->d_name);
        if (tid > 0 && tid != self) {
            syscall(SYS_tgkill, getpid(), tid, SIGUSR2);
            usleep(20000);
        }
    }
    dump_handler(0);
    _exit(97);
}

static int die(int rc) {
    printf("ERR %d %s\n", rc, snapgpu_last_error());
    return 3;
}
static void print_list(char *names, size_t n) {
    const char *p = names;
    for (size_t i = 0; i < n; i++) {
        puts(p);
        p += strlen(p) + 1;
    }
    snapgpu_free(names);
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (const char *w = getenv("HOSTSIM_WATCHDOG")) {
        signal(SIGUSR2, dump_handler);
        signal(SIGALRM, alarm_handler);
        alarm((unsigned)atoi(w));
    }
    const std::string cmd = argv[1];
    char *out = nullptr;
    size_t len = 0;
    int rc = 0;
    if (cmd == "hashes_yaml" && argc == 4) {
        if ((rc = snapgpu_hashes_yaml(argv[2], argv[3], &out, &len))) return die(rc);
        fwrite(out, 1, len, stdout);
        snapgpu_free(out);
    } else if (cmd == "repeat" && argc == 5) {
        std::string first;
        for (int i = 0; i < atoi(argv[2]); i++) {
            if ((rc = snapgpu_hashes_yaml(argv[3], argv[4], &out, &len))) return die(rc);
            std::string doc(out, len);
            snapgpu_free(out);
            if (i == 0) first = doc;
            else if (doc != first) {
                printf("ERR 0 run %d differs from run 0\n", i);
                return 3;
            }
        }
        fwrite(first.data(), 1, first.size(), stdout);
    } else if (cmd == "warm" && argc == 4) {
        // snapgpu_warm from a second thread while the first already hashes the tree (a build's goroutine may lose the race)
        int warm_rc = 0;
        std::thread warmer([&] { warm_rc = snapgpu_warm(); });
        rc = snapgpu_hashes_yaml(argv[2], argv[3], &out, &len);
        warmer.join();
        if (rc) return die(rc);
        if (warm_rc) return die(warm_rc);
        std::string first(out, len);
        snapgpu_free(out);
        if ((rc = snapgpu_warm())) return die(rc);
        if ((rc = snapgpu_hashes_yaml(argv[2], argv[3], &out, &len))) return die(rc);
        if (first != std::string(out, len)) {
            printf("ERR 0 the document after the warm-up differs\n");
            return 3;
        }
        fwrite(out, 1, len, stdout);
        snapgpu_free(out);
    } else if (cmd == "write_hashes" && argc == 4) {
        if ((rc = snapgpu_write_hashes(argv[2], argv[3]))) return die(rc);
    } else if (cmd == "verify" && (argc == 4 || argc == 5)) {
        if ((rc = snapgpu_verify_hashes(argv[2], argv[3], argc == 5 ? argv[4] : nullptr, &out, &len))) return die(rc);
        print_list(out, len);
    } else if (cmd == "copy" && (argc == 5 || argc == 6)) {
        if ((rc = snapgpu_copy_to_build_dir(argv[2], argv[3], atoi(argv[4])))) return die(rc);
        if (argc == 6) {
            if ((rc = snapgpu_hashes_yaml(argv[3], argv[5], &out, &len))) return die(rc);
            fwrite(out, 1, len, stdout);
            snapgpu_free(out);
        }
    } else if (cmd == "copy_edit" && argc == 6) {
        // copyToBuildDir (copy forced), then every byte of DST/<argv[5]> is inverted in place and its
        // mtime put back, then writeHashes: the digest cache must not serve the stale digest
        if ((rc = snapgpu_copy_to_build_dir(argv[2], argv[3], SNAPGPU_COPY_NO_LINK))) return die(rc);
        const std::string victim = std::string(argv[3]) + "/" + argv[5];
        struct stat st;
        if (stat(victim.c_str(), &st) != 0) return 4;
        FILE *f = fopen(victim.c_str(), "r+b");
        if (!f) return 4;
        std::string body((size_t)st.st_size, '\0');
        if (fread(&body[0], 1, body.size(), f) != body.size()) return 4;
        for (char &c : body) c = (char)~c;
        rewind(f);
        fwrite(body.data(), 1, body.size(), f);
        fclose(f);
        struct timespec times[2] = {st.st_atim, st.st_mtim};
        if (utimensat(AT_FDCWD, victim.c_str(), times, 0) != 0) return 4;
        if ((rc = snapgpu_hashes_yaml(argv[3], argv[4], &out, &len))) return die(rc);
        fwrite(out, 1, len, stdout);
        snapgpu_free(out);
    } else if (cmd == "dir_updated" && argc == 5) {
        if ((rc = snapgpu_dir_updated(argv[2], argv[3], argv[4], &out, &len))) return die(rc);
        print_list(out, len);
    } else if (cmd == "sha512sum" && argc == 3) {
        char hex[129];
        if ((rc = snapgpu_sha512sum_file(argv[2], hex))) return die(rc);
        puts(hex);
    } else {
        return 2;
    }
    return 0;
}
