// io_probe -- per-file cost of scan + open/fstat/read/close on a tree, by thread count and strategy
// (mode 0 lstat+open by path, 1 fstatat+openat, 2 openat+fstat+O_NOATIME, 3 openat only), with
// UNSHARE=1 giving every thread a private descriptor table, NOREAD=1 skipping the read and DEFER=n closing n
// descriptors with one close_range call (only meaningful with UNSHARE=1).
//   g++ -O2 -std=c++17 -o io_probe io_probe.cpp -lpthread;  io_probe TREE THREADS MODE
// probe: per-file cost of open/read/close on tmpfs with several strategies
#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <linux/close_range.h>
#include <sys/wait.h>
#include <time.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec*1e3 + ts.tv_nsec*1e-6; }
struct linux_dirent64 { ino64_t d_ino; off64_t d_off; unsigned short d_reclen; unsigned char d_type; char d_name[]; };
int main(int argc, char **argv) {
    std::string root = argv[1];
    int nthreads = atoi(argv[2]);
    int mode = atoi(argv[3]);
    // list dirs
    std::vector<std::string> dirs;
    { DIR *d = opendir(root.c_str()); while (dirent *e = readdir(d)) if (e->d_name[0] != '.') dirs.push_back(e->d_name); closedir(d); }
    std::sort(dirs.begin(), dirs.end());
    size_t cap = (size_t)2 << 30;
    uint8_t *buf = (uint8_t*)mmap(nullptr, cap, PROT_READ|PROT_WRITE, MAP_PRIVATE|MAP_ANONYMOUS|MAP_POPULATE, -1, 0);
    for (int rep = 0; rep < 3; rep++) {
    std::atomic<size_t> next{0}, bump{0}, files{0}, bytes{0};
    double t0 = now();
    auto work = [&]() {
        if (getenv("UNSHARE")) syscall(SYS_close_range, 3, ~0U, CLOSE_RANGE_UNSHARE);
        std::vector<char> dbuf(1 << 20);
        size_t nf = 0, nb = 0;
        for (size_t k; (k = next.fetch_add(1)) < dirs.size();) {
            std::string dp = root + "/" + dirs[k];
            int dfd = open(dp.c_str(), O_RDONLY|O_DIRECTORY|O_CLOEXEC);
            std::vector<std::pair<std::string, unsigned char>> names;
            for (;;) { long n = syscall(SYS_getdents64, dfd, dbuf.data(), dbuf.size()); if (n <= 0) break;
                for (long p = 0; p < n;) { auto *e = (linux_dirent64*)(dbuf.data()+p); p += e->d_reclen; if (e->d_name[0]=='.' && (!e->d_name[1] || (e->d_name[1]=='.'&&!e->d_name[2]))) continue; names.emplace_back(e->d_name, e->d_type);} }
            std::sort(names.begin(), names.end());
            size_t chunk = 0, chunk_end = 0;
            const int defer = getenv("DEFER") ? atoi(getenv("DEFER")) : 0;
            int lo_fd = -1, hi_fd = -1, held = 0;
            for (auto &nm : names) {
                struct stat st;
                int fd;
                if (mode == 0) { // lstat by full path, open by full path
                    std::string fp = dp + "/" + nm.first;
                    lstat(fp.c_str(), &st);
                    fd = open(fp.c_str(), O_RDONLY|O_CLOEXEC);
                } else if (mode == 1) { // fstatat + openat
                    fstatat(dfd, nm.first.c_str(), &st, AT_SYMLINK_NOFOLLOW);
                    fd = openat(dfd, nm.first.c_str(), O_RDONLY|O_CLOEXEC);
                } else if (mode == 2) { // openat + fstat
                    fd = openat(dfd, nm.first.c_str(), O_RDONLY|O_CLOEXEC|O_NOFOLLOW|O_NOATIME);
                    fstat(fd, &st);
                } else { // openat only, read with big buffer (no stat)
                    fd = openat(dfd, nm.first.c_str(), O_RDONLY|O_CLOEXEC|O_NOFOLLOW|O_NOATIME);
                    st.st_size = 65536;
                }
                size_t need = (st.st_size + 1 + 15) & ~15ull;
                if (chunk + need > chunk_end) { chunk = bump.fetch_add(4 << 20); chunk_end = chunk + (4 << 20); }
                ssize_t r = getenv("NOREAD") ? st.st_size : read(fd, buf + chunk, st.st_size + 1);
                if (mode == 3) need = (r + 15) & ~15ull;
                chunk += need;
                if (!defer) close(fd);
                else {
                    if (!held) lo_fd = hi_fd = fd; else { lo_fd = std::min(lo_fd, fd); hi_fd = std::max(hi_fd, fd); }
                    if (++held == defer) { syscall(SYS_close_range, (unsigned)lo_fd, (unsigned)hi_fd, 0); held = 0; }
                }
                nf++; nb += r;
            }
            if (held) syscall(SYS_close_range, (unsigned)lo_fd, (unsigned)hi_fd, 0);
            close(dfd);
        }
        files += nf; bytes += nb;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work);
    work();
    for (auto &x : th) x.join();
    double ms = now() - t0;
    printf("mode %d threads %d: %zu files %zu bytes in %.2f ms = %.2f us/file/thread, %.2f GB/s\n", mode, nthreads, files.load(), bytes.load(), ms, ms*1e3*nthreads/files.load(), bytes.load()/ms/1e6);
    }
}
