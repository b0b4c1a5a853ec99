// tree_hasher.hpp -- writeHashes (snappy/build.go:216-270) as a streaming pipeline.
// Included by host_path.cpp only, inside its anonymous namespace, after the YAML emitter, the
// digest cache and the Go-style error helpers it uses.
//
// The reference walks the tree with one goroutine and reads and hashes one file at a time
// (build.go:228-259).  Here four stages overlap:
//
//   scan    pool workers read directories (getdents64), sort the names the way filepath.Walk's
//           readDirNames does, lstat what is not a regular file and queue sub-directories;
//   pack    pool workers open/fstat/read/close every regular file straight into pinned chunk
//           buffers (16-byte aligned slots, ONE read per file);
//   hash    the calling thread hands finished chunks to a batch session (runtime.hpp): H2D copy,
//           length binning, SHA-512 kernel and digest copy-back run behind it, a chunk returns
//           to the pool as soon as its copy is done;
//   chains  the archive (build.go:222) and every file too long for a batch are single SHA-512
//           chains; a streamer thread advances all of them together, piece by piece, beside the
//           batches, starting with the archive before the walk begins.
//
// and the document is written afterwards by the pool in one parallel pass.  Results are
// assembled in filepath.Walk order, errors are reported as the reference would meet them: the
// archive first, then the first failing entry in walk order, then the first unknown mode.
//
// The pool's threads have private descriptor tables (close_range(CLOSE_RANGE_UNSHARE)): opening
// and closing a hundred thousand files from sixteen threads of one process otherwise serialises
// on the process-wide table's lock (measured on tmpfs: 2 threads no faster than 1, 8 threads
// 2.1x; with private tables 1.9x and 5.5x).  Such a thread must not touch descriptors owned by
// the rest of the process -- so the workers never call CUDA: pinned chunks are allocated and
// every GPU call is made by ordinary threads.
// (no #include here: the file is included inside a namespace; host_path.cpp includes
// <condition_variable>, <deque> and <sys/syscall.h> for it)
#pragma once

#ifndef CLOSE_RANGE_UNSHARE
#define CLOSE_RANGE_UNSHARE (1U << 1)
#endif

// ------------------------------------------------------------------------------------------
// I/O pool
// ------------------------------------------------------------------------------------------

class IoPool {
public:
    static IoPool &instance() {
        static IoPool *p = new IoPool();        // never destroyed: its threads outlive static destructors
        return *p;
    }
    unsigned size() const { return nthreads_; }

    // fn(worker) runs on `workers` pool threads; `meanwhile` (optional) runs on the calling
    // thread beside them.  Returns when all have returned.  One job at a time.
    void run(unsigned workers, const std::function<void(unsigned)> &fn, const std::function<void()> &meanwhile = nullptr) {
        std::lock_guard<std::mutex> one(user_mu_);
        workers = std::max(1u, std::min(workers, nthreads_));
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &fn;
            job_workers_ = workers;
            running_ = workers;
            seq_++;
        }
        cv_work_.notify_all();
        if (meanwhile) meanwhile();
        std::unique_lock<std::mutex> lk(mu_);
        cv_done_.wait(lk, [&] { return running_ == 0; });
        job_ = nullptr;
    }

private:
    IoPool() {
        unsigned n = 0;
        if (const char *e = getenv("SNAPGPU_PACK_THREADS")) n = (unsigned)atoi(e);
        if (!n) n = std::max(2u, std::thread::hardware_concurrency());
        nthreads_ = std::min(n, 128u);
        const bool share = getenv("SNAPGPU_SHARED_FDS") != nullptr;
        for (unsigned t = 0; t < nthreads_; t++) std::thread([this, t, share] { main(t, share); }).detach();
    }
    void main(unsigned index, bool share) {
        // a private, empty descriptor table (see the header comment); harmless if unsupported
        if (!share) syscall(SYS_close_range, 0u, ~0u, CLOSE_RANGE_UNSHARE);
        // One worker per core keeps every core busy, and the thread that drives the GPU -- it sleeps
        // on 50 us timers and has a few microseconds of work each time -- then waits for a worker's
        // time slice to end before it runs (measured: the first batch's completion was noticed 4 ms
        // late, the batches behind it piled up).  The workers give way to it.
        setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 10);
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(unsigned)> *job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return seq_ != seen; });
                seen = seq_;
                if (index >= job_workers_) continue;
                job = job_;
            }
            (*job)(index);
            {
                std::lock_guard<std::mutex> lk(mu_);
                running_--;
            }
            cv_done_.notify_all();
        }
    }
    unsigned nthreads_ = 0;
    std::mutex user_mu_, mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(unsigned)> *job_ = nullptr;
    unsigned job_workers_ = 0, running_ = 0;
    uint64_t seq_ = 0;
};

// ------------------------------------------------------------------------------------------
// pinned chunks
// ------------------------------------------------------------------------------------------

// 4 MiB: a span of that size copies at 52.5 GB/s, one of 2 MiB at 49.9, of 1 MiB at 45.5 (64 MiB: 55.3;
// profiles/r02_h2d_pieces.jsonl -- where sixteen threads streaming through DRAM, as the packers do,
// cost the copy engine another 27 %, whatever the span size and however many copy streams)
constexpr size_t kSmallChunk = (size_t)4 << 20;      // files up to kSmallMax are packed into these
constexpr size_t kSmallMax = (size_t)256 << 10;
constexpr size_t kLargeChunk = (size_t)32 << 20;     // files up to kMidMax, a few per chunk
constexpr size_t kMidMax = (size_t)16 << 20;         // longer files are chains of their own (ChainStreamer)
constexpr size_t kChunkSlack = 256;
constexpr size_t kTreeBatchBytes = (size_t)64 << 20; // one batch of the tree hasher: at least two large chunks
constexpr size_t kHungryFlush = (size_t)384 << 10;   // a worker hands over a chunk this full when the GPU is idle

struct PackedRef {
    uint8_t *digest;       // where the digest goes (a TEntry's)
    uint32_t off;          // slot in the chunk
    uint32_t len;
};

struct Chunk {
    uint8_t *base = nullptr;
    size_t cap = 0, used = 0;
    int cls = 0;           // 0 small, 1 large
    std::vector<PackedRef> files;
};

// The chunks of one size class.  They stay pinned for the life of the process (pinning costs
// ~0.4 ms per MiB, which a warm caller should not pay again); only as many as a tree needs are
// ever allocated.  Workers take and return chunks; allocation needs CUDA and is done for them by
// the thread that drives the session (serve_allocations).
class ChunkPool {
public:
    ChunkPool(size_t chunk_bytes, size_t max_chunks, size_t first_slab, int cls)
        : bytes_(chunk_bytes), max_(max_chunks), first_slab_(first_slab), cls_(cls) {}

    // worker side: a free chunk or nullptr.  A worker that got none announces itself with
    // begin_wait() and keeps trying; the driver allocates for the waiters it sees (up to the cap),
    // so a chunk taken by somebody else in between is simply made up for on its next round.
    Chunk *try_get() {
        std::lock_guard<std::mutex> lk(mu_);
        if (free_.empty()) return nullptr;
        Chunk *c = free_.back();
        free_.pop_back();
        return c;
    }
    void begin_wait() {
        std::lock_guard<std::mutex> lk(mu_);
        waiters_++;
    }
    void end_wait() {
        std::lock_guard<std::mutex> lk(mu_);
        waiters_--;
    }
    void put(Chunk *c) {
        c->used = 0;
        c->files.clear();
        std::lock_guard<std::mutex> lk(mu_);
        free_.push_back(c);
    }
    // driver side: starts growing the pool when workers wait for chunks; returns -1 when they wait, there is not a
    // single chunk to recycle for them and pinned memory cannot be had at all, else 0.  The pool grows in slabs that
    // double it (at least `first_slab` chunks): a handful of allocations during the first large tree of a process,
    // none afterwards.  The allocation itself runs on a thread of its own (pinning 256 MiB takes 10-50 ms, 115 ms
    // through cudaHostAlloc): the driver keeps submitting and recycling meanwhile, and the workers, which poll the
    // pool every 150 us while they wait, pick the new chunks up as they appear.  The pool outlives every tree
    // (process-lifetime singleton), so a slab that arrives after its tree has finished is simply there for the next.
    int serve_allocations() {
        size_t want = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (failed_ && all_.empty() && waiters_ > 0) return -1;
            if (!growing_ && !failed_ && waiters_ > free_.size() && all_.size() < max_) {
                // doubling, but no slab larger than 128 MiB: registering a slab holds up the process's other CUDA calls
                // for a few milliseconds per 64 MiB, and a smaller slab is in the workers' hands sooner
                const size_t slab_cap = std::max<size_t>(first_slab_, ((size_t)128 << 20) / bytes_);
                want = std::min({std::max(first_slab_, all_.size()), slab_cap, max_ - all_.size()});
                growing_ = want != 0;
            }
        }
        if (!want) return 0;
        std::thread([this, want]() mutable {
            uint8_t *p = nullptr;
            for (; want; want /= 2)              // a smaller slab if the large one cannot be pinned
                if ((p = static_cast<uint8_t *>(snapgpu_alloc_pinned(want * bytes_)))) break;
            if (getenv("SNAPGPU_TRACE"))
                fprintf(stderr, "[snapgpu] chunk pool %d: +%zu chunks of %zu MiB (%s)\n", cls_, p ? want : (size_t)0, bytes_ >> 20,
                        p ? "pinned" : "no pinned memory");
            std::lock_guard<std::mutex> lk(mu_);
            growing_ = false;
            failed_ = p == nullptr;              // no more growing; the workers keep waiting for recycled chunks
            for (size_t k = 0; p && k < want; k++) {
                Chunk *c = new Chunk();
                c->base = p + k * bytes_;
                c->cap = bytes_;
                c->cls = cls_;
                all_.push_back(c);
                free_.push_back(c);
            }
        }).detach();
        return 0;
    }
    size_t allocated() {
        std::lock_guard<std::mutex> lk(mu_);
        return all_.size();
    }
    // Grow the pool to at least n chunks now, on the calling thread (snapgpu_warm: ahead of the first tree).
    bool reserve(size_t n) {
        for (;;) {
            size_t want;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (all_.size() >= std::min(n, max_)) return true;
                if (failed_) return false;
                if (growing_) want = 0;              // a tree is growing it at this moment: let that slab land first
                else {
                    want = std::min(std::min(n, max_) - all_.size(), std::max<size_t>(first_slab_, ((size_t)128 << 20) / bytes_));
                    growing_ = true;
                }
            }
            if (!want) {
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
                continue;
            }
            uint8_t *p = static_cast<uint8_t *>(snapgpu_alloc_pinned(want * bytes_));
            std::lock_guard<std::mutex> lk(mu_);
            growing_ = false;
            failed_ = p == nullptr;
            for (size_t k = 0; p && k < want; k++) {
                Chunk *c = new Chunk();
                c->base = p + k * bytes_;
                c->cap = bytes_;
                c->cls = cls_;
                all_.push_back(c);
                free_.push_back(c);
            }
        }
    }

private:
    const size_t bytes_, max_, first_slab_;
    const int cls_;
    std::mutex mu_;
    std::vector<Chunk *> all_, free_;
    size_t waiters_ = 0;
    bool growing_ = false, failed_ = false;
};

inline ChunkPool &small_chunks() {
    static ChunkPool *p = new ChunkPool(kSmallChunk, 192, 8, 0);   // 32 MiB, then slabs of up to 128 MiB, up to 768 MiB pinned
    return *p;
}
inline ChunkPool &large_chunks() {
    static ChunkPool *p = new ChunkPool(kLargeChunk, 24, 2, 1);    // 64 MiB, then slabs of 128 MiB, up to 768 MiB pinned
    return *p;
}

// ------------------------------------------------------------------------------------------
// chains: the archive and the files too long for a batch
// ------------------------------------------------------------------------------------------

// One SHA-512 chain advances at ~70 MB/s on the GPU however idle the rest of it is, and any
// number of chains advance side by side at that speed (one lane pair each, sha512_pair.cuh).
// The streamer therefore moves all its chains forward together: every round reads the next
// piece of each open file into pinned memory and hashes the pieces as continuation segments of
// one call, on a pipe of its own beside the tree's batches; the pieces of round r+1 are read
// while the GPU works on round r.
class ChainStreamer {
public:
    struct Chain {
        std::string path;
        uint8_t *digest_out = nullptr;     // 64 bytes, written when the chain ends
        int *err_out = nullptr;            // errno of a failed open/read ...
        uint8_t *op_out = nullptr;         // ... and which (2 open, 3 read)
        int fd = -1;
        uint64_t prefix = 0;
        bool eof = false;
        uint8_t state[64] = {0};
    };

    ~ChainStreamer() { finish(); }

    // Called by the thread that owns the streamer (an ordinary thread: the streamer makes GPU
    // calls, so it must not be started from a pool worker, whose descriptor table is private).
    void start() {
        if (!th_.joinable()) th_ = std::thread([this] { main(); });
    }
    // Any thread, pool workers included.
    void add(const std::string &path, uint8_t *digest_out, int *err_out, uint8_t *op_out) {
        Chain *c = new Chain();
        c->path = path;
        c->digest_out = digest_out;
        c->err_out = err_out;
        c->op_out = op_out;
        {
            std::lock_guard<std::mutex> lk(mu_);
            incoming_.push_back(c);
        }
        cv_.notify_all();
    }
    // no more chains will be added; waits for all of them.  Returns the first GPU-side failure.
    int finish() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            closing_ = true;
        }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
        if (rc_) return fail(rc_, "%s", err_.c_str());
        return 0;
    }
    double busy_ms() const { return busy_ms_; }

private:
    static constexpr size_t kPiece = (size_t)2 << 20;      // per chain and round: ~30 ms of chain
    static constexpr size_t kMaxChains = 32;

    struct Buffers {
        std::mutex mu;
        uint8_t *buf[2] = {nullptr, nullptr};
        size_t chains = 0;                                  // each buffer holds this many pieces
        bool in_use = false;
    };
    static Buffers &buffers() {
        static Buffers *b = new Buffers();
        return *b;
    }

    void main() {
        const double t0 = wall_ms();
        Buffers &B = buffers();
        uint8_t *own[2] = {nullptr, nullptr};               // when the shared buffers are taken by another streamer
        size_t own_chains = 0;
        bool shared = false;
        {
            std::lock_guard<std::mutex> lk(B.mu);
            if (!B.in_use) {
                B.in_use = true;
                shared = true;
            }
        }
        auto ensure = [&](size_t chains) -> bool {
            uint8_t **buf = shared ? B.buf : own;
            size_t &have = shared ? B.chains : own_chains;
            if (have >= chains) return true;
            size_t want = std::max<size_t>(have ? have : 1, 1);
            while (want < chains) want *= 2;
            for (int k = 0; k < 2; k++) {
                if (buf[k]) snapgpu_free_pinned(buf[k]);
                buf[k] = static_cast<uint8_t *>(snapgpu_alloc_pinned(want * kPiece));
                if (!buf[k]) {
                    have = 0;
                    return false;
                }
            }
            have = want;
            return true;
        };
        std::vector<Chain *> active;
        std::vector<HostSeg> segs[2];
        std::vector<uint8_t> states[2];
        std::vector<Chain *> round[2];
        std::thread gpu;
        int gpu_rc = 0;
        std::string gpu_err;
        int cur = 0;
        auto retire_round = [&](int r) {                     // after its GPU call: states back, finished chains out
            if (gpu.joinable()) gpu.join();
            if (gpu_rc && !rc_) {
                rc_ = gpu_rc;
                err_ = gpu_err;
            }
            for (size_t i = 0; i < round[r].size(); i++) {
                Chain *c = round[r][i];
                memcpy(c->state, &states[r][64 * i], 64);
                // the piece of THIS round was the chain's last (c->eof may already belong to the next read)
                if (!(segs[r][i].flags & kHostSegNoFinal)) {
                    memcpy(c->digest_out, c->state, 64);
                    if (c->fd >= 0) ::close(c->fd);
                    delete c;
                }
            }
            round[r].clear();
        };
        for (;;) {
            // take on new chains
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (active.empty() && round[cur ^ 1].empty())
                    cv_.wait(lk, [&] { return closing_ || !incoming_.empty(); });
                while (!incoming_.empty() && active.size() < kMaxChains) {
                    active.push_back(incoming_.front());
                    incoming_.pop_front();
                }
                if (active.empty() && incoming_.empty() && closing_ && round[cur ^ 1].empty()) break;
            }
            // read the next piece of every active chain (the previous round is on the GPU meanwhile)
            const int r = cur;
            segs[r].clear();
            round[r].clear();
            if (active.size() > (shared ? B.chains : own_chains)) retire_round(r ^ 1);      // the old buffers are about to go
            if (!active.empty() && !ensure(active.size())) {
                if (!rc_) {
                    rc_ = SNAPGPU_ECUDA;
                    err_ = snapgpu_last_error();
                }
                for (Chain *c : active) {
                    if (c->fd >= 0) ::close(c->fd);
                    delete c;
                }
                active.clear();
            }
            uint8_t *buf = (shared ? B.buf : own)[r];
            std::vector<Chain *> still;
            for (Chain *c : active) {
                if (c->fd < 0 && c->prefix == 0 && !c->eof) {
                    c->fd = ::open(c->path.c_str(), O_RDONLY | O_CLOEXEC);
                    if (c->fd < 0) {
                        *c->err_out = errno;
                        *c->op_out = 2;
                        delete c;
                        continue;
                    }
                }
                const size_t slot = round[r].size();
                const ssize_t got = read_full(c->fd, buf + slot * kPiece, kPiece);
                if (got < 0) {
                    *c->err_out = errno;
                    *c->op_out = 3;
                    ::close(c->fd);
                    delete c;
                    continue;
                }
                c->eof = (size_t)got < kPiece;
                segs[r].push_back(HostSeg{slot * kPiece, (uint64_t)got, c->prefix,
                                          (c->prefix ? kHostSegContinue : 0u) | (c->eof ? 0u : kHostSegNoFinal)});
                round[r].push_back(c);
                if (!c->eof) still.push_back(c);
            }
            active.swap(still);
            // the previous round has to be back before this one's chaining values are known
            retire_round(r ^ 1);
            if (round[r].empty()) continue;
            states[r].resize(64 * round[r].size());
            for (size_t i = 0; i < round[r].size(); i++) {
                memcpy(&states[r][64 * i], round[r][i]->state, 64);
                round[r][i]->prefix += segs[r][i].len;
            }
            gpu_rc = 0;
            if (rc_) {                                       // after a failure: let every chain go, without the GPU
                for (HostSeg &s : segs[r]) s.flags &= ~kHostSegNoFinal;      // every chain of the round ends with it
                active.clear();                              // (the chains still active are all in this round)
            } else {
                gpu = std::thread([&, r, buf] {
                    gpu_rc = sha512_host_segments(buf, segs[r].data(), segs[r].size(), states[r].data());
                    if (gpu_rc) gpu_err = snapgpu_last_error();
                });
            }
            cur ^= 1;
        }
        retire_round(0);
        retire_round(1);
        if (shared) {
            std::lock_guard<std::mutex> lk(B.mu);
            B.in_use = false;
        } else {
            for (auto &p : own)
                if (p) snapgpu_free_pinned(p);
        }
        busy_ms_ = wall_ms() - t0;
    }

    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Chain *> incoming_;
    bool closing_ = false;
    std::thread th_;
    int rc_ = 0;
    std::string err_;
    double busy_ms_ = 0;
};

// ------------------------------------------------------------------------------------------
// the tree
// ------------------------------------------------------------------------------------------

struct TDir;

struct TEntry {
    uint32_t name_off = 0, name_len = 0;     // in TDir::names
    mode_t mode = 0;
    int64_t size = 0;
    uint8_t kind = 0;          // 0 neither, 1 regular file, 2 directory
    uint8_t err_op = 0;        // 1 lstat, 2 open, 3 read, 4 mkdir, 5 write; | 0x80: on the destination path (copy mode)
    bool cached = false;       // digest came from the digest cache
    int err = 0;               // errno of that step
    TDir *child = nullptr;
    uint8_t digest[64];
};

// copy mode: what the library wrote for an entry, the key of its digest in the cache
struct CopyRec {
    bool written = false;      // its bytes went to the destination out of the pinned chunk the GPU hashed
    dev_t dev = 0;
    ino_t ino = 0;
    off_t size = 0;
    struct timespec mtime = {0, 0}, ctime = {0, 0};
};

struct TDir {
    std::string path;          // as the reference would spell it: buildDir + "/" + rel
    std::string rel;           // relative to the build dir with a trailing slash ("" for the root)
    std::string names;         // the entries' names back to back
    std::vector<TEntry> entries;
    std::vector<CopyRec> copies;   // copy mode only, parallel to entries
    int open_err = 0;          // could not be read: Walk hands it to the callback a second time
};

// copyToBuildDir (snappy/build.go:362-418) on the same engine: the scan applies shouldExclude
// instead of the /DEBIAN rule and creates the directories, the packers hard-link what can be
// linked and copy the rest -- a copied file is read ONCE, its bytes go to the destination out of
// the pinned chunk that the GPU hashes.
struct CopySpec {
    std::string dest;
    bool no_link = false;
};

struct linux_dirent64_ {
    uint64_t d_ino;
    int64_t d_off;
    unsigned short d_reclen;
    unsigned char d_type;
    char d_name[1];
};

struct FlatEntry {
    TDir *dir;
    TEntry *e;
};

class TreeHasher {
public:
    // hash: read and hash the regular files on the GPU.  false: scan and lstat only (the test
    // hook that takes its digests from the caller).
    TreeHasher(const std::string &root, bool hash, const CopySpec *copy = nullptr)
        : root_(root), hash_(hash), copy_(copy ? new CopySpec(*copy) : nullptr) {}
    ~TreeHasher() {
        chains.finish();                          // it writes into the entries
        for (TDir *d : dirs_) delete d;
        if (session_) session_close(session_);
        delete copy_;
    }

    ChainStreamer chains;

    int run() {
        IoPool &pool = IoPool::instance();
        struct stat st;
        if (lstat(root_.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) return 0;      // Walk visits only the root: no entries
        if (hash_ && !copy_) {                 // copy mode opens it with the first batch: a tree that links needs no GPU
            int rc = session_open(&session_, kTreeBatchBytes);
            if (rc) return rc;
        }
        TDir *root = new_dir(root_, "");
        root_dir_ = root;
        push_task(Task{root, 0, 0, 0});
        const double t0 = wall_ms();
        pool.run(pool.size(), [this](unsigned w) { worker(w); }, [this] { drive(); });
        t_pack_ms_ = wall_ms() - t0;
        if (fatal_rc_) return fail(fatal_rc_, "%s", fatal_err_.c_str());
        return 0;
    }

    // entries in filepath.Walk order (laid out once; during the GPU tail when there is one)
    const std::vector<FlatEntry> &flat() {
        if (!flattened_) {
            flatten(flat_);
            flattened_ = true;
        }
        return flat_;
    }
    void flatten(std::vector<FlatEntry> &out) const {
        out.clear();
        out.reserve(nentries_.load());
        if (root_dir_) flatten_dir(root_dir_, out);
    }

    // the first failing entry in walk order, the way the reference's Walk would stop at it
    int first_error(const std::vector<FlatEntry> &flat) const {
        // copy mode: a directory that cannot be read is the error Walk hands to the callback, which
        // returns it (build.go:374-376); writeHashes ignores it (build.go:228,241)
        if (copy_ && root_dir_ && root_dir_->open_err)
            return fail(SNAPGPU_EIO, "%s", go_path_error("open", root_dir_->path, root_dir_->open_err).c_str());
        for (const FlatEntry &f : flat) {
            if (copy_ && f.e->child && f.e->child->open_err && !f.e->err)
                return fail(SNAPGPU_EIO, "%s", go_path_error("open", f.e->child->path, f.e->child->open_err).c_str());
            if (f.e->err) {
                static const char *const ops[] = {"", "lstat", "open", "read", "mkdir", "write"};
                const std::string name = f.dir->names.substr(f.e->name_off, f.e->name_len);
                const std::string path = (f.e->err_op & 0x80) && copy_ ? copy_->dest + "/" + f.dir->rel + name : entry_path(f);
                return fail(SNAPGPU_EIO, "%s", go_path_error(ops[f.e->err_op & 7], path, f.e->err).c_str());
            }
        }
        return 0;
    }

    static std::string entry_path(const FlatEntry &f) { return f.dir->path + "/" + f.dir->names.substr(f.e->name_off, f.e->name_len); }

    const std::vector<TDir *> &dirs() const { return dirs_; }
    size_t files_linked() const { return nlinked_.load(); }
    size_t files_hashed() const { return nhashed_.load(); }
    size_t files_cached() const { return ncached_.load(); }
    size_t batches() const { return nbatches_; }
    double pack_ms() const { return t_pack_ms_; }
    double drain_ms() const { return t_drain_ms_; }

private:
    struct Task {
        TDir *dir;
        uint32_t lo, hi;
        int type;              // 0 scan the directory, 1 pack entries [lo, hi)
    };
    static constexpr uint32_t kPackRun = 48;       // files per pack task

    struct WorkerState {
        Chunk *cur[2] = {nullptr, nullptr};
        bool noatime = true;
        std::vector<char> dents;
    };

    // ---- task queue -------------------------------------------------------------------------
    void push_task(const Task &t) {
        {
            std::lock_guard<std::mutex> lk(q_mu_);
            queue_.push_back(t);
            pending_++;
        }
        q_cv_.notify_one();
    }
    // Before a worker blocks for want of tasks it hands over its half-filled chunks: a blocked
    // worker never holds one (see chunk_for), and the tail of the tree reaches the GPU early.
    bool pop_task(WorkerState &W, Task *t) {
        std::unique_lock<std::mutex> lk(q_mu_);
        for (;;) {
            if (!queue_.empty()) {
                *t = queue_.front();
                queue_.pop_front();
                return true;
            }
            if (pending_ == 0 || abort_.load()) return false;
            if (W.cur[0] || W.cur[1]) {
                lk.unlock();
                for (int cls = 0; cls < 2; cls++) flush_chunk(W, cls, true);
                lk.lock();
                continue;
            }
            q_cv_.wait(lk);
        }
    }
    void task_done() {
        bool last;
        {
            std::lock_guard<std::mutex> lk(q_mu_);
            last = --pending_ == 0;
        }
        if (last) q_cv_.notify_all();
    }

    TDir *new_dir(const std::string &path, const std::string &rel) {
        TDir *d = new TDir();
        d->path = path;
        d->rel = rel;
        std::lock_guard<std::mutex> lk(dirs_mu_);
        dirs_.push_back(d);
        return d;
    }

    // ---- workers ----------------------------------------------------------------------------
    void worker(unsigned) {
        WorkerState W;
        W.dents.resize(256 << 10);
        Task t;
        while (pop_task(W, &t)) {
            if (!abort_.load()) {
                if (t.type == 0) scan(W, t.dir);
                else pack(W, t.dir, t.lo, t.hi, -1);
            }
            task_done();
        }
        for (int cls = 0; cls < 2; cls++) flush_chunk(W, cls, true);
        {
            std::lock_guard<std::mutex> lk(r_mu_);
            workers_done_++;
        }
        r_cv_.notify_all();
    }
    // clflushopt over [p, p + n): the lines leave the cache hierarchy, the data is in memory (x86 with CLFLUSHOPT;
    // elsewhere a no-op -- it is an optimisation of the copy that follows, not a requirement)
    static bool have_clflushopt() {
#if defined(__x86_64__)
        static const bool have = [] {
            unsigned a = 7, b = 0, c = 0, d = 0;
            __asm__ volatile("cpuid" : "+a"(a), "=b"(b), "+c"(c), "=d"(d));
            return (b >> 23 & 1) != 0;
        }();
        return have;
#else
        return false;
#endif
    }
    static void cache_write_back(const uint8_t *p, size_t n) {
#if defined(__x86_64__)
        if (!have_clflushopt()) return;
        const uintptr_t lo = (uintptr_t)p & ~(uintptr_t)63, hi = (uintptr_t)p + n;
        for (uintptr_t a = lo; a < hi; a += 64) __asm__ volatile("clflushopt (%0)" ::"r"(a) : "memory");
        __asm__ volatile("sfence" ::: "memory");
#else
        (void)p; (void)n;
#endif
    }
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }

    void scan(WorkerState &W, TDir *dir) {
        const int dfd = ::open(dir->path.c_str(), O_RDONLY | O_DIRECTORY | O_CLOEXEC);
        if (dfd < 0) {
            dir->open_err = errno;
            return;
        }
        // names, sorted bytewise like readDirNames (sort.Strings)
        struct Name {
            uint32_t off, len;
            uint8_t type;
        };
        std::vector<Name> names;
        std::string &arena = dir->names;
        const bool top = dir->rel.empty();
        for (;;) {
            const long n = syscall(SYS_getdents64, dfd, W.dents.data(), W.dents.size());
            if (n < 0 && errno == EINTR) continue;
            if (n <= 0) {
                if (n < 0 && names.empty()) {                // readdir failed: like a directory that cannot be opened
                    dir->open_err = errno;
                    ::close(dfd);
                    return;
                }
                break;
            }
            for (long p = 0; p < n;) {
                const linux_dirent64_ *e = reinterpret_cast<const linux_dirent64_ *>(W.dents.data() + p);
                p += e->d_reclen;
                const char *nm = e->d_name;
                if (nm[0] == '.' && (!nm[1] || (nm[1] == '.' && !nm[2]))) continue;
                // build.go:229: anything whose path below the build dir starts with "/DEBIAN" is
                // skipped, which the top-level name decides for the whole subtree
                if (!copy_ && top && strncmp(nm, "DEBIAN", 6) == 0) continue;
                if (copy_ && should_exclude(nm)) continue;          // build.go:378-383: the entry, a directory with its subtree
                const size_t len = strlen(nm);
                names.push_back(Name{(uint32_t)arena.size(), (uint32_t)len, e->d_type});
                arena.append(nm, len);
            }
        }
        const char *base = arena.data();
        std::sort(names.begin(), names.end(), [base](const Name &a, const Name &b) {
            const int c = memcmp(base + a.off, base + b.off, std::min(a.len, b.len));
            return c ? c < 0 : a.len < b.len;
        });
        dir->entries.resize(names.size());
        if (copy_) dir->copies.resize(names.size());
        nentries_ += names.size();
        uint32_t nreg = 0;
        std::string child_name;
        for (size_t i = 0; i < names.size(); i++) {
            TEntry &e = dir->entries[i];
            e.name_off = names[i].off;
            e.name_len = names[i].len;
            if (names[i].type == DT_REG && hash_) {          // lstat'ed through its descriptor when it is packed
                e.kind = 1;
                nreg++;
                continue;
            }
            child_name.assign(base + e.name_off, e.name_len);
            struct stat st;
            if (fstatat(dfd, child_name.c_str(), &st, AT_SYMLINK_NOFOLLOW) != 0) {
                e.err = errno;
                e.err_op = 1;
                continue;
            }
            e.mode = st.st_mode;
            e.size = st.st_size;
            if (S_ISREG(st.st_mode)) {
                e.kind = 1;
                if (hash_) nreg++;
            } else if (S_ISDIR(st.st_mode)) {
                e.kind = 2;
                if (copy_ && ::mkdir((copy_->dest + "/" + dir->rel + child_name).c_str(), st.st_mode & 07777) != 0) {
                    e.err = errno;                             // os.Mkdir(dest, info.Mode()), build.go:386-388
                    e.err_op = 4 | 0x80;
                    continue;
                }
                e.child = new_dir(dir->path + "/" + child_name, dir->rel + child_name + "/");
                push_task(Task{e.child, 0, 0, 0});
            } else if (copy_) {
                // a symlink, a fifo ...: linked if possible, else copied the way the reference does it
                // (os.Open follows a symlink; build.go:391-416)
                const std::string dest = copy_->dest + "/" + dir->rel + child_name;
                if (copy_->no_link || ::linkat(dfd, child_name.c_str(), AT_FDCWD, dest.c_str(), 0) != 0)
                    plain_copy(dir->path + "/" + child_name, dest, st.st_mode, e);
            }
        }
        if (nreg) {
            // a small directory is packed right here through the descriptor that is already
            // open, a large one is cut into runs that any worker takes
            const uint32_t n = (uint32_t)names.size();
            // (copy mode: creating files serialises on the directory's lock, so one directory is one
            // worker's unless it is huge -- the workers are then in different directories)
            if (n <= 2 * kPackRun || (copy_ && n <= 8192)) {
                pack(W, dir, 0, n, dfd);
            } else {
                for (uint32_t lo = kPackRun; lo < n; lo += kPackRun) push_task(Task{dir, lo, std::min(n, lo + kPackRun), 1});
                pack(W, dir, 0, kPackRun, dfd);
            }
        }
        ::close(dfd);
    }

    // a chunk with room for `need` bytes of class cls; nullptr when the run was aborted
    Chunk *chunk_for(WorkerState &W, int cls, size_t need) {
        Chunk *c = W.cur[cls];
        if (c && c->used + need <= c->cap - kChunkSlack) return c;
        flush_chunk(W, cls);
        ChunkPool &pool = cls ? large_chunks() : small_chunks();
        if (!(c = pool.try_get())) {
            // nothing free: hand over what this worker holds (nobody waits while sitting on a
            // half-filled chunk: every chunk is then free, being filled by a running worker,
            // ready or in flight), wake the driver, wait for a new or a recycled chunk
            flush_chunk(W, cls ^ 1);
            bool asked = false;
            // a chunk whose copy has just finished is usually back within a few hundred microseconds; only a worker
            // that has waited longer asks for the pool to grow -- and much longer once the pool holds a quarter of a
            // gigabyte: from there a wait is a hiccup of the pipeline, not a pool that is too small, and pinning
            // another slab in the middle of a warm call costs that call more than the wait (measured: +37 ms)
            const int patience = pool.allocated() * (cls ? kLargeChunk : kSmallChunk) >= ((size_t)256 << 20) ? 40 : 2;
            for (int spins = 0; !(c = pool.try_get()); spins++) {
                if (spins == patience && !asked) {
                    pool.begin_wait();
                    asked = true;
                }
                std::unique_lock<std::mutex> lk(r_mu_);
                r_cv_.notify_all();
                chunk_cv_.wait_for(lk, std::chrono::microseconds(150));
                if (abort_.load()) break;
            }
            if (asked) pool.end_wait();
            if (!c) return nullptr;
        }
        W.cur[cls] = c;
        return c;
    }
    // going_idle: the worker is about to block.  What it wrote last is still in its core's cache, and the copy
    // engine's reads of lines held by an IDLE core are slow: measured 200 us for a 1.3 MiB chunk (6.5 GB/s)
    // against 82 us for a full 4 MiB chunk handed over while its writer kept running -- 3 ms on the tail of a
    // 40 ms tree, where every worker hands over its last chunk and stops.  (Keeping the workers spinning through
    // the tail fixes the copies as well, and costs more than it gains as soon as the process has a CPU quota.)
    // So a worker that is going idle writes the last 2 MiB of the chunk back first.
    void flush_chunk(WorkerState &W, int cls, bool going_idle = false) {
        Chunk *c = W.cur[cls];
        if (!c) return;
        W.cur[cls] = nullptr;
        if (c->files.empty()) {
            (cls ? large_chunks() : small_chunks()).put(c);
            return;
        }
        if (going_idle && !copy_ && write_back_) {
            const size_t span = std::min<size_t>(c->used, (size_t)2 << 20);
            cache_write_back(c->base + (c->used - span), span);
        }
        {
            std::lock_guard<std::mutex> lk(r_mu_);
            ready_.push_back(c);
        }
        r_cv_.notify_all();
    }

    void pack(WorkerState &W, TDir *dir, uint32_t lo, uint32_t hi, int dfd_in) {
        int dfd = dfd_in;
        if (dfd < 0) {
            dfd = ::open(dir->path.c_str(), O_RDONLY | O_DIRECTORY | O_CLOEXEC);
            if (dfd < 0) {                                   // it was readable a moment ago
                const int err = errno;
                for (uint32_t i = lo; i < hi; i++)
                    if (dir->entries[i].kind == 1) {
                        dir->entries[i].err = err;
                        dir->entries[i].err_op = 2;
                    }
                return;
            }
        }
        std::string name;
        const char *base = dir->names.data();
        size_t hashed = 0, cached = 0, linked = 0;
        const bool try_cache = !copy_ && cache_nonempty();
        int ddfd = -1;                                         // copy mode: the destination directory
        if (copy_) {
            ddfd = ::open((copy_->dest + "/" + dir->rel).c_str(), O_RDONLY | O_DIRECTORY | O_CLOEXEC);
            if (ddfd < 0) {
                const int err = errno;
                for (uint32_t i = lo; i < hi; i++)
                    if (dir->entries[i].kind == 1) {
                        dir->entries[i].err = err;
                        dir->entries[i].err_op = 2 | 0x80;
                    }
                if (dfd_in < 0) ::close(dfd);
                return;
            }
        }
        for (uint32_t i = lo; i < hi && !abort_.load(); i++) {
            TEntry &e = dir->entries[i];
            if (e.kind != 1) continue;
            name.assign(base + e.name_off, e.name_len);
            if (copy_ && !copy_->no_link && ::linkat(dfd, name.c_str(), ddfd, name.c_str(), 0) == 0) {
                linked++;                                      // "whee" (build.go:392-395): nothing to read
                continue;
            }
            if (try_cache) {                                   // after copyToBuildDir: an lstat decides, the file is not opened
                struct stat cst;
                if (fstatat(dfd, name.c_str(), &cst, AT_SYMLINK_NOFOLLOW) == 0 && S_ISREG(cst.st_mode) &&
                    cache_lookup(cst, e.digest)) {
                    e.mode = cst.st_mode;
                    e.size = cst.st_size;
                    e.cached = true;
                    cached++;
                    continue;
                }
            }
            // O_NONBLOCK: should the name have become a fifo since the scan, the open returns
            int fd = ::openat(dfd, name.c_str(), O_RDONLY | O_CLOEXEC | O_NOFOLLOW | O_NONBLOCK | (W.noatime ? O_NOATIME : 0));
            if (fd < 0 && errno == EPERM && W.noatime) {     // O_NOATIME is for the owner only
                W.noatime = false;
                fd = ::openat(dfd, name.c_str(), O_RDONLY | O_CLOEXEC | O_NOFOLLOW | O_NONBLOCK);
            }
            if (fd < 0) {
                e.err = errno;
                e.err_op = 2;
                continue;
            }
            struct stat st;
            if (fstat(fd, &st) != 0) {
                e.err = errno;
                e.err_op = 1;
                ::close(fd);
                continue;
            }
            e.mode = st.st_mode;
            e.size = st.st_size;
            if (!S_ISREG(st.st_mode)) {                      // replaced since the scan: listed with its mode, not hashed
                e.kind = 0;
                ::close(fd);
                continue;
            }
            if (!copy_ && cache_lookup(st, e.digest)) {
                e.cached = true;
                cached++;
                ::close(fd);
                continue;
            }
            const uint64_t size = (uint64_t)st.st_size;
            if (copy_ && size > kMidMax) {                   // too long for a chunk: plain io.Copy, hashed later by writeHashes
                ::close(fd);
                plain_copy(dir->path + "/" + name, copy_->dest + "/" + dir->rel + name, st.st_mode, e);
                continue;
            }
            hashed++;
            if (size > kMidMax) {                            // a chain of its own
                ::close(fd);
                chains.add(dir->path + "/" + name, e.digest, &e.err, &e.err_op);
                continue;
            }
            const int cls = size > kSmallMax ? 1 : 0;
            const size_t need = align_up(size + 1);          // one spare byte: a file that grew shows
            Chunk *c = chunk_for(W, cls, need);
            if (!c) {
                ::close(fd);
                break;
            }
            uint8_t *dst = c->base + c->used;
            size_t got = 0;
            int err = 0;
            for (;;) {
                const ssize_t r = ::read(fd, dst + got, size + 1 - got);
                if (r < 0) {
                    if (errno == EINTR) continue;
                    err = errno;
                    break;
                }
                got += (size_t)r;
                if (r == 0 || got == size || got == size + 1) break;
            }
            ::close(fd);
            if (err) {
                e.err = err;
                e.err_op = 3;
                continue;
            }
            if (got > size) {                                // grew after its fstat: hashed to EOF as a chain, like io.Copy would
                if (copy_) plain_copy(dir->path + "/" + name, copy_->dest + "/" + dir->rel + name, st.st_mode, e);
                else chains.add(dir->path + "/" + name, e.digest, &e.err, &e.err_op);
                continue;
            }
            if (copy_) {
                // the same bytes the GPU is about to hash go to the destination (build.go:403-414)
                const int out = ::openat(ddfd, name.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_CLOEXEC, st.st_mode & 07777);
                if (out < 0) {
                    e.err = errno;
                    e.err_op = 2 | 0x80;
                    continue;
                }
                struct stat wst;
                int werr = 0;
                for (size_t done = 0; done < got && !werr;) {
                    const ssize_t w = ::write(out, dst + done, got - done);
                    if (w < 0 && errno == EINTR) continue;
                    if (w < 0) werr = errno;
                    else done += (size_t)w;
                }
                if (!werr && fstat(out, &wst) != 0) werr = errno;
                if (::close(out) != 0 && !werr) werr = errno;
                if (werr) {
                    e.err = werr;
                    e.err_op = 5 | 0x80;
                    continue;
                }
                CopyRec &r = dir->copies[i];
                r.written = true;
                r.dev = wst.st_dev;
                r.ino = wst.st_ino;
                r.size = wst.st_size;
                r.mtime = wst.st_mtim;
                r.ctime = wst.st_ctim;
            }
            c->files.push_back(PackedRef{e.digest, (uint32_t)c->used, (uint32_t)got});
            c->used += align_up(got + 1);
            // the GPU has nothing to do (the start of a tree, or a slow disk): do not sit on a
            // chunk until it is full
            if (hungry_.load(std::memory_order_relaxed) && c->used >= kHungryFlush) flush_chunk(W, cls);
        }
        if (dfd_in < 0) ::close(dfd);
        if (ddfd >= 0) ::close(ddfd);
        nhashed_ += hashed;
        ncached_ += cached;
        nlinked_ += linked;
    }

    // io.Copy of one entry without the GPU (build.go:396-416): what cannot go through a chunk.  Its
    // digest is not remembered; the writeHashes that follows reads the copy.
    static void plain_copy(const std::string &src, const std::string &dest, mode_t mode, TEntry &e) {
        const int in = ::open(src.c_str(), O_RDONLY | O_CLOEXEC);
        if (in < 0) {
            e.err = errno;
            e.err_op = 2;
            return;
        }
        const int out = ::open(dest.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_CLOEXEC, mode & 07777);
        if (out < 0) {
            e.err = errno;
            e.err_op = 2 | 0x80;
            ::close(in);
            return;
        }
        std::vector<uint8_t> buf(1 << 20);
        for (;;) {
            const ssize_t r = ::read(in, buf.data(), buf.size());
            if (r < 0 && errno == EINTR) continue;
            if (r < 0) {
                e.err = errno;
                e.err_op = 3;
                break;
            }
            if (r == 0) break;
            ssize_t done = 0;
            while (done < r) {
                const ssize_t w = ::write(out, buf.data() + done, (size_t)(r - done));
                if (w < 0 && errno == EINTR) continue;
                if (w < 0) {
                    e.err = errno;
                    e.err_op = 5 | 0x80;
                    break;
                }
                done += w;
            }
            if (e.err) break;
        }
        ::close(in);
        if (::close(out) != 0 && !e.err) {
            e.err = errno;
            e.err_op = 5 | 0x80;
        }
    }

    static bool cache_lookup(const struct stat &st, uint8_t digest[64]);
    static bool cache_nonempty();

    // ---- the driving thread: chunk allocation, batch submission, chunk recycling --------------
    void drive() {
        if (!hash_) return;
        const unsigned nworkers = IoPool::instance().size();
        std::vector<Chunk *> take;
        std::vector<HostSpan> spans;
        std::vector<SpanSeg> segs;
        std::vector<uint8_t *> dst;
        std::vector<uint64_t> copied;
        // 64 MiB batches: four of them in flight (H2D ~1.2 ms, kernel >= 2 ms each) carry more than
        // twice what sixteen packers produce, and the staging buffers (four of this size per device)
        // are allocated once and never grow -- cudaMalloc is slow enough to show in a 50 ms call
        const size_t cap_bytes = kTreeBatchBytes - 4096, cap_items = (size_t)1 << 20, kMinBatch = (size_t)32 << 20;
        bool all_done = false;
        const bool trace = getenv("SNAPGPU_TRACE") != nullptr;
        const double t_start = wall_ms();
        while (!fatal_rc_) {
            bool progressed = false;
            const int made_small = small_chunks().serve_allocations(), made_large = large_chunks().serve_allocations();
            if (made_small < 0 || made_large < 0) {
                fatal_rc_ = SNAPGPU_ECUDA;
                fatal_err_ = std::string("no pinned memory for the file packer: ") + snapgpu_last_error();
                break;
            }
            if (made_small + made_large > 0) {
                chunk_cv_.notify_all();
                progressed = true;
            }
            // chunks whose copy has finished go back to the pool
            copied.clear();
            int rc = poll_session(&copied, false);
            if (rc) { fatal(rc); break; }
            if (recycle(copied)) progressed = true;
            // Submit what is ready when a slot is free -- but not crumbs: every batch costs the GPU at
            // least the chain of its longest file (2 ms for a 64 KiB one) and holds a slot for that
            // long, so unless the GPU has nothing to do, or the workers are done, a batch waits until
            // kMinBatch bytes are ready (the packers fill that in about a millisecond).  While every
            // slot is taken the chunks keep collecting and the next batch is as large as the GPU's
            // pace allows.
            take.clear();
            if (in_flight() < capacity()) {
                std::lock_guard<std::mutex> lk(r_mu_);
                size_t bytes = 0, items = 0;
                const bool workers_done = workers_done_ == nworkers;
                if (!workers_done && in_flight() > 0) {
                    size_t ready_bytes = 0;
                    for (Chunk *c : ready_) ready_bytes += c->used;
                    if (ready_bytes < kMinBatch) bytes = cap_bytes + 1;          // not yet
                }
                while (!ready_.empty() && bytes <= cap_bytes) {
                    Chunk *c = ready_.front();
                    if (!take.empty() && (bytes + c->used > cap_bytes || items + c->files.size() > cap_items)) break;
                    bytes += (c->used + 255) & ~(size_t)255;
                    items += c->files.size();
                    take.push_back(c);
                    ready_.pop_front();
                }
                all_done = workers_done_ == nworkers && ready_.empty();
            }
            hungry_.store(take.empty() && in_flight() == 0, std::memory_order_relaxed);
            if (!take.empty()) {
                spans.clear();
                segs.clear();
                dst.clear();
                for (size_t k = 0; k < take.size(); k++) {
                    spans.push_back(HostSpan{take[k]->base, take[k]->used});
                    for (const PackedRef &f : take[k]->files) {
                        segs.push_back(SpanSeg{(uint32_t)k, f.off, f.len});
                        dst.push_back(f.digest);
                    }
                }
                uint64_t ticket = 0;
                copied.clear();
                const double t_submit = wall_ms();
                rc = 0;
                if (!session_ && !(rc = ensure_init())) rc = session_open(&session_, kTreeBatchBytes);     // copy mode: first batch
                if (!rc) rc = session_submit(session_, spans.data(), spans.size(), segs.data(), dst.data(), segs.size(), &ticket, &copied);
                if (rc) { fatal(rc); break; }
                in_copy_.emplace_back(ticket, take);
                recycle(copied);
                nbatches_++;
                if (trace) {
                    size_t bytes = 0;
                    for (Chunk *c : take) bytes += c->used;
                    fprintf(stderr, "[snapgpu] tree batch %zu at %.2f ms: %zu chunks, %.1f MiB, %zu files; %zu in flight, submit took %.2f ms\n",
                            nbatches_, wall_ms() - t_start, take.size(), bytes / 1048576.0, segs.size(), in_flight(),
                            wall_ms() - t_submit);
                }
                continue;
            }
            if (all_done) break;
            if (!progressed) {
                std::unique_lock<std::mutex> lk(r_mu_);
                r_cv_.wait_for(lk, std::chrono::microseconds(in_flight() ? 50 : 300));
            }
        }
        if (fatal_rc_) {
            abort_ = true;
            q_cv_.notify_all();
            chunk_cv_.notify_all();
            // let the workers run out, taking back what they hand over
            for (;;) {
                std::unique_lock<std::mutex> lk(r_mu_);
                for (Chunk *c : ready_) (c->cls ? large_chunks() : small_chunks()).put(c);
                ready_.clear();
                if (workers_done_ == nworkers) break;
                r_cv_.wait_for(lk, std::chrono::microseconds(500));
            }
        }
        const double t0 = wall_ms();
        if (trace) fprintf(stderr, "[snapgpu] tree: workers done at %.2f ms, %zu batches in flight\n", t0 - t_start, in_flight());
        // every directory has been read: the walk order can be laid out while the last batches are
        // still on the GPU (it needs names and modes, not digests)
        if (!fatal_rc_) {
            flatten(flat_);
            flattened_ = true;
        }
        // the last batches: first until everything has been copied (the workers stay awake for that, see worker()),
        // then until the digests are back
        int rc = 0;
        while (!fatal_rc_ && !in_copy_.empty()) {
            copied.clear();
            if ((rc = poll_session(&copied, false))) break;
            if (!recycle(copied)) cpu_relax();
        }
        copied.clear();
        if (!rc) rc = poll_session(&copied, true);
        if (rc && !fatal_rc_) fatal(rc);
        for (auto &p : in_copy_)
            for (Chunk *c : p.second) (c->cls ? large_chunks() : small_chunks()).put(c);
        in_copy_.clear();
        t_drain_ms_ = wall_ms() - t0;
    }
    // the session opens with the first batch in copy mode (a tree that hard-links needs no GPU)
    size_t in_flight() const { return session_ ? session_in_flight(session_) : 0; }
    size_t capacity() const { return session_ ? session_capacity(session_) : 1; }
    int poll_session(std::vector<uint64_t> *copied, bool wait_all) { return session_ ? session_poll(session_, copied, wait_all) : 0; }
    bool recycle(const std::vector<uint64_t> &copied) {
        bool any = false;
        for (uint64_t t : copied)
            for (size_t k = 0; k < in_copy_.size(); k++)
                if (in_copy_[k].first == t) {
                    for (Chunk *c : in_copy_[k].second) (c->cls ? large_chunks() : small_chunks()).put(c);
                    in_copy_.erase(in_copy_.begin() + (long)k);
                    any = true;
                    break;
                }
        if (any) chunk_cv_.notify_all();
        return any;
    }
    void fatal(int rc) {
        fatal_rc_ = rc;
        fatal_err_ = snapgpu_last_error();
    }

    void flatten_dir(TDir *d, std::vector<FlatEntry> &out) const {
        for (TEntry &e : d->entries) {
            out.push_back(FlatEntry{d, &e});
            if (e.child) {
                // Go reports a directory it cannot read to the callback a second time, with the
                // error, and writeHashes ignores that error (build.go:228,241): it is listed twice
                if (e.child->open_err) out.push_back(FlatEntry{d, &e});
                else flatten_dir(e.child, out);
            }
        }
    }

    const std::string root_;
    const bool hash_;
    const CopySpec *const copy_;
    std::vector<FlatEntry> flat_;
    bool flattened_ = false;
    BatchSession *session_ = nullptr;
    TDir *root_dir_ = nullptr;
    std::mutex dirs_mu_;
    std::vector<TDir *> dirs_;
    std::atomic<size_t> nentries_{0}, nhashed_{0}, ncached_{0}, nlinked_{0};
    // task queue
    std::mutex q_mu_;
    std::condition_variable q_cv_;
    std::deque<Task> queue_;
    size_t pending_ = 0;
    std::atomic<bool> abort_{false}, hungry_{true};
    // ready chunks
    std::mutex r_mu_;
    std::condition_variable r_cv_, chunk_cv_;
    std::deque<Chunk *> ready_;
    unsigned workers_done_ = 0;
    const bool write_back_ = getenv("SNAPGPU_NO_WRITE_BACK") == nullptr;    // A/B switch for the measurement at flush_chunk
    std::vector<std::pair<uint64_t, std::vector<Chunk *>>> in_copy_;
    int fatal_rc_ = 0;
    std::string fatal_err_;
    size_t nbatches_ = 0;
    double t_pack_ms_ = 0, t_drain_ms_ = 0;
};
