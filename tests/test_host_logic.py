"""Host logic of libsnapgpu that needs no GPU: the tree walk + yaml.v2-exact writer (fed
with digests from the oracle), the length binning, the multi-GPU sharder and the chunker."""
import ctypes
import os
import stat

import numpy as np
import pytest

from conftest import make_reference_tree


def cxx_yaml(native, oracle, tree: str, tar: str) -> bytes:
    """YAML from the C++ walk/emitter with digests supplied by the oracle."""
    L = native.lib()
    need = ctypes.c_size_t()
    native.check(L.snapgpu_test_yaml_from_digests(os.fsencode(tree), None, 0, None, ctypes.byref(need)))
    archive, entries = oracle.collect_hashes(tree, tar)
    hexes = [archive] + [e["sha512"] for e in entries if e["size"] is not None]
    assert need.value == len(hexes)
    dg = np.frombuffer(bytes.fromhex("".join(hexes)), dtype=np.uint8).copy()
    ptr, ln = ctypes.c_void_p(), ctypes.c_size_t()
    rc = L.snapgpu_test_yaml_from_digests(os.fsencode(tree), dg.ctypes.data, len(hexes), ctypes.byref(ptr),
                                          ctypes.byref(ln))
    native.check(rc)
    return native.take_string(ptr, ln.value)


def test_golden_tree(native, oracle, golden_dir, tmp_path):
    tree = tmp_path / "tree"
    tree.mkdir()
    make_reference_tree(tree)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(b"")
    assert cxx_yaml(native, oracle, str(tree), str(tar)) == (golden_dir / "hashes_simple.yaml").read_bytes()


NASTY = ["true", "123", "1e3", "~", "null", "#x", "a: b", "a #b", " lead", "trail ", "- x", "-x", "0x1f", "1_000",
         "12:30", "it's", "tab\there", "multi\nline", "café", "x" * 70 + " " + "y" * 30 + " z", "\U0001F600",
         "[abc]", "{a}", "*star", "&amp", "!bang", "%pct", "@at", "`tick", "a\\b", 'q"uote', "...", "---", "?",
         ":", "y", "No", ".5", "+.inf", "0b101", "<<", "080", "1.", "+1", "e5", "-", "--", "a,b", "k: ",
         "sp  ace", "w" * 90 + "  double", "ctrl\x01x", "del\x7f", "nbsp x", "bom﻿x", "trail\n",
         "\nlead", " \n", "a \nb", "0o17", "1e400", "9" * 25, "-0b11", "0b" + "1" * 65, "٣", "12:30:45.5", "1:2:x"]


def test_nasty_names_match_oracle(native, oracle, tmp_path):
    """Names outside the plain-safe subset: C++ writer == Python restatement (parity unpinned
    against yaml.v2 itself)."""
    tree = tmp_path / "t"
    tree.mkdir()
    made = 0
    for i, n in enumerate(NASTY):
        try:
            p = tree / n
            p.write_bytes(("content %d" % i).encode())
            made += 1
        except (OSError, ValueError):
            continue
    assert made > 50
    os.mkfifo(tree / "zz-fifo") if False else None
    (tree / b"bin\xff\xfe".decode("utf-8", "surrogateescape")).write_bytes(b"invalid utf8 name")
    (tree / (b"\xff" * 80).decode("utf-8", "surrogateescape")).write_bytes(b"long invalid name")
    tar = tmp_path / "d.tar.gz"
    tar.write_bytes(b"tar")
    want = oracle.write_hashes(str(tree), str(tar))
    got = cxx_yaml(native, oracle, str(tree), str(tar))
    assert got == want


def test_random_trees_match_oracle(native, oracle, tmp_path):
    rng = np.random.default_rng(7)
    alphabet = list("abcXYZ019._-+ #:'\"\t") + ["é", "\n"]
    for t in range(6):
        tree = tmp_path / f"t{t}"
        tree.mkdir()
        for d in range(int(rng.integers(1, 4))):
            sub = tree / ("sub%d" % d)
            sub.mkdir(mode=0o750)
            for f in range(int(rng.integers(0, 12))):
                name = "".join(rng.choice(alphabet, size=int(rng.integers(1, 40)))).strip("/") or "x"
                if name in (".", ".."):
                    continue
                try:
                    p = sub / name
                    p.write_bytes(os.urandom(int(rng.integers(0, 300))))
                    os.chmod(p, int(rng.choice([0o644, 0o600, 0o755, 0o4711, 0o000, 0o664])))
                except OSError:
                    pass
            if rng.random() < 0.5:
                os.symlink("target", sub / "link")
        tar = tmp_path / f"d{t}"
        tar.write_bytes(os.urandom(50))
        if os.geteuid() == 0:
            pass   # root can read mode-000 files, so the oracle can hash them
        want = oracle.write_hashes(str(tree), str(tar))
        got = cxx_yaml(native, oracle, str(tree), str(tar))
        assert got == want


def test_unknown_file_mode(native, oracle, tmp_path):
    tree = tmp_path / "t"
    tree.mkdir()
    os.mkfifo(tree / "pipe", 0o644)
    tar = tmp_path / "d"
    tar.write_bytes(b"")
    with pytest.raises(oracle.UnknownFileMode):
        oracle.write_hashes(str(tree), str(tar))
    L = native.lib()
    dg = np.zeros(64, dtype=np.uint8)
    ptr, ln = ctypes.c_void_p(), ctypes.c_size_t()
    rc = L.snapgpu_test_yaml_from_digests(os.fsencode(str(tree)), dg.ctypes.data, 1, ctypes.byref(ptr), ctypes.byref(ln))
    assert rc == native.EMODE
    assert native.last_error() == "Unknown file mode prw-r--r--"


def test_walk_order_and_trailing_slash(native, oracle, tmp_path):
    tree = tmp_path / "t"
    for name in ("b", "a", "a.b", "a-b", "a/z", "B", "a0", "_", "~t"):
        p = tree / name
        p.parent.mkdir(parents=True, exist_ok=True)
        if not p.exists():
            p.write_bytes(name.encode()) if name != "a" else None
    tar = tmp_path / "d"
    tar.write_bytes(b"")
    want = oracle.write_hashes(str(tree), str(tar))
    assert cxx_yaml(native, oracle, str(tree) + "/", str(tar)) == want
    names = [l.split(b": ", 1)[1] for l in want.splitlines() if l.startswith(b"- name")]
    assert names == sorted(names, key=lambda n: n.replace(b"/", b"\x00"))   # Walk = sorted by path components


# ---- launch plan -----------------------------------------------------------------------------

def test_sharder_balances_and_covers(native):
    rng = np.random.default_rng(4)
    for ndev in (1, 2, 4, 8):
        for weights in (np.full(2000, 513, dtype=np.uint64),
                        rng.integers(9, 514, 20000).astype(np.uint64),
                        np.concatenate([rng.integers(9, 514, 50000), [8_388_609] * 4]).astype(np.uint64)):
            dev = np.full(len(weights), -1, dtype=np.int32)
            native.check(native.lib().snapgpu_test_shard(weights.ctypes.data, len(weights), ndev, dev.ctypes.data))
            assert dev.min() >= 0 and dev.max() < ndev
            load = np.bincount(dev, weights=weights.astype(np.float64), minlength=ndev)
            ideal = max(weights.sum() / ndev, weights.max())
            assert load.max() <= ideal * 1.02 + weights.max() * 0 + 600, (ndev, load)
            # small items stay in contiguous runs per device (dense H2D spans)
            small = weights <= max(weights.sum() // (4 * ndev), 1)
            d_small = dev[small]
            assert np.all(np.diff(d_small) >= 0)


def test_contiguous_split_for_several_devices(native):
    """The in-place multi-device split of a plain file list: contiguous index ranges that cover the list,
    weights (blocks + 1) within one file of equal, and a refusal when one file is too heavy to be cut around."""
    rng = np.random.default_rng(9)
    blocks = lambda ln: (ln + 144) // 128 + 1
    for ndev in (2, 4, 8):
        for lengths in (np.full(2000, 65536, dtype=np.uint64), rng.integers(0, 70_000, 20_000).astype(np.uint64),
                        np.clip(np.round(np.exp(rng.normal(np.log(8192), 1.0, 100_000))), 1024, 65536).astype(np.uint64),
                        np.zeros(50, dtype=np.uint64), np.array([5, 5, 5], dtype=np.uint64)):
            cut = np.zeros(ndev + 1, dtype=np.uint64)
            rc = native.lib().snapgpu_test_split(lengths.ctypes.data, len(lengths), ndev, cut.ctypes.data)
            w = blocks(lengths.astype(np.int64))
            if rc == 0:                                   # only when a single file outweighs 1/(4 ndev) of the job
                assert w.max() > max(w.sum() // (4 * ndev), 1)
                continue
            assert rc == 1
            c = cut.astype(np.int64)
            assert c[0] == 0 and c[-1] == len(lengths) and np.all(np.diff(c) >= 0)
            loads = np.array([w[c[d]:c[d + 1]].sum() for d in range(ndev)])
            assert loads.sum() == w.sum()
            if len(lengths) >= 100 * ndev:
                assert loads.max() - loads.min() <= 2 * w.max() + 2, (ndev, loads)
        heavy = np.concatenate([rng.integers(0, 5000, 1000), [1 << 30]]).astype(np.uint64)
        cut = np.zeros(ndev + 1, dtype=np.uint64)
        assert native.lib().snapgpu_test_split(heavy.ctypes.data, len(heavy), ndev, cut.ctypes.data) == 0


def test_chunker_covers_every_byte_once(native):
    rng = np.random.default_rng(5)
    lengths = np.concatenate([rng.integers(0, 40000, 300), [1_500_000, 0, 999_999, 128, 1_048_576]]).astype(np.uint64)
    offsets = np.zeros(len(lengths), dtype=np.uint64)
    pos = 0
    for i, l in enumerate(lengths):
        pos = (pos + 15) // 16 * 16 + int(rng.integers(0, 3))        # not always aligned
        offsets[i] = pos
        pos += int(l)
    cap = 1 << 20
    for is_sha in (1, 0):
        rows = np.zeros((4096, 6), dtype=np.uint64)
        n = native.lib().snapgpu_test_chunks(offsets.ctypes.data, lengths.ctypes.data, len(lengths), cap, is_sha,
                                             rows.ctypes.data, len(rows))
        assert 0 < n <= len(rows)
        rows = rows[:n]
        covered = {}
        for user, off, ln, prefix, flags, chunk in rows.tolist():
            assert off == offsets[user] + prefix
            assert prefix == covered.get(user, 0)              # pieces arrive in order
            covered[user] = prefix + ln
            if is_sha:
                first, last = prefix == 0, prefix + ln == lengths[user]
                assert bool(flags & 1) == (not first)          # continue
                assert bool(flags & 2) == (not last)           # no-final
                if not last:
                    assert ln % 128 == 0
        assert all(covered[i] == int(lengths[i]) for i in range(len(lengths)))
        # each chunk's span fits the staging buffer
        for c in np.unique(rows[:, 5]):
            r = rows[rows[:, 5] == c]
            begin = int(r[:, 1].min()) // 16 * 16
            end = int((r[:, 1] + r[:, 2]).max())
            assert end - begin <= cap


def test_pipeline_chunks_ramp_up_and_taper_off(native):
    """The chunks sha512_shard cuts a shard into (csrc/snapgpu.cu: ramp_cap, taper_cap): 64 MiB, 256 MiB, then the
    staging size -- and the end of the shard as 128 MiB and 64 MiB, because nothing overlaps the hashing of the
    last chunk.  Every file in exactly one chunk, in order; short shards and unordered lists are cut sanely too."""
    MiB = 1 << 20

    def chunks_of(lengths, cap, offsets=None):
        lengths = np.asarray(lengths, dtype=np.uint64)
        if offsets is None:
            offsets = np.zeros(len(lengths), dtype=np.uint64)
            pos = 0
            for i, l in enumerate(lengths):
                offsets[i] = pos
                pos += (int(l) + 15) // 16 * 16
        rows = np.zeros((len(lengths) + 64, 6), dtype=np.uint64)
        n = native.lib().snapgpu_test_chunks(offsets.ctypes.data, lengths.ctypes.data, len(lengths), cap, 2,
                                             rows.ctypes.data, len(rows))
        assert n == len(lengths)                           # nothing here is larger than a chunk: no pieces
        rows = rows[:n]
        assert rows[:, 0].tolist() == list(range(len(lengths)))
        assert (np.diff(rows[:, 5].astype(np.int64)) >= 0).all()
        spans = []
        for c in np.unique(rows[:, 5]):
            r = rows[rows[:, 5] == c]
            spans.append(int((r[:, 1] + r[:, 2]).max()) - int(r[:, 1].min()))
        return spans

    rng = np.random.default_rng(3)
    lengths = rng.integers(1024, 65537, 36_000)            # ~1.2 GB, the size of config 2
    total = int(((lengths + 15) // 16 * 16).sum())
    spans = chunks_of(lengths, 1024 * MiB)
    assert len(spans) == 5 and sum(spans) <= total
    assert 63 * MiB < spans[0] <= 64 * MiB and 255 * MiB < spans[1] <= 256 * MiB
    assert 127 * MiB < spans[3] <= 129 * MiB and 63 * MiB < spans[4] <= 65 * MiB
    assert spans[2] == max(spans)
    # a staging size below the ramp: every chunk at most that size, and still the two short ones at the end
    spans = chunks_of(lengths, 200 * MiB)
    assert max(spans) <= 200 * MiB and 63 * MiB < spans[-1] <= 65 * MiB and 127 * MiB < spans[-2] <= 129 * MiB
    # shards shorter than the taper: one chunk, or the first 64 MiB and a last chunk of at most 96 MiB
    assert len(chunks_of(lengths[:1000], 1024 * MiB)) == 1
    spans = chunks_of(lengths[:4500], 1024 * MiB)           # ~150 MB
    assert len(spans) == 2 and spans[0] <= 64 * MiB and spans[1] <= 96 * MiB
    # files listed back to front: no estimate of what is left, the plain ramp
    lens = rng.integers(1024, 65537, 3000).astype(np.uint64)
    offs = np.zeros(len(lens), dtype=np.uint64)
    pos = 0
    for i in range(len(lens) - 1, -1, -1):
        offs[i] = pos
        pos += (int(lens[i]) + 15) // 16 * 16
    assert sum(1 for _ in chunks_of(lens, 1024 * MiB, offs)) >= 1


def test_long_file_bin_policy(native):
    """Which files of a launch leave the batched kernel for the long-file bin (csrc/snapgpu.cu: select_long_bin,
    pair_cta_shape) -- the files whose serial chain would set the launch's makespan:
    lane-pair form: at least 256 blocks (32 KiB), at least four times the launch's blocks per lane and within 31/64 of
    the longest file; with more than 4,096 files above the lower bound only those of 128 KiB and more are looked at;
    at most one CTA per SM of 16 files; one file per CTA up to a quarter of the SMs, two up to half, 16 beyond."""
    from snappy_b200 import synth

    def bin_of(lengths, sm=148, mode=2, min_blocks=0):
        lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        flags = np.zeros(len(lengths), dtype=np.uint8)
        per_cta = ctypes.c_uint32(0)
        native.check(native.lib().snapgpu_test_long_bin(lengths.ctypes.data, len(lengths), sm, mode, min_blocks,
                                                        flags.ctypes.data, ctypes.addressof(per_cta)))
        return flags.astype(bool), per_cta.value

    cfg2 = synth.lognormal_sizes(100_000)
    blocks = synth.blocks(cfg2)
    # config 3 as one batch: thousands of files above 32 KiB, four of 1 GiB -> exactly those four, one per CTA
    flags, per_cta = bin_of(np.concatenate([cfg2[:50_000], np.full(4, 1 << 30, dtype=np.uint64)]))
    assert np.nonzero(flags)[0].tolist() == [50_000, 50_001, 50_002, 50_003] and per_cta == 1
    # a small chain-bound batch: the files of 256+ blocks that are within 31/64 of the longest
    flags, per_cta = bin_of(cfg2[:3000])
    b = blocks[:3000]
    want = b >= max(256, int(b.max()) * 31 // 64)
    assert np.array_equal(flags, want) and 100 < flags.sum() < 600 and per_cta == 16
    # a launch that is bound by throughput keeps everything in the batched kernel
    assert not bin_of(cfg2)[0].any()
    assert not bin_of(np.concatenate([np.full(2000, 65536), np.full(400_000, 4096)]))[0].any()
    # nothing as long as 32 KiB: nothing to gain
    assert not bin_of(np.minimum(cfg2[:3000], 30_000))[0].any()
    # the one-lane form gains nothing below 128 KiB; long_min_blocks overrides either bound
    assert not bin_of(cfg2[:3000], mode=1)[0].any()
    assert bin_of(cfg2[:3000], min_blocks=1024)[0].sum() == 0 and bin_of(cfg2[:3000], min_blocks=400)[0].sum() > 0
    assert not bin_of(cfg2[:3000], mode=0)[0].any()
    # room: one CTA per SM of 16 files
    assert bin_of(np.full(2368, 130 * 1024))[0].all() and not bin_of(np.full(2369, 130 * 1024))[0].any()
    assert bin_of(np.full(32, 130 * 1024), sm=2)[0].all() and not bin_of(np.full(33, 130 * 1024), sm=2)[0].any()
    # placement: one file per CTA up to a quarter of the SMs, two up to half, then 16
    for n, shape in ((1, 1), (37, 1), (38, 2), (74, 2), (75, 16), (2000, 16)):
        assert bin_of(np.full(n, 1 << 20))[1] == shape, n
    # the few dominant files of a mixed launch, whatever their order in the list
    mixed = np.concatenate([np.full(5, 4 << 20), cfg2[:20_000], [7 << 20]]).astype(np.uint64)
    flags, per_cta = bin_of(mixed)
    assert np.nonzero(flags)[0].tolist() == [0, 1, 2, 3, 4, 20_005] and per_cta == 1
    # ... but not the ones that would finish in the batched kernel before the longest does on its lane pair (3 < 7 * 31/64)
    mixed[:5] = 3 << 20
    assert np.nonzero(bin_of(mixed)[0])[0].tolist() == [20_005]


# ---- copyToBuildDir (snappy/build.go:362-418): host logic that needs no GPU ---------------------

EXCLUDE_CASES = ["foo.snap", "foo.click", ".foo.swp", "..swp", ".swp", "foo~", "~", ",,x", ",x", ".#lock", ".~tmp",
                 ".bzr", ".bzrx", "x.bzr", ".git", ".gitignore", ".gitfoo", "CVS", "CVS2", "DEADJOE", "RCS", "_MTN",
                 "_darcs", "{arch}", "arch", ".hgtags", ".shelf", ".svn", ".arch-ids", "foo.snap.txt", "snap",
                 ".a.swo", ".a.sw", "a.swp", "foo.click~", ".bzr.backup", ".bzr.tags", ".bzr-builddeb", "normal.txt",
                 ".x\n.swp", ".x.sw\n", "foo~\n", "README"]


def test_should_exclude_matches_oracle(native, oracle):
    from snappy_b200 import build
    for name in EXCLUDE_CASES:
        assert build.shouldExclude(name) == oracle.should_exclude(name), repr(name)
    assert build.shouldExclude("foo~") and build.shouldExclude(".bzr") and not build.shouldExclude("README")


def snapshot(root):
    out = {}
    for dirpath, dirnames, filenames in os.walk(root):
        for n in dirnames + filenames:
            p = os.path.join(dirpath, n)
            st = os.lstat(p)
            rel = os.path.relpath(p, root)
            kind = "d" if stat.S_ISDIR(st.st_mode) else "l" if stat.S_ISLNK(st.st_mode) else "f"
            body = os.readlink(p) if kind == "l" else (open(p, "rb").read() if kind == "f" else None)
            out[rel] = (kind, stat.S_IMODE(st.st_mode), body, st.st_nlink > 1 if kind == "f" else None)
    return out


def make_source_tree(root):
    """makeExampleSnapSourceDir-like tree plus what the copy tests add (snappy/build_test.go:294-345)."""
    (root / "meta").mkdir(parents=True)
    (root / "meta" / "package.yaml").write_text("name: hello\n")
    (root / "bin").mkdir()
    (root / "bin" / "hello-world").write_text("#!/bin/sh\necho hello\n")
    os.chmod(root / "bin" / "hello-world", 0o755)
    (root / "bin" / "empty").write_bytes(b"")
    os.symlink("hello-world", root / "bin" / "link")
    (root / "foo~").write_text("hi")                       # TestCopyExcludesBackups
    (root / ".bzr").mkdir()                                # TestCopyExcludesWholeDirs
    (root / ".bzr" / "foo").write_text("hi")
    (root / "lib").mkdir(mode=0o750)
    (root / "lib" / "data.bin").write_bytes(bytes(range(256)) * 40)
    (root / "lib" / "x.click").write_text("excluded")
    os.chmod(root / "lib", 0o750)


def test_copy_to_build_dir_links_like_the_oracle(native, oracle, tmp_path):
    """TestCopyCopies / TestCopyExcludesBackups / TestCopyExcludesWholeDirs: same file system, so every
    file is hard-linked and no GPU is needed."""
    from snappy_b200 import build
    src = tmp_path / "src"
    make_source_tree(src)
    got, want = tmp_path / "got", tmp_path / "want"
    got.mkdir()                                            # an empty target is removed and re-created
    build.copyToBuildDir(str(src), str(got))
    oracle.copy_to_build_dir(str(src), str(want))
    assert snapshot(got) == snapshot(want)
    assert "foo~" not in snapshot(got) and ".bzr" not in snapshot(got) and "lib/x.click" not in snapshot(got)
    assert snapshot(got)["bin/hello-world"][3] is True     # linked, not copied
    # a non-empty target is an error in both (os.Remove fails, build.go:368-372)
    with pytest.raises(OSError):
        build.copyToBuildDir(str(src), str(got))
    with pytest.raises(OSError):
        oracle.copy_to_build_dir(str(src), str(want))
    # a missing source is the Walk error
    with pytest.raises(OSError):
        build.copyToBuildDir(str(tmp_path / "nope"), str(tmp_path / "t2"))
    # an excluded source directory copies nothing at all
    bak = tmp_path / "tree~"
    make_source_tree(bak)
    build.copyToBuildDir(str(bak), str(tmp_path / "t3"))
    assert not (tmp_path / "t3").exists()


def test_copy_to_build_dir_wide_tree_parallel_walk(native, oracle, tmp_path):
    """Many sub-directories: the root's subtrees are walked and linked by several threads; the result
    (including exclusions below the top level and an unreadable directory's error) equals the oracle's."""
    from snappy_b200 import build
    rng = np.random.default_rng(7)
    src = tmp_path / "src"
    src.mkdir()
    for d in range(40):
        sub = src / f"d{d:03d}"
        (sub / "deep" / "er").mkdir(parents=True)
        for f in range(25):
            (sub / f"f{f:02d}.bin").write_bytes(rng.bytes(int(rng.integers(0, 300))))
        (sub / "deep" / "x~").write_text("backup")
        (sub / "deep" / ".git").mkdir()
        (sub / "deep" / ".git" / "HEAD").write_text("ref")
        (sub / "deep" / "er" / "leaf").write_text(f"leaf {d}")
        os.symlink("../f00.bin", sub / "deep" / "ln")
    (src / "top.txt").write_text("top")
    (src / "CVS").mkdir()
    (src / "CVS" / "Entries").write_text("x")
    got, want = tmp_path / "got", tmp_path / "want"
    build.copyToBuildDir(str(src), str(got))
    oracle.copy_to_build_dir(str(src), str(want))
    a, b = snapshot(got), snapshot(want)
    assert a == b and len(a) == 40 * (25 + 5) + 1
    assert all(v[3] for v in a.values() if v[0] == "f")    # every file hard-linked
    if os.geteuid() != 0:                                  # root reads everything
        os.chmod(src / "d017" / "deep", 0)
        try:
            with pytest.raises(OSError) as e1:
                build.copyToBuildDir(str(src), str(tmp_path / "g2"))
            with pytest.raises(OSError) as e2:
                oracle.copy_to_build_dir(str(src), str(tmp_path / "w2"))
            assert "d017/deep" in str(e1.value) and "d017/deep" in str(e2.value)
        finally:
            os.chmod(src / "d017" / "deep", 0o755)


def test_missing_archive_is_reported_before_anything_else(native, oracle, tmp_path):
    """writeHashes makes DEBIAN/, hashes the archive, then walks (snappy/build.go:218-228): with no
    archive the open error comes first -- before a walk error, before the GPU is touched -- and
    DEBIAN/ exists afterwards in both implementations."""
    from snappy_b200 import build
    for impl in (lambda t, a: build.writeHashes(t, a), lambda t, a: oracle.write_hashes(t, a)):
        tree = tmp_path / f"t{id(impl)}"
        tree.mkdir()
        os.mkfifo(tree / "pipe", 0o644)                    # would be "Unknown file mode" during the walk
        with pytest.raises(OSError) as e:
            impl(str(tree), str(tmp_path / "no-such.tar.gz"))
        assert "no-such.tar.gz" in str(e.value)
        assert (tree / "DEBIAN").is_dir()
        assert not (tree / "DEBIAN" / "hashes.yaml").exists()


def test_read_archive_sha512_like_new_snap_part(native, golden_dir, tmp_path):
    """The reader side (snappy/snapp.go:466-478): TestLocalSnapHash (snappy/snapp_test.go:159-170), the
    "{}" fixture of makeInstalledMockSnap (snappy/common_test.go:77), the golden document, and the mode
    decoding of yamlFileMode.UnmarshalYAML (snappy/hashes.go:59-88)."""
    from snappy_b200 import build
    f = tmp_path / "hashes.yaml"
    f.write_bytes(b"archive-sha512: F00F00")                     # no trailing newline, as the reference writes it
    assert build.readArchiveSha512(str(f)) == "F00F00"
    f.write_bytes(b"{}")
    assert build.readArchiveSha512(str(f)) == ""
    f.write_bytes(b"")
    assert build.readArchiveSha512(str(f)) == ""
    assert build.readArchiveSha512(str(golden_dir / "hashes_simple.yaml")) == (
        "cf83e1357eefb8bdf1542850d66d8007d620e4050b5715dc83f4a921d36ce9ce47d0d13c5d85f2b0ff8318d2877eec2f63b931bd47417a81a538327af927da3e")
    f.write_bytes(b'archive-sha512: "12345"\nfiles:\n- name: a\n  mode: drwxr-xr-x\n- name: b\n  size: 0\n  sha512: 00\n  mode: frw-r--r--\n')
    assert build.readArchiveSha512(str(f)) == "12345"           # a quoted (number-like) scalar is unquoted
    f.write_bytes(b"archive-sha512: ab\nfiles:\n- name: a\n  mode: prw-r--r--\n")
    with pytest.raises(build.UnknownFileMode) as e:
        build.readArchiveSha512(str(f))
    assert "Unknown file mode prw-r--r--" in str(e.value)
    with pytest.raises(OSError):
        build.readArchiveSha512(str(tmp_path / "missing.yaml"))


def test_filehash_fragment_through_the_product_emitter(native, golden_dir):
    """TestHashesYamlMarshal (snappy/hashes_test.go:30-55) on the C++ emitter of libsnapgpu: the
    fileHash {Name: "foo", Size: 10, Mode: dir 0644} renders as the reference's fragment, omitempty
    drops a nil size and an empty sha512, and the mode string follows yamlFileMode.MarshalYAML."""
    import ctypes
    import stat

    def render(name, size, sha, mode):
        ptr, ln = ctypes.c_void_p(), ctypes.c_size_t()
        native.check(native.lib().snapgpu_test_filehash_yaml(name.encode(), size, sha.encode() if sha else None, mode,
                                                             ctypes.byref(ptr), ctypes.byref(ln)))
        return native.take_string(ptr, ln.value)

    assert render("foo", 10, None, stat.S_IFDIR | 0o644) == (golden_dir / "filehash_fragment.yaml").read_bytes()
    assert render("foo", -1, None, stat.S_IFLNK | 0o777) == b"name: foo\nmode: lrwxrwxrwx\n"
    assert render("bin/bar", 4, "cc" * 64, stat.S_IFREG | 0o644) == \
        b"name: bin/bar\nsize: 4\nsha512: " + b"cc" * 64 + b"\nmode: frw-r--r--\n"
    assert render("true", 0, None, stat.S_IFREG | 0o600) == b'name: "true"\nsize: 0\nmode: frw-------\n'
    ptr, ln = ctypes.c_void_p(), ctypes.c_size_t()
    rc = native.lib().snapgpu_test_filehash_yaml(b"pipe", -1, None, stat.S_IFIFO | 0o644, ctypes.byref(ptr), ctypes.byref(ln))
    assert rc == native.EMODE and "Unknown file mode" in native.last_error()
