"""The host side of libsnapgpu (tree walk, packer pool, chunk recycling, chain streamer, YAML
writer, copyToBuildDir, the digest cache) run WITHOUT a GPU against a fake runtime, under
AddressSanitizer + UndefinedBehaviorSanitizer and under ThreadSanitizer (tests/hostsim/).

The fake runtime takes its SHA-512 from the CPU oracle -- it exists only in this test build.
Documents are compared with the oracle's write_hashes (snappy/build.go:216-270 restated) and
with the reference's golden document (snappy/hashes_test.go:89-103).  The same entry points
run on the real kernels in tests/test_gpu_parity.py.
"""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, make_reference_tree

SIM = ROOT / "tests" / "hostsim"


@pytest.fixture(scope="module")
def sim():
    subprocess.check_call(["make", "-s", "-C", str(SIM)])
    return {"asan": SIM / "_build" / "hostsim_asan", "tsan": SIM / "_build" / "hostsim_tsan"}


def run(exe, *args, env=None, ok=True):
    e = dict(os.environ)
    e["ASAN_OPTIONS"] = "detect_leaks=1:abort_on_error=0"
    e["UBSAN_OPTIONS"] = "halt_on_error=1:print_stacktrace=1"
    e["TSAN_OPTIONS"] = "halt_on_error=0:exitcode=66"
    e.update(env or {})
    p = subprocess.run([str(exe), *map(str, args)], capture_output=True, env=e, timeout=120)
    assert b"ERROR: AddressSanitizer" not in p.stderr and b"runtime error" not in p.stderr and \
        b"WARNING: ThreadSanitizer" not in p.stderr and b"LeakSanitizer" not in p.stderr, p.stderr.decode(errors="replace")[-4000:]
    if ok:
        assert p.returncode == 0, (p.returncode, p.stdout[-500:], p.stderr[-2000:])
    return p


def make_mixed_tree(root: Path, rng, nfiles=700):
    """Nested directories, empty ones, symlinks, names yaml.v2 quotes, DEBIAN-prefixed names and
    every size class of the packer (small <= 512 KiB, mid <= 16 MiB, a chain beyond)."""
    root.mkdir()
    dirs = [root]
    for i in range(40):
        d = dirs[int(rng.integers(len(dirs)))] / f"dir{i:02d}"
        d.mkdir()
        dirs.append(d)
    (root / "empty").mkdir()
    (root / "DEBIAN").mkdir()
    (root / "DEBIAN" / "control").write_bytes(b"skipped")
    (root / "DEBIANfoo").write_bytes(b"also skipped: the prefix rule of build.go:229")
    (dirs[1] / "DEBIAN").mkdir(exist_ok=True)                    # only the top level is special
    (dirs[1] / "DEBIAN" / "kept").write_bytes(b"kept")
    sizes = [0, 1, 111, 112, 127, 128, 129, 4096, 65536, (512 << 10) - 1, 512 << 10, (512 << 10) + 1,
             5 << 20, (16 << 20) + 5]
    sizes += [int(s) for s in np.clip(np.rint(np.exp(rng.normal(np.log(6000.0), 1.3, size=nfiles))), 0, 300000)]
    for i, size in enumerate(sizes):
        d = dirs[int(rng.integers(len(dirs)))]
        p = d / f"f{i:05d}.bin"
        p.write_bytes(rng.integers(0, 256, size=size, dtype=np.uint8).tobytes())
        os.chmod(p, [0o644, 0o755, 0o600, 0o444][i % 4])
    for name in ["true", "123", "1e3", "~", "a b", "x: y", "-", "# c", "Grüße", "tab\there", "null", "0x1f", "1:30"]:
        (dirs[2] / name).write_bytes(name.encode())
    os.symlink("f00000.bin", root / "link-to-file")
    os.symlink("dir00", root / "link-to-dir")
    os.symlink("/nonexistent/target", root / "dangling")
    big = root / "wide"                                          # one directory cut into several pack tasks
    big.mkdir()
    for i in range(400):
        (big / f"w{i:04d}").write_bytes(rng.integers(0, 256, size=int(rng.integers(0, 3000)), dtype=np.uint8).tobytes())


def test_golden_document(sim, golden_dir, tmp_path):
    tree = tmp_path / "tree"
    tree.mkdir()
    make_reference_tree(tree)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(b"")
    got = run(sim["asan"], "hashes_yaml", tree, tar).stdout
    assert got == (golden_dir / "hashes_simple.yaml").read_bytes()
    run(sim["asan"], "write_hashes", tree, tar)
    f = tree / "DEBIAN" / "hashes.yaml"
    assert f.read_bytes() == got and (f.stat().st_mode & 0o777) == 0o644


@pytest.mark.parametrize("flavour", ["asan", "tsan"])
def test_mixed_tree_matches_the_oracle(sim, oracle, tmp_path, flavour):
    rng = np.random.default_rng(7)
    tree = tmp_path / "tree"
    make_mixed_tree(tree, rng)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(rng.integers(0, 256, size=(5 << 20) + 77, dtype=np.uint8).tobytes())
    want = oracle.write_hashes(str(tree), str(tar))
    # ThreadSanitizer keys its descriptor bookkeeping by number and so reports "races" between
    # threads whose tables are private; its run shares one table (the locking is the same)
    env = {"SNAPGPU_SHARED_FDS": "1"} if flavour == "tsan" else {}
    for threads in ("3", "16"):
        got = run(sim[flavour], "repeat", 2, tree, tar, env={**env, "SNAPGPU_PACK_THREADS": threads}).stdout
        assert got == want


@pytest.mark.parametrize("flavour", ["asan", "tsan"])
def test_warm_up_beside_a_running_write_hashes(sim, oracle, tmp_path, flavour):
    """snapgpu_warm from a second thread while the first is already inside writeHashes (both grow the chunk pool,
    both open a session), and again afterwards: same document, nothing for the sanitizers to report."""
    rng = np.random.default_rng(17)
    tree = tmp_path / "tree"
    make_mixed_tree(tree, rng)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(rng.integers(0, 256, size=(1 << 20) + 5, dtype=np.uint8).tobytes())
    env = {"SNAPGPU_SHARED_FDS": "1"} if flavour == "tsan" else {}
    assert run(sim[flavour], "warm", tree, tar, env=env).stdout == oracle.write_hashes(str(tree), str(tar))


def test_empty_tree_and_missing_root(sim, oracle, tmp_path):
    tar = tmp_path / "t"
    tar.write_bytes(b"x")
    empty = tmp_path / "empty"
    empty.mkdir()
    assert run(sim["asan"], "hashes_yaml", empty, tar).stdout == oracle.write_hashes(str(empty), str(tar))
    p = run(sim["asan"], "hashes_yaml", tmp_path / "tree", tmp_path / "no-such.tar.gz", ok=False)
    assert p.returncode == 3 and b"no-such.tar.gz: No such file or directory" in p.stdout


def test_unknown_file_mode_and_walk_order_of_errors(sim, tmp_path):
    tree = tmp_path / "tree"
    (tree / "a").mkdir(parents=True)
    (tree / "a" / "ok").write_bytes(b"1")
    os.mkfifo(tree / "a" / "pipe")
    os.mkfifo(tree / "b-pipe")
    tar = tmp_path / "t"
    tar.write_bytes(b"")
    p = run(sim["asan"], "hashes_yaml", tree, tar, ok=False)
    assert p.returncode == 3 and p.stdout.startswith(b"ERR -4 Unknown file mode p")


def test_injected_batch_failure_does_not_hang(sim, tmp_path):
    rng = np.random.default_rng(3)
    tree = tmp_path / "tree"
    make_mixed_tree(tree, rng, nfiles=3000)
    tar = tmp_path / "t"
    tar.write_bytes(b"abc")
    for nth in ("1", "2"):
        p = run(sim["asan"], "hashes_yaml", tree, tar, env={"HOSTSIM_FAIL_SUBMIT": nth, "SNAPGPU_PACK_THREADS": "8"}, ok=False)
        assert p.returncode == 3 and b"injected batch failure" in p.stdout


def test_copy_then_write_hashes_and_the_digest_cache(sim, oracle, tmp_path):
    """copyToBuildDir reads each copied file once and remembers its digest for the writeHashes that
    follows; a file edited in place afterwards -- same size, mtime put back -- is hashed again
    (the cache compares ctime too), and the cache does not outlive that writeHashes."""
    rng = np.random.default_rng(11)
    src = tmp_path / "src"
    make_mixed_tree(src, rng, nfiles=200)
    (src / "DEBIAN" / "control").unlink()
    (src / "DEBIAN").rmdir()
    (src / "DEBIANfoo").unlink()
    for link in ("link-to-file", "link-to-dir", "dangling"):     # with the copy forced, os.Open would follow them
        (src / link).unlink()
    tar = tmp_path / "t"
    tar.write_bytes(b"tarball")
    dst = tmp_path / "dst"
    got = run(sim["asan"], "copy", src, dst, 1, tar).stdout
    ref = tmp_path / "ref"
    oracle.copy_to_build_dir(str(src), str(ref), no_link=True)   # shouldExclude drops "~" here too
    want = oracle.write_hashes(str(ref), str(tar))
    assert got == want
    assert got == oracle.write_hashes(str(dst), str(tar))        # and it describes what was written
    # in one process: copy, edit a copied file in place (same size, mtime put back), writeHashes
    dst2 = tmp_path / "dst2"
    victim = "wide/w0007"
    got = run(sim["asan"], "copy_edit", src, dst2, tar, victim).stdout
    assert got != want
    assert got == oracle.write_hashes(str(dst2), str(tar))


def test_verify_reports(sim, oracle, tmp_path):
    rng = np.random.default_rng(5)
    tree = tmp_path / "tree"
    make_mixed_tree(tree, rng, nfiles=100)
    tar = tmp_path / "t"
    tar.write_bytes(b"tar")
    run(sim["asan"], "write_hashes", tree, tar)
    doc = tmp_path / "hashes.yaml"
    doc.write_bytes((tree / "DEBIAN" / "hashes.yaml").read_bytes())
    for f in (tree / "DEBIAN").iterdir():
        f.unlink()
    (tree / "DEBIAN").rmdir()
    assert run(sim["asan"], "verify", tree, doc, tar).stdout == b""
    (tree / "wide" / "w0003").write_bytes(b"changed")
    (tree / "wide" / "w0004").unlink()
    (tree / "new").write_bytes(b"n")
    got = run(sim["asan"], "verify", tree, doc).stdout.decode().splitlines()
    assert sorted(got) == sorted(oracle.verify_hashes(str(tree), str(doc)))


def test_compare_path_and_one_file_entry_points(sim, oracle, golden_dir, tmp_path):
    """DirUpdated (helpers/cmp.go:97-114) and Sha512sum of one file through the same sanitizer build:
    the cases of helpers/cmp_test.go:84-133 plus sizes across the 16 KiB chunk and the pinned staging."""
    import json
    rng = np.random.default_rng(9)
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir()
    b.mkdir()
    for i, n in enumerate([0, 1, 16383, 16384, 16385, 100_000, 3_000_000]):
        body = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        (a / f"same{i}").write_bytes(body)
        (b / f"same{i}").write_bytes(body)
        if n:
            other = bytearray(body)
            other[n // 2] ^= 1
            (a / f"diff{i}").write_bytes(body)
            (b / f"diff{i}").write_bytes(bytes(other))
    (a / "only-in-a").write_bytes(b"x")
    (b / "only-in-b").write_bytes(b"y")
    (a / "size").write_bytes(b"12")
    (b / "size").write_bytes(b"123")
    (a / "subdir").mkdir()
    (b / "subdir").mkdir()
    got = run(sim["asan"], "dir_updated", a, b, "pfx_").stdout.decode().splitlines()
    assert sorted(got) == sorted(oracle.dir_updated(str(a), str(b), "pfx_"))
    for k in json.loads((golden_dir / "sha512_kats.json").read_text()):
        f = tmp_path / "kat"
        f.write_bytes(k["message"].encode())
        assert run(sim["asan"], "sha512sum", f).stdout.decode().strip() == k["sha512"]
    p = run(sim["asan"], "sha512sum", tmp_path / "missing", ok=False)
    assert p.returncode == 3 and b"No such file or directory" in p.stdout
