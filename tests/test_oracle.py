"""The oracle against the reference's own golden vectors and independent implementations.

This is what pins the oracle before any GPU result is compared with it (task section 3).
"""
import hashlib
import json
import os
import stat

import numpy as np
import pytest

from conftest import make_reference_tree


def test_reference_kats(oracle, golden_dir):
    kats = json.loads((golden_dir / "sha512_kats.json").read_text())
    assert len(kats) == 4
    for k in kats:
        assert oracle.sha512(k["message"].encode()).hex() == k["sha512"], k["source"]


def test_sha512_every_length_vs_hashlib(oracle):
    rng = np.random.default_rng(1)
    for n in list(range(0, 301)) + [1023, 1024, 4096, 32767, 32768, 32769, 65536, 100003]:
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.sha512(d) == hashlib.sha512(d).digest(), n


def test_padding_boundary(oracle):
    # 111 bytes is the last length whose padding fits the same block, 112 needs one more
    for n in (111, 112, 127, 128, 239, 240):
        d = bytes([n & 0xFF]) * n
        assert oracle.sha512(d) == hashlib.sha512(d).digest()


def test_batch_threads_and_openssl(oracle):
    rng = np.random.default_rng(2)
    lengths = rng.integers(0, 5000, 257).astype(np.uint64)
    offsets = np.concatenate([[0], np.cumsum(lengths[:-1] + 3)]).astype(np.uint64)
    data = rng.integers(0, 256, int(offsets[-1] + lengths[-1]) + 8, dtype=np.uint8)
    want = np.stack([np.frombuffer(hashlib.sha512(data[int(o):int(o + l)].tobytes()).digest(), dtype=np.uint8)
                     for o, l in zip(offsets, lengths)])
    for threads, ossl in ((1, False), (4, False), (1, True), (3, True)):
        got = oracle.sha512_batch(data, offsets, lengths, threads, ossl)
        assert np.array_equal(got, want), (threads, ossl)


def test_sha512sum_file(oracle, tmp_path):
    p = tmp_path / "foo"
    p.write_bytes(b"x")
    assert oracle.sha512sum(str(p)) == hashlib.sha512(b"x").hexdigest()
    with pytest.raises(OSError):
        oracle.sha512sum(str(tmp_path / "missing"))


def test_golden_hashes_yaml(oracle, golden_dir, tmp_path):
    """TestBuildCreateDebianHashesSimple (snappy/hashes_test.go:57-104)."""
    tree = tmp_path / "tree"
    tree.mkdir()
    make_reference_tree(tree)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(b"")
    got = oracle.write_hashes(str(tree), str(tar))
    assert got == (golden_dir / "hashes_simple.yaml").read_bytes()
    assert (tree / "DEBIAN" / "hashes.yaml").read_bytes() == got
    assert stat.S_IMODE(os.stat(tree / "DEBIAN" / "hashes.yaml").st_mode) == 0o644


def test_golden_filehash_fragment(oracle, golden_dir):
    """TestHashesYamlMarshal / Unmarshal (snappy/hashes_test.go:30-55)."""
    mode = oracle.file_mode_string(0o644 | stat.S_IFDIR)
    assert mode == "drw-r--r--"
    got = oracle.marshal_single_file_hash({"name": b"foo", "size": 10, "sha512": "", "mode": mode})
    assert got == (golden_dir / "filehash_fragment.yaml").read_bytes()
    back = oracle.parse_mode_string(mode)
    assert stat.S_ISDIR(back) and stat.S_IMODE(back) == 0o644


def test_mode_strings(oracle):
    assert oracle.file_mode_string(stat.S_IFREG | 0o644) == "frw-r--r--"
    assert oracle.file_mode_string(stat.S_IFREG | 0o4755) == "frwxr-xr-x"     # setuid not represented
    assert oracle.file_mode_string(stat.S_IFLNK | 0o777) == "lrwxrwxrwx"
    with pytest.raises(oracle.UnknownFileMode) as e:
        oracle.file_mode_string(stat.S_IFIFO | 0o644)
    assert str(e.value) == "Unknown file mode prw-r--r--"


def test_yaml_matches_pyyaml_on_safe_names(oracle):
    import yaml
    files = [{"name": n.encode(), "size": s, "sha512": h, "mode": "frw-r--r--"}
             for n, s, h in (("d0000/f0000001.bin", 4096, "ab" * 64), ("usr/lib/x86_64/libfoo.so.1", 0, "0" * 128))]
    files.append({"name": b"usr", "size": None, "sha512": "", "mode": "drwxr-xr-x"})
    got = oracle.marshal_hashes("cd" * 64, files)
    doc = {"archive-sha512": "cd" * 64, "files": [
        {k: (v.decode() if isinstance(v, bytes) else v) for k, v in f.items() if v not in (None, "")} for f in files]}
    want = yaml.dump(doc, default_flow_style=False, sort_keys=False, width=10**9).encode()
    # PyYAML quotes the all-digit digest; yaml.v2 double-quotes it too ("0"*128 parses as a float)
    assert yaml.safe_load(got) == yaml.safe_load(want)
    assert got.replace(b'"', b"'") == want.replace(b'"', b"'")


def test_yaml_quoting_roundtrips(oracle):
    """Names that yaml.v2 must quote: parse back (PyYAML) to the same string."""
    import yaml
    names = ["true", "123", "1e3", "~", "null", "#x", "a: b", "a #b", " lead", "trail ", "- x", "-x", "0x1f",
             "1_000", "12:30", "it's", "tab\there", "multi\nline", "café", "x" * 70 + " " + "y" * 30 + " z",
             "\U0001F600", "[abc]", "{a}", "*star", "&amp", "!bang", "%pct", "@at", "`tick", "a\\b", 'q"uote',
             "..." , "---", "?", ":", "y", "No", ".5", "+.inf", "0b101", "<<", "080", "1.", "+1", "e5"]
    for n in names:
        doc = oracle.marshal_single_file_hash({"name": n.encode(), "size": None, "sha512": "", "mode": "frw-r--r--"})
        back = yaml.safe_load(doc.decode())
        assert back == {"name": n, "mode": "frw-r--r--"}, (n, doc)


def test_empty_tree(oracle, tmp_path):
    tree = tmp_path / "t"
    tree.mkdir()
    tar = tmp_path / "d.tar.gz"
    tar.write_bytes(b"abc")
    got = oracle.write_hashes(str(tree), str(tar))
    assert got == b"archive-sha512: " + hashlib.sha512(b"abc").hexdigest().encode() + b"\nfiles: []\n"


def test_debian_prefix_rule(oracle, tmp_path):
    """build.go:229 is a string-prefix test: DEBIAN-x and DEBIANfoo/* are skipped as well."""
    tree = tmp_path / "t"
    (tree / "DEBIAN-x").mkdir(parents=True)
    (tree / "DEBIANfoo").mkdir()
    (tree / "DEBIANfoo" / "f").write_bytes(b"1")
    (tree / "debian").mkdir()
    (tree / "debian" / "f").write_bytes(b"2")
    tar = tmp_path / "d"
    tar.write_bytes(b"")
    _, entries = oracle.collect_hashes(str(tree), str(tar))
    assert [e["name"] for e in entries] == [b"debian", b"debian/f"]


# ---- cmp: the reference's truth tables (helpers/cmp_test.go) --------------------------------

def test_cmp_self_across_chunk_boundary(oracle, tmp_path):
    foo = tmp_path / "foo"
    with open(foo, "wb") as f:
        for _ in range(1100):
            f.flush()
            assert oracle.files_are_equal(str(foo), str(foo))
            f.write(b"*" * 16)


def test_cmp_empty_missing_nonempty(oracle, tmp_path):
    foo, bar = tmp_path / "foo", tmp_path / "bar"
    foo.write_bytes(b"")
    assert not oracle.files_are_equal(str(foo), str(bar))
    assert not oracle.files_are_equal(str(bar), str(foo))
    bar.write_bytes(b"x")
    assert not oracle.files_are_equal(str(foo), str(bar))
    assert not oracle.files_are_equal(str(bar), str(foo))


def test_cmp_streams(oracle):
    assert oracle.streams_equal(b"hello", b"hello")
    assert not oracle.streams_equal(b"hello", b"world")
    assert not oracle.streams_equal(b"hello", b"hell")
    assert oracle.streams_equal(b"", b"")
    a = os.urandom(16384 * 2 + 5)
    assert oracle.streams_equal(a, a)
    assert not oracle.streams_equal(a, a[:-1])
    assert not oracle.streams_equal(a[:16384], a[:16385])
    b = bytearray(a)
    b[16384] ^= 1
    assert not oracle.streams_equal(a, bytes(b))


def test_dir_updated(oracle, tmp_path):
    d1, d2 = tmp_path / "d1", tmp_path / "d2"
    d1.mkdir()
    d2.mkdir()
    assert oracle.dir_updated(str(d1), str(d2), "") == {}
    (d2 / "foo").write_bytes(b"x")
    assert oracle.dir_updated(str(d1), str(d2), "") == {}
    assert oracle.dir_updated(str(d2), str(d1), "") == {}
    (d1 / "foo").write_bytes(b"x")
    assert oracle.dir_updated(str(d1), str(d2), "") == {}
    (d1 / "dir").mkdir()
    assert oracle.dir_updated(str(d1), str(d2), "") == {}
    (d1 / "foo").write_bytes(b"y")
    (d1 / "bar").write_bytes(b"x")
    (d2 / "bar").write_bytes(b"y")
    (d2 / "baz").write_bytes(b"x")
    assert oracle.dir_updated(str(d1), str(d2), "") == {"bar": True, "foo": True}
    assert oracle.dir_updated(str(d1), str(d2), "foo_") == {"foo_bar": True, "foo_foo": True}


def test_apparmor_delta(oracle, tmp_path):
    """policy/policy_test.go:164-182."""
    orig, dest = tmp_path / "orig", tmp_path / "dest"
    for root, suffix in ((orig, ""), (dest, " 2")):
        base = root / "meta" / "framework-policy" / "apparmor" / "policygroups"
        base.mkdir(parents=True)
        for k in range(3):
            (base / f"policygroups{k}").write_text(f"apparmor::policygroups{k}{suffix}")
    (orig / "meta" / "framework-policy" / "apparmor" / "templates").mkdir()
    (orig / "meta" / "framework-policy" / "apparmor" / "templates" / "t0").write_text("x")
    ps, ts = oracle.apparmor_delta(str(orig), str(dest), "x-")
    assert ps == {"x-policygroups0": True, "x-policygroups1": True, "x-policygroups2": True}
    assert ts == {}


def test_sha512sum_files_threaded_equals_hashlib(oracle, tmp_path):
    """The CPU tree comparator of bench.py (Sha512sum over a file list, 1 and several threads, both
    block functions) against hashlib, incl. an empty file, a >32 KiB file and a missing one."""
    import hashlib
    rng = np.random.default_rng(1)
    paths, sizes = [], []
    for i, n in enumerate([0, 1, 111, 112, 4096, 32768, 32769, 100_000] + [int(x) for x in rng.integers(0, 5000, 40)]):
        p = tmp_path / f"f{i:03d}"
        p.write_bytes(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        paths.append(str(p))
        sizes.append(n)
    want = np.array([list(hashlib.sha512(open(p, "rb").read()).digest()) for p in paths], dtype=np.uint8)
    for threads in (1, 3, 8):
        for ossl in (False, True):
            assert np.array_equal(oracle.sha512sum_files(paths, sizes, threads, ossl), want)
    with pytest.raises(OSError):
        oracle.sha512sum_files(paths + [str(tmp_path / "missing")], sizes + [0], 2)
