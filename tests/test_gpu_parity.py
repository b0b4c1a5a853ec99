"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs.  Bit-exact is the bar for digests, flags and emitted YAML."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import make_reference_tree

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def default_options(gpu):
    gpu.set_option("staging_bytes", 256 << 20)
    gpu.set_option("sha_variant", 0)
    gpu.set_option("sha_warps_per_sm", 0)
    gpu.set_option("long_kernel", 2)
    gpu.set_option("pair_form", 0)
    gpu.set_option("pair_files_per_cta", 0)
    gpu.set_option("long_min_blocks", 0)
    gpu.set_option("two_ended", 1)
    yield


def pack(lengths, rng, align=16, jitter=False):
    lengths = np.asarray(lengths, dtype=np.uint64)
    offsets = np.zeros(len(lengths), dtype=np.uint64)
    pos = 0
    for i, l in enumerate(lengths):
        pos = (pos + align - 1) // align * align
        if jitter:
            pos += int(rng.integers(0, 16))
        offsets[i] = pos
        pos += int(l)
    data = rng.integers(0, 256, pos + 64, dtype=np.uint8)
    return data, offsets, lengths


# ---- SHA-512 -----------------------------------------------------------------------------------

def test_reference_kats_through_files(gpu, golden_dir, tmp_path):
    """TestSha512sum (helpers/helpers_test.go:167-176) and the other three KATs."""
    from snappy_b200 import helpers
    for i, k in enumerate(json.loads((golden_dir / "sha512_kats.json").read_text())):
        p = tmp_path / f"kat{i}"
        p.write_bytes(k["message"].encode())
        assert helpers.Sha512sum(str(p)) == k["sha512"], k["source"]
    with pytest.raises(OSError) as e:
        helpers.Sha512sum(str(tmp_path / "missing"))
    assert "no such file or directory" in str(e.value).lower()


def test_reference_kats_through_the_hasher(gpu, golden_dir):
    """The other crypto/sha512 call shape in the reference: sha512.New(); Write; hex(Sum(nil))
    (SystemImagePart.Hash, snappy/systemimage.go:122-128; its KAT is snappy/systemimage_test.go:104,117)."""
    from snappy_b200 import helpers
    for k in json.loads((golden_dir / "sha512_kats.json").read_text()):
        h = helpers.Sha512Stream()
        h.Write(k["message"].encode())
        assert h.Sum().hex() == k["sha512"], k["source"]
        msg = k["message"].encode()                      # the same in two writes
        h2 = helpers.Sha512Stream()
        h2.Write(msg[: len(msg) // 2])
        h2.Write(msg[len(msg) // 2:])
        assert h2.Sum().hex() == k["sha512"]


def test_every_length_0_to_300(gpu, oracle):
    from snappy_b200 import helpers
    rng = np.random.default_rng(11)
    data, off, ln = pack(np.arange(0, 301), rng)
    got = helpers.sha512_batch(data, off, ln)
    assert np.array_equal(got, oracle.sha512_batch(data, off, ln))
    for i in (0, 1, 111, 112, 128, 300):
        assert got[i].tobytes() == hashlib.sha512(data[int(off[i]):int(off[i] + ln[i])].tobytes()).digest()


@pytest.mark.parametrize("variant", list(range(6)))
@pytest.mark.parametrize("warps", [0, 1, 2, 3])
def test_kernel_variants_bit_exact(gpu, oracle, variant, warps):
    from snappy_b200 import helpers
    gpu.set_option("sha_variant", variant)
    gpu.set_option("sha_warps_per_sm", warps)
    rng = np.random.default_rng(100 + variant)
    lengths = np.concatenate([rng.integers(0, 9000, 700), [65536, 65535, 40000, 0, 0, 111, 112, 239, 240]])
    data, off, ln = pack(lengths, rng)
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 4))


def test_device_plan_is_longest_first(gpu):
    """Length binning (plan_kernels.cuh): a permutation, non-increasing in min(blocks, 65535)."""
    rng = np.random.default_rng(3)
    for extra in ([0, 111, 112, 2**30, 2**31 + 5, 9_000_000, 8_388_608 * 128], []):
        for n in (1, 31, 33, 5000, 70_001):
            lengths = np.concatenate([rng.integers(0, 70000, n), extra]).astype(np.uint64)
            order = np.zeros(len(lengths), dtype=np.uint32)
            gpu.check(gpu.lib().snapgpu_test_plan_order(lengths.ctypes.data, len(lengths), order.ctypes.data))
            assert sorted(order.tolist()) == list(range(len(lengths)))
            blocks = np.minimum((lengths[order] + np.uint64(144)) // np.uint64(128), np.uint64(65535)).astype(np.int64)
            assert np.all(blocks[:-1] >= blocks[1:])
    # uniform lengths (config 5): one bucket, every warp-aggregated atomic hits the same counter
    lengths = np.full(100_000, 65536, dtype=np.uint64)
    order = np.zeros(len(lengths), dtype=np.uint32)
    gpu.check(gpu.lib().snapgpu_test_plan_order(lengths.ctypes.data, len(lengths), order.ctypes.data))
    assert sorted(order.tolist()) == list(range(len(lengths)))


def test_unaligned_offsets_take_the_generic_path(gpu, oracle):
    from snappy_b200 import helpers
    rng = np.random.default_rng(12)
    lengths = np.concatenate([rng.integers(0, 3000, 400), np.arange(0, 40)])
    data, off, ln = pack(lengths, rng, align=1, jitter=True)
    assert (off % 16 != 0).any()
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 4))
    # and a misaligned base pointer
    backing = np.zeros(len(data) + 32, np.uint8)
    start = (data.ctypes.data + 5 - backing.ctypes.data) % 16         # base pointer 5 bytes off data's phase
    shifted = backing[start:start + len(data)]
    shifted[:] = data
    assert shifted.ctypes.data % 16 != data.ctypes.data % 16
    assert np.array_equal(helpers.sha512_batch(shifted, off, ln), oracle.sha512_batch(data, off, ln, 4))


def test_config1_1000_files_of_4k(gpu, oracle):
    """BASELINE config 1: 1,000 files x 4 KiB."""
    from snappy_b200 import helpers, synth
    data, off, ln = synth.make_host_batch(np.full(1000, 4096, dtype=np.uint64))
    got = helpers.sha512_batch(data, off, ln)
    assert np.array_equal(got, oracle.sha512_batch(data, off, ln, 4))
    assert got[7].tobytes() == hashlib.sha512(synth.file_bytes(7, 4096)).digest()


def test_lognormal_batch_and_checksum_of_checksums(gpu, oracle):
    """A slice of config 2 (log-normal 1-64 KiB) and an order-independence property."""
    from snappy_b200 import helpers, synth
    lengths = synth.lognormal_sizes(6000)
    data, off, ln = synth.make_host_batch(lengths)
    got = helpers.sha512_batch(data, off, ln)
    assert np.array_equal(got, oracle.sha512_batch(data, off, ln, 8, True))
    perm = np.random.default_rng(5).permutation(len(ln))
    got_perm = helpers.sha512_batch(data, off[perm], ln[perm])
    assert np.array_equal(got_perm, got[perm])
    assert hashlib.sha512(got.tobytes()).digest() == hashlib.sha512(got_perm[np.argsort(perm)].tobytes()).digest()


def test_multi_chunk_pipeline_and_streamed_big_files(gpu, oracle):
    """Small staging buffer: many chunks, plus files larger than the buffer that must be
    hashed as continuation segments."""
    from snappy_b200 import helpers
    gpu.set_option("staging_bytes", 1 << 20)
    rng = np.random.default_rng(13)
    lengths = np.concatenate([rng.integers(0, 50000, 300), [3_000_001, 1 << 20, (1 << 20) - 64, 2_500_000 + 128 * 7]])
    lengths = lengths[rng.permutation(len(lengths))]
    data, off, ln = pack(lengths, rng)
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 8, True))


def test_stream_api_matches_one_shot(gpu):
    from snappy_b200 import helpers
    rng = np.random.default_rng(14)
    msg = rng.integers(0, 256, 700_001, dtype=np.uint8).tobytes()
    h = helpers.Sha512Stream()
    pos = 0
    for step in (1, 127, 128, 129, 32768, 100_000, 5):
        h.Write(msg[pos:pos + step])
        pos += step
    h.Write(msg[pos:])
    assert h.Sum() == hashlib.sha512(msg).digest()
    assert helpers.Sha512Stream().Sum() == hashlib.sha512(b"").digest()


def test_hasher_streams_like_hash_hash(gpu, tmp_path):
    """snapgpu_hasher_*: Write in arbitrary pieces across many 512 KiB buffers, Sum in mid-stream without
    disturbing the state, and the archive-sha512 use: hashing data.tar.gz WHILE it is written
    (io.MultiWriter shape; clickdeb/deb.go:360-366 + snappy/build.go:222) equals hashing the finished file."""
    import tarfile
    from snappy_b200 import helpers
    rng = np.random.default_rng(15)
    msg = rng.integers(0, 256, 13 * (1 << 20) + 12345, dtype=np.uint8).tobytes()
    h = helpers.Sha512Stream()
    assert h.Sum() == hashlib.sha512(b"").digest()
    pos = 0
    for step in (0, 1, 127, 128, 512 << 10, (512 << 10) - 256, 3, (512 << 10) - 3 + 128, 4 << 20, 5_000_000):
        h.Write(msg[pos:pos + step])
        pos += step
        assert h.Sum() == hashlib.sha512(msg[:pos]).digest()
        assert h.Sum() == hashlib.sha512(msg[:pos]).digest()          # Sum twice: state untouched
    h.Write(msg[pos:])
    assert h.Sum() == hashlib.sha512(msg).digest()

    class Tee:
        def __init__(self, f, hasher):
            self.f, self.h = f, hasher

        def write(self, b):
            self.h.Write(bytes(b))
            return self.f.write(b)

        def flush(self):
            self.f.flush()

    src = tmp_path / "tree"
    src.mkdir()
    for i in range(30):
        (src / f"f{i}").write_bytes(rng.integers(0, 256, int(rng.integers(1, 400_000)), dtype=np.uint8).tobytes())
    tar_path = tmp_path / "data.tar.gz"
    hasher = helpers.Sha512Stream()
    with open(tar_path, "wb") as raw:
        with tarfile.open(fileobj=Tee(raw, hasher), mode="w|gz") as tar:
            tar.add(src, arcname=".")
    assert hasher.Sum().hex() == helpers.Sha512sum(str(tar_path)) == hashlib.sha512(tar_path.read_bytes()).hexdigest()


def test_device_resident_api(gpu, oracle):
    import torch
    from snappy_b200 import device, synth
    lengths = np.concatenate([synth.lognormal_sizes(3000), [0, 5, 200_000]]).astype(np.uint64)
    off, total = synth.layout(lengths)
    d = torch.zeros(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    dg = device.sha512_batch_device(d, off, lengths)
    torch.cuda.synchronize()
    host = d.cpu().numpy()
    # the device generator agrees with the host generator ...
    for i in (0, 17, len(lengths) - 1, len(lengths) - 2):
        assert host[int(off[i]):int(off[i] + lengths[i])].tobytes() == synth.file_bytes(i, int(lengths[i]))
    # ... and the digests with the oracle
    assert np.array_equal(dg.cpu().numpy(), oracle.sha512_batch(host, off, lengths, 8, True))


def test_long_file_chain(gpu):
    """One long serial chain (64 MiB) next to short files."""
    import torch
    from snappy_b200 import device, synth
    lengths = np.array([64 << 20, 100, 4096], dtype=np.uint64)
    off, total = synth.layout(lengths)
    d = torch.zeros(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    dg = device.sha512_batch_device(d, off, lengths).cpu().numpy()
    assert dg[0].tobytes() == hashlib.sha512(synth.file_bytes(0, 64 << 20)).digest()
    assert dg[1].tobytes() == hashlib.sha512(synth.file_bytes(1, 100)).digest()


def test_chain_bound_small_batch_uses_the_pair_bin(gpu, oracle):
    """A small batch whose launch is bound by the chain of its longest files (64 KiB files among a few
    thousand small ones: the last chunk of a host-buffer call, a tree batch): the files of at least 32 KiB
    that are within 31/64 of the longest go to the lane-pair kernel, whose chain is 2.1x shorter; option
    long_min_blocks 1024 restores round 1's rule (nothing below 128 KiB).  Same digests either way."""
    from snappy_b200 import helpers, synth
    rng = np.random.default_rng(77)
    lengths = synth.lognormal_sizes(100_000)[:3000]
    data, off, ln = pack(lengths, rng)
    want = oracle.sha512_batch(data, off, ln, 8)
    for min_blocks, binned in ((0, True), (1024, False), (300, True)):
        gpu.set_option("long_min_blocks", min_blocks)
        gpu.reset_stats()
        got = helpers.sha512_batch(data, off, ln)
        st = gpu.stats()
        assert np.array_equal(got, want), min_blocks
        assert (st.sha512_long_launches >= 1) == binned, (min_blocks, st.sha512_long_launches)
    gpu.set_option("long_min_blocks", 0)
    # nothing as long as 32 KiB: no bin
    data, off, ln = pack(np.minimum(lengths, 30000), rng)
    gpu.reset_stats()
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 8))
    assert gpu.stats().sha512_long_launches == 0


@pytest.mark.parametrize("mode,pair_form,per_cta", [(1, 0, 0), (2, 0, 0), (2, 0, 16), (2, 0, 5), (2, 0, 2), (2, 1, 16)])
def test_long_file_kernel(gpu, oracle, mode, pair_form, per_cta):
    """The long-file bin: files whose chain would dominate a launch leave the batched kernel when a
    launch has at most 256 of them -- mode 1 one lane per file (sha512_long.cuh), mode 2 a lane pair
    per file (sha512_pair.cuh), its lanes exchanging through shared-memory mailboxes (pair_form 0, the
    default) or by warp shuffle (pair_form 1).  The pair form gives a chain its own CTA while a quarter of the
    SMs last (per_cta 0: then the rounds of a block are one branch-free region), two files per CTA up to
    half the SMs' worth, 16 beyond (two regions); per_cta forces the CTA shape the comments below name.  Same digests with the bin
    switched off, and all equal the oracle."""
    from snappy_b200 import helpers
    gpu.set_option("long_kernel", mode)
    gpu.set_option("pair_form", pair_form)
    gpu.set_option("pair_files_per_cta", per_cta)
    rng = np.random.default_rng(21)
    MiB = 1 << 20
    cases = [
        [3 * MiB + 5],                                                   # one file, one producer lane per block
        [2 * MiB - 17, 2 * MiB, 2 * MiB + 111, 5 * MiB + 112],           # padding edges around the threshold
        list(rng.integers(0, 9000, 300)) + [4 * MiB + 1, 2 * MiB + 128, 0, 7 * MiB],
        [2 * MiB + 64 * i + (i % 3) for i in range(33)],                 # 33 long files: two CTAs, 32 + 1
        [2 * MiB + 1024 * i for i in range(5)] + list(rng.integers(1, 70000, 64)),   # 5 files: 32 % 5 idle lanes
        [2 * MiB + 128 * i + (i % 5) for i in range(17)],                # 17 long files: 16 + 1 lane pairs
        [2 * MiB + 777 * i for i in range(16)] + [1, 2, 3],              # exactly 16
        [MiB + 4096 * i for i in range(11)],                             # 11 files: ring of 23 steps, 2 per round
    ]
    for lengths in cases:
        data, off, ln = pack(lengths, rng)
        want = oracle.sha512_batch(data, off, ln, 8)
        gpu.reset_stats()
        gpu.set_option("long_kernel", mode)
        got = helpers.sha512_batch(data, off, ln)
        launches_with = gpu.stats().kernel_launches
        gpu.set_option("long_kernel", 0)
        plain = helpers.sha512_batch(data, off, ln)
        gpu.set_option("long_kernel", mode)
        assert np.array_equal(got, want), lengths[-4:]
        assert np.array_equal(plain, want)
        assert launches_with >= 1
    # long files at odd offsets: the producer realigns, the consumer does not care
    data, off, ln = pack([3 * MiB + 5, 100, 2 * MiB + 1, 17], rng, align=1, jitter=True)
    assert (off % 16 != 0).any()
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 4))
    # 260 long files: more than the one-lane form takes (all stay in the batched kernel), 17 CTAs of the pair form
    lengths = [2 * MiB] * 260
    data, off, ln = pack(lengths, rng)
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 16))
    # more long files than one CTA per SM of the pair form takes (148 x 16): the batched kernel keeps them all
    lengths = [130 * 1024 + 128 * (i % 9) for i in range(2400)]
    data, off, ln = pack(lengths, rng)
    gpu.reset_stats()
    got = helpers.sha512_batch(data, off, ln)
    assert np.array_equal(got, oracle.sha512_batch(data, off, ln, 16))
    # ... and just below that limit they all go to the bin (mode 2)
    data, off, ln = pack(lengths[:2300], rng)
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 16))
    # a long message streamed in pieces (continuation segments go through the long kernel too)
    msg = rng.integers(0, 256, 9 * MiB + 77, dtype=np.uint8).tobytes()
    h = helpers.Sha512Stream()
    h.Write(msg[: 4 * MiB])
    h.Write(msg[4 * MiB: 4 * MiB + 100])
    h.Write(msg[4 * MiB + 100:])
    assert h.Sum() == hashlib.sha512(msg).digest()
    # a file larger than the host staging buffer: pieces of 4 MiB, each a long segment
    gpu.set_option("staging_bytes", 4 * MiB)
    data, off, ln = pack([11 * MiB + 3, 100, 3 * MiB], rng)
    assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 4))


# ---- hashes.yaml ---------------------------------------------------------------------------------

def test_golden_hashes_yaml(gpu, golden_dir, tmp_path):
    """TestBuildCreateDebianHashesSimple (snappy/hashes_test.go:57-104), byte for byte."""
    from snappy_b200 import build
    tree = tmp_path / "tree"
    tree.mkdir()
    make_reference_tree(tree)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(b"")
    build.writeHashes(str(tree), str(tar))
    assert (tree / "DEBIAN" / "hashes.yaml").read_bytes() == (golden_dir / "hashes_simple.yaml").read_bytes()
    assert oct(os.stat(tree / "DEBIAN" / "hashes.yaml").st_mode & 0o777) == "0o644"


def test_config1_tree_yaml_matches_oracle(gpu, oracle, tmp_path):
    """Config 1 as a real tree: 1,000 x 4 KiB in d%04d/f%07d.bin plus a tarball stand-in."""
    from snappy_b200 import build, synth
    tree = tmp_path / "snap"
    names = synth.tree_names(1000)
    for i, n in enumerate(names):
        p = tree / n
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(synth.file_bytes(i, 4096))
    (tree / "meta").mkdir()
    (tree / "meta" / "empty").write_bytes(b"")
    os.symlink("d0000/f0000000.bin", tree / "current")
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(os.urandom(300_000))
    got = build.hashes_yaml(str(tree), str(tar))
    (tree / "DEBIAN" / "hashes.yaml").unlink(missing_ok=True)
    want = oracle.write_hashes(str(tree), str(tar))
    assert got == want
    assert got.count(b"- name: ") == 1000 + 1 + 2 + 1


def test_write_hashes_errors(gpu, tmp_path):
    from snappy_b200 import build
    tree = tmp_path / "t"
    tree.mkdir()
    (tree / "f").write_bytes(b"x")
    with pytest.raises(OSError):
        build.writeHashes(str(tree), str(tmp_path / "no-such-tar"))
    tar = tmp_path / "d"
    tar.write_bytes(b"")
    os.mkfifo(tree / "pipe")
    with pytest.raises(build.UnknownFileMode):
        build.writeHashes(str(tree), str(tar))
    assert not (tree / "DEBIAN" / "hashes.yaml").exists()


def test_tree_larger_than_host_staging(gpu, oracle, tmp_path):
    """Files and a tarball that do not fit one staging batch (streamed + multi-batch)."""
    from snappy_b200 import build
    gpu.set_option("staging_bytes", 1 << 20)
    tree = tmp_path / "t"
    tree.mkdir()
    rng = np.random.default_rng(15)
    for i in range(40):
        (tree / f"f{i:03d}").write_bytes(rng.integers(0, 256, int(rng.integers(0, 200_000)), dtype=np.uint8).tobytes())
    (tree / "big").write_bytes(rng.integers(0, 256, 2_700_003, dtype=np.uint8).tobytes())
    tar = tmp_path / "d.tar.gz"
    tar.write_bytes(rng.integers(0, 256, 3_333_333, dtype=np.uint8).tobytes())
    got = build.hashes_yaml(str(tree), str(tar))
    assert got == oracle.write_hashes(str(tree), str(tar))


def test_archive_rides_along_in_slices(gpu, tmp_path):
    """writeHashes hashes data.tar.gz (archive-sha512, snappy/build.go:222) as a chain of 2 MiB pieces that
    starts before the walk and runs beside the tree's batches (round 1 carried it in slices inside them,
    hence the name): every archive size around the piece boundaries gives hashlib's digest, for a tree of
    several batches, of one batch, and with no regular file at all."""
    from snappy_b200 import build
    rng = np.random.default_rng(33)
    gpu.set_option("staging_bytes", 1 << 20)            # small staging for the chain's calls
    big = tmp_path / "big"
    big.mkdir()
    for i in range(60):
        (big / f"f{i:03d}").write_bytes(rng.integers(0, 256, int(rng.integers(20_000, 90_000)), dtype=np.uint8).tobytes())
    small = tmp_path / "small"
    small.mkdir()
    (small / "one").write_bytes(b"1")
    empty = tmp_path / "empty"
    (empty / "sub").mkdir(parents=True)
    K = 128 << 10
    P = 2 << 20                                         # the chain streamer's piece
    sizes = [0, 1, 127, 128, K - 1, K, K + 1, 3 * K + 77, (1 << 20) + 5, P - 1, P, P + 1, P + 128, 2 * P, 2 * P + 129, 3_000_001]
    for tree in (big, small, empty):
        for n in sizes:
            tar = tmp_path / "data.tar.gz"
            payload = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
            tar.write_bytes(payload)
            doc = build.hashes_yaml(str(tree), str(tar))
            first = doc.split(b"\n", 1)[0]
            assert first == b"archive-sha512: " + hashlib.sha512(payload).hexdigest().encode(), (tree.name, n)
            if tree is big:
                assert doc.count(b"- name: ") == 60
    # a missing archive is the first error, before anything is hashed (build.go:222-226)
    with pytest.raises(OSError):
        build.hashes_yaml(str(big), str(tmp_path / "nope.tar.gz"))


# ---- cmp -----------------------------------------------------------------------------------------

def test_cmp_reference_truth_tables(gpu, tmp_path):
    """helpers/cmp_test.go:29-82 through FilesAreEqual."""
    from snappy_b200 import helpers
    foo = tmp_path / "foo"
    with open(foo, "wb") as f:
        for i in range(1100):
            f.flush()
            if i % 37 == 0 or i in (1023, 1024, 1025, 1099):          # sizes around the 16 KiB chunk
                assert helpers.FilesAreEqual(str(foo), str(foo))
            f.write(b"*" * 16)
    bar = tmp_path / "bar"
    empty = tmp_path / "empty"
    empty.write_bytes(b"")
    assert not helpers.FilesAreEqual(str(empty), str(bar))
    assert not helpers.FilesAreEqual(str(bar), str(empty))
    bar.write_bytes(b"x")
    assert not helpers.FilesAreEqual(str(empty), str(bar))
    assert not helpers.FilesAreEqual(str(bar), str(empty))
    assert helpers.FilesAreEqual(str(empty), str(empty))
    for a, b, r in ((b"hello", b"hello", True), (b"hello", b"world", False), (b"hello", b"hell", False)):
        (tmp_path / "a").write_bytes(a)
        (tmp_path / "b").write_bytes(b)
        assert helpers.FilesAreEqual(str(tmp_path / "a"), str(tmp_path / "b")) is r


def test_cmp_batch_vs_oracle(gpu, oracle):
    from snappy_b200 import helpers
    rng = np.random.default_rng(16)
    lengths = np.concatenate([rng.integers(0, 70000, 500), [0, 1, 15, 16, 17, 16383, 16384, 16385, 1 << 20]])
    a, off, ln = pack(lengths, rng)
    b = a.copy()
    flipped = rng.choice(len(ln), 60, replace=False)
    for i in flipped:
        if ln[i]:
            pos = int(off[i]) + int(rng.integers(0, ln[i]))
            b[pos] ^= 1 << int(rng.integers(0, 8))
    # bytes just outside a pair must not matter
    b[int(off[10] + ln[10])] ^= 0xFF
    got = helpers.cmp_batch(a, b, off, ln)
    assert np.array_equal(got, oracle.cmp_batch(a, b, off, ln, 4))
    assert got.sum() == len(ln) - sum(1 for i in flipped if ln[i])
    # last-byte and first-byte differences
    for i in (3, 4):
        c = a.copy()
        c[int(off[i]) + (int(ln[i]) - 1 if i == 3 else 0)] ^= 0x80
        assert helpers.cmp_batch(a, c, off, ln)[i] == 0


def test_cmp_unaligned_and_chunked(gpu, oracle):
    from snappy_b200 import helpers
    gpu.set_option("staging_bytes", 1 << 20)
    rng = np.random.default_rng(17)
    lengths = np.concatenate([rng.integers(0, 30000, 200), [1_300_000, 700_000]])
    a, off, ln = pack(lengths, rng, align=1, jitter=True)
    b = a.copy()
    b[int(off[-2]) + 1_200_000] ^= 1
    b[int(off[5]) + 2] ^= 1
    assert np.array_equal(helpers.cmp_batch(a, b, off, ln), oracle.cmp_batch(a, b, off, ln, 4))


def test_cmp_device_config4_slice(gpu, oracle):
    """Config 4 at reduced count: 1 MiB pairs, 1 % differing by one byte."""
    import torch
    from snappy_b200 import device, synth
    n = 200
    lengths = np.full(n, 1 << 20, dtype=np.uint64)
    off, total = synth.layout(lengths)
    da = torch.zeros(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(da, off, lengths)
    db = da.clone()
    rng = np.random.default_rng(18)
    differ = sorted(rng.choice(n, 2, replace=False).tolist())
    for i in differ:
        db[int(off[i]) + int(rng.integers(0, 1 << 20))] ^= 0x10
    eq = device.cmp_batch_device(da, db, off, lengths).cpu().numpy()
    assert np.nonzero(eq == 0)[0].tolist() == differ
    assert np.array_equal(eq, oracle.cmp_batch(da.cpu().numpy(), db.cpu().numpy(), off, lengths, 8))


def test_dir_updated_and_apparmor_delta(gpu, oracle, tmp_path):
    """helpers/cmp_test.go:84-133 and policy/policy_test.go:164-182."""
    from snappy_b200 import helpers, policy
    d1, d2 = tmp_path / "d1", tmp_path / "d2"
    d1.mkdir()
    d2.mkdir()
    assert helpers.DirUpdated(str(d1), str(d2), "") == {}
    (d2 / "foo").write_bytes(b"x")
    assert helpers.DirUpdated(str(d1), str(d2), "") == {}
    assert helpers.DirUpdated(str(d2), str(d1), "") == {}
    (d1 / "foo").write_bytes(b"x")
    assert helpers.DirUpdated(str(d1), str(d2), "") == {}
    (d1 / "dir").mkdir()
    assert helpers.DirUpdated(str(d1), str(d2), "") == {}
    (d1 / "foo").write_bytes(b"y")
    (d1 / "bar").write_bytes(b"x")
    (d2 / "bar").write_bytes(b"y")
    (d2 / "baz").write_bytes(b"x")
    assert helpers.DirUpdated(str(d1), str(d2), "") == {"bar": True, "foo": True}
    assert helpers.DirUpdated(str(d1), str(d2), "foo_") == {"foo_bar": True, "foo_foo": True}
    assert helpers.DirUpdated(str(d1), str(d2), "foo_") == oracle.dir_updated(str(d1), str(d2), "foo_")

    orig, dest = tmp_path / "orig", tmp_path / "dest"
    for root, suffix in ((orig, ""), (dest, " 2")):
        base = root / "meta" / "framework-policy" / "apparmor" / "policygroups"
        base.mkdir(parents=True)
        for k in range(3):
            (base / f"policygroups{k}").write_text(f"apparmor::policygroups{k}{suffix}")
    (orig / "meta" / "framework-policy" / "apparmor" / "templates").mkdir()
    (orig / "meta" / "framework-policy" / "apparmor" / "templates" / "t0").write_text("x")
    ps, ts = policy.AppArmorDelta(str(orig), str(dest), "x-")
    assert ps == {"x-policygroups0": True, "x-policygroups1": True, "x-policygroups2": True}
    assert ts == {}
    assert (ps, ts) == oracle.apparmor_delta(str(orig), str(dest), "x-")


def test_native_kernels_ran(gpu):
    """The numbers above came from this library's kernels, not from a fallback."""
    s = gpu.stats()
    assert s.sha512_launches > 0 and s.cmp_launches > 0 and s.kernel_launches >= s.sha512_launches + s.cmp_launches


# ---- copyToBuildDir fused with hashing (snappy/build.go:362-418) -----------------

def test_copy_to_build_dir_fused_with_hashing(gpu, oracle, tmp_path):
    """TestCopyActuallyCopies (snappy/build_test.go:302-312) with the copy path forced: the copied tree
    equals the oracle's, and the writeHashes that follows takes every copied file's digest from the
    cache (one read per file for copy + hash) yet emits the document the oracle emits."""
    from snappy_b200 import build
    from test_host_logic import make_source_tree, snapshot
    rng = np.random.default_rng(33)
    src = tmp_path / "src"
    make_source_tree(src)
    for i in range(40):
        (src / "lib" / f"blob{i:02d}").write_bytes(rng.integers(0, 256, int(rng.integers(0, 50_000)), dtype=np.uint8).tobytes())
    (src / "lib" / "big.bin").write_bytes(rng.integers(0, 256, 3_000_001, dtype=np.uint8).tobytes())
    got, want = tmp_path / "got", tmp_path / "want"
    gpu.lib().snapgpu_digest_cache_clear()
    build.copyToBuildDir(str(src), str(got), no_link=True)
    oracle.copy_to_build_dir(str(src), str(want), no_link=True)
    assert snapshot(got) == snapshot(want)
    assert snapshot(got)["bin/hello-world"][3] is False          # really copied
    nfiles = sum(1 for v in snapshot(got).values() if v[0] == "f")
    entries, hits = build.digest_cache_stats()
    # bin/link (a symlink whose target is copied: it "grows" from the 0 bytes planned for it) did not go
    # to the build dir out of the pinned batch, so its digest is not remembered
    assert entries == nfiles - 1 and hits == 0
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(b"not really a tarball")
    doc = build.hashes_yaml(str(got), str(tar))
    assert doc == oracle.write_hashes(str(want), str(tar))
    assert build.digest_cache_stats() == (0, nfiles - 1)         # the others were not read a second time, and
                                                                 # the cache served this one writeHashes only
    # a file changed after the copy is hashed again (size or mtime no longer match)
    victim = got / "lib" / "blob07"
    body = bytearray(victim.read_bytes() or b"x")
    body[0] ^= 0xFF
    victim.write_bytes(bytes(body))
    os.utime(victim, ns=(1, 1))
    (want / "lib" / "blob07").write_bytes(bytes(body))
    assert build.hashes_yaml(str(got), str(tar)) == oracle.write_hashes(str(want), str(tar))
    # files larger than a packer batch are copied by the streaming path
    gpu.set_option("staging_bytes", 1 << 20)
    got2 = tmp_path / "got2"
    build.copyToBuildDir(str(src), str(got2), no_link=True)
    want2 = tmp_path / "want2"
    oracle.copy_to_build_dir(str(src), str(want2), no_link=True)
    assert snapshot(got2) == snapshot(want2)
    assert build.hashes_yaml(str(got2), str(tar)) == oracle.write_hashes(str(want2), str(tar))
    gpu.lib().snapgpu_digest_cache_clear()


# ---- hashes.yaml verification -------------------------------------------------

def test_verify_hashes(gpu, oracle, tmp_path):
    """Build -> hashes.yaml -> 'install' (copy of the tree with meta/hashes.yaml, snappy/click.go:330-338)
    -> verify.  The report equals the oracle's (PyYAML parse + CPU re-hash) for every kind of drift."""
    import shutil
    from snappy_b200 import build
    from test_host_logic import make_source_tree
    rng = np.random.default_rng(44)
    src = tmp_path / "build"
    make_source_tree(src)
    for i in range(25):
        (src / "lib" / f"blob{i:02d}").write_bytes(rng.integers(0, 256, int(rng.integers(1, 30_000)), dtype=np.uint8).tobytes())
    (src / "lib" / "with space and a very long name that yaml folds at eighty columns because it has spaces in it.txt").write_text("x")
    (src / "lib" / "true").write_text("quoted name")
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(b"archive bytes")
    build.writeHashes(str(src), str(tar))
    inst = tmp_path / "inst"
    shutil.copytree(src, inst, symlinks=True, ignore=shutil.ignore_patterns("DEBIAN"))
    shutil.copy(src / "DEBIAN" / "hashes.yaml", inst / "meta" / "hashes.yaml")
    y = str(inst / "meta" / "hashes.yaml")
    assert build.verifyHashes(str(inst), y) == [] == oracle.verify_hashes(str(inst), y)
    assert build.verifyHashes(str(inst), y, str(tar)) == []
    assert not (inst / "DEBIAN").exists()                            # verification does not touch the tree
    other = tmp_path / "other.tar.gz"
    other.write_bytes(b"another archive")
    assert build.verifyHashes(str(inst), y, str(other)) == ["archive-sha512 differs"] == oracle.verify_hashes(str(inst), y, str(other))
    # drift of every kind
    b = bytearray((inst / "lib" / "blob03").read_bytes())
    b[-1] ^= 1
    (inst / "lib" / "blob03").write_bytes(bytes(b))                  # same size, other content
    (inst / "lib" / "blob04").write_bytes(b"shorter")                # size and content
    os.chmod(inst / "bin" / "hello-world", 0o700)                    # mode only
    os.remove(inst / "lib" / "blob05")                               # missing
    (inst / "lib" / "new-file").write_text("added")                  # extra
    os.remove(inst / "bin" / "link")
    (inst / "bin" / "link").write_text("was a symlink")              # type change: size/sha512 appear, mode differs
    got = build.verifyHashes(str(inst), y)
    assert got == oracle.verify_hashes(str(inst), y)
    assert got == ["changed: bin/hello-world (mode)", "changed: bin/link (size sha512 mode)", "changed: lib/blob03 (sha512)",
                   "changed: lib/blob04 (size sha512)", "missing: lib/blob05", "extra: lib/new-file"]
    with pytest.raises(OSError):
        build.verifyHashes(str(inst), str(tmp_path / "no-such.yaml"))


# ---- BASELINE.json configs at full size: size-independent properties --------------------------------

def _spot_check(dg, lengths, picks, first_index=0):
    from snappy_b200 import synth
    for i in picks:
        want = hashlib.sha512(synth.file_bytes(first_index + int(i), int(lengths[i]))).digest()
        assert dg[i].tobytes() == want, f"file {i} differs from hashlib"


def test_config2_full_size_properties(gpu):
    """Config 2 as benchmarked (100,000 files, log-normal 1-64 KiB, 1.29 GB, generated in HBM): spot checks
    against hashlib, order independence (checksum of checksums), launch-shape independence, and the
    host-buffer pipeline on the same bytes."""
    import torch
    from snappy_b200 import device, helpers, synth
    lengths = synth.lognormal_sizes(100_000)
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    dg = device.sha512_batch_device(d, off, lengths).cpu().numpy()
    rng = np.random.default_rng(2)
    longest = np.argsort(lengths)[-8:]
    _spot_check(dg, lengths, np.concatenate([rng.integers(0, len(lengths), 48), longest, [0, len(lengths) - 1]]))
    perm = rng.permutation(len(lengths))
    dg_perm = device.sha512_batch_device(d, off[perm], lengths[perm]).cpu().numpy()
    assert hashlib.sha512(dg_perm[np.argsort(perm)].tobytes()).digest() == hashlib.sha512(dg.tobytes()).digest()
    for warps in (2, 3):
        gpu.set_option("sha_warps_per_sm", warps)
        assert np.array_equal(device.sha512_batch_device(d, off, lengths).cpu().numpy(), dg)
    gpu.set_option("sha_warps_per_sm", 0)
    host = d.cpu().numpy()
    assert np.array_equal(helpers.sha512_batch(host, off, lengths), dg)


def test_config5_shard_full_size_properties(gpu):
    """One GPU's shard of config 5 (250,000 files x 64 KiB = 16.4 GB in HBM)."""
    import torch
    from snappy_b200 import device, synth
    n = 250_000
    lengths = np.full(n, 65536, dtype=np.uint64)
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths, first_index=3 * n)          # rank 3's files
    dg = device.sha512_batch_device(d, off, lengths).cpu().numpy()
    rng = np.random.default_rng(3)
    _spot_check(dg, lengths, np.concatenate([rng.integers(0, n, 24), [0, n - 1]]), first_index=3 * n)
    assert len(np.unique(dg.view(np.dtype((np.void, 64))))) == n        # distinct contents, distinct digests
    gpu.set_option("sha_warps_per_sm", 1)
    assert np.array_equal(device.sha512_batch_device(d, off, lengths).cpu().numpy(), dg)
    gpu.set_option("sha_warps_per_sm", 0)


def test_config4_full_size_properties(gpu):
    """Config 4 as benchmarked: 10,000 pairs x 1 MiB, 100 of them differing in one byte."""
    import torch
    from snappy_b200 import device, synth
    n = 10_000
    lengths = np.full(n, 1 << 20, dtype=np.uint64)
    off, total = synth.layout(lengths)
    da = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(da, off, lengths)
    db = da.clone()
    rng = np.random.default_rng(synth.SEED)
    differ = np.sort(rng.choice(n, 100, replace=False))
    pos = rng.integers(0, 1 << 20, 100)
    pos[:3] = [0, (1 << 20) - 1, 16384]                                    # first byte, last byte, a tile edge
    db[torch.from_numpy(off[differ].astype(np.int64) + pos).to("cuda:0")] ^= 0x80
    eq_ab = device.cmp_batch_device(da, db, off, lengths).cpu().numpy()
    eq_ba = device.cmp_batch_device(db, da, off, lengths).cpu().numpy()
    assert np.nonzero(eq_ab == 0)[0].tolist() == differ.tolist()
    assert np.array_equal(eq_ab, eq_ba)                                    # symmetric
    assert device.cmp_batch_device(da, da, off, lengths).cpu().numpy().all()   # reflexive
    del db


def test_config3_shape_at_reduced_size(gpu):
    """Config 3's shape (50,000 small files + 4 long ones) with 48 MiB instead of 1 GiB long files; the
    full-size run is tools/cfg3_tail.py (profiles/), which checks the 1 GiB digests against hashlib too."""
    import torch
    from snappy_b200 import device, synth
    lengths = np.concatenate([synth.lognormal_sizes(100_000)[:50_000], np.full(4, 48 << 20, dtype=np.uint64)])
    off, total = synth.layout(lengths)
    d = torch.empty(total, dtype=torch.uint8, device="cuda:0")
    device.synth_fill_device(d, off, lengths)
    gpu.reset_stats()
    dg = device.sha512_batch_device(d, off, lengths).cpu().numpy()
    # the four long files leave the batched kernel although the launch has thousands of files above the bin's lower
    # bound (32 KiB): with too many candidates of that size the bin falls back to the files of 128 KiB and more
    assert gpu.stats().sha512_long_launches == 1
    _spot_check(dg, lengths, [50_000, 50_001, 50_002, 50_003, 0, 49_999, 123, 31_337])
    gpu.set_option("long_kernel", 0)
    small = device.sha512_batch_device(d, off[:50_000], lengths[:50_000]).cpu().numpy()
    gpu.set_option("long_kernel", 2)
    assert np.array_equal(small, dg[:50_000])


# ---- the boundary under concurrency and with several devices -------------------------------------------

def test_concurrent_callers(gpu, oracle, tmp_path):
    """cgo calls arrive on arbitrary OS threads: eight threads hash, compare and write hashes.yaml at the
    same time (ctypes releases the GIL around every call); every result equals the oracle's."""
    import threading
    from snappy_b200 import build, helpers
    rng = np.random.default_rng(55)
    jobs = []
    for t in range(8):
        lengths = np.concatenate([rng.integers(0, 40_000, 400 + 50 * t), [0, 111, 112, 3_000_000 if t % 3 == 0 else 5]])
        data, off, ln = pack(lengths, rng)
        b = data.copy()
        b[int(off[7]) + int(ln[7]) // 2] ^= 4
        jobs.append((data, b, off, ln))
    tree = tmp_path / "tree"
    tree.mkdir()
    make_reference_tree(tree)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(b"")
    want_yaml = oracle.write_hashes(str(tree), str(tar))
    results, errors = [None] * 8, []

    def work(t):
        try:
            data, b, off, ln = jobs[t]
            for _ in range(3):
                dg = helpers.sha512_batch(data, off, ln)
                eq = helpers.cmp_batch(data, b, off, ln)
                doc = build.hashes_yaml(str(tree), str(tar))
            results[t] = (dg, eq, doc)
        except Exception as e:                        # noqa: BLE001 -- reported below
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t, (data, b, off, ln) in enumerate(jobs):
        dg, eq, doc = results[t]
        assert np.array_equal(dg, oracle.sha512_batch(data, off, ln, 4))
        assert np.array_equal(eq, oracle.cmp_batch(data, b, off, ln, 4))
        assert doc == want_yaml


def test_in_process_sharding_over_all_devices(native, oracle):
    """snapgpu_init over every visible GPU: the file list is sharded across them inside one call (no
    collective; results gathered by index).  With one GPU this is the single-device path again."""
    import torch
    from snappy_b200 import helpers, synth
    ndev = torch.cuda.device_count()
    native.init(list(range(ndev)))
    try:
        lengths = np.concatenate([synth.lognormal_sizes(20_000), [5_000_000, 0, 3]]).astype(np.uint64)
        data, off, ln = synth.make_host_batch(lengths)
        assert np.array_equal(helpers.sha512_batch(data, off, ln), oracle.sha512_batch(data, off, ln, 8))
        b = data.copy()
        flips = [0, 9_999, 19_999, 20_000]
        for i in flips:
            b[int(off[i])] ^= 1
        assert np.nonzero(helpers.cmp_batch(data, b, off, ln) == 0)[0].tolist() == flips
        assert native.lib().snapgpu_num_devices() == ndev
    finally:
        native.init([int(os.environ.get("LOCAL_RANK", "0"))])


def test_cmp_large_pageable_buffers(gpu, oracle):
    """Pairs of 1-6 MiB in ordinary (pageable) numpy memory: the spans go through the pinned bounce buffers
    (h2d_span) and several staging chunks; differences at the first byte, the last byte and a tile edge."""
    from snappy_b200 import helpers
    gpu.set_option("staging_bytes", 16 << 20)
    rng = np.random.default_rng(61)
    lengths = [6 << 20, (1 << 20) + 5, 3 << 20, 4096, 0, (2 << 20) - 1, 5 << 20]
    a, off, ln = pack(lengths, rng)
    b = a.copy()
    b[int(off[0])] ^= 1                                  # first byte of a pair
    b[int(off[2]) + (3 << 20) - 1] ^= 1                  # last byte
    b[int(off[6]) + 16384 * 77] ^= 1                     # first byte of a 16 KiB tile
    got = helpers.cmp_batch(a, b, off, ln)
    assert got.tolist() == [0, 1, 0, 1, 1, 1, 0]
    assert np.array_equal(got, oracle.cmp_batch(a, b, off, ln, 4))
    assert np.array_equal(helpers.sha512_batch(a, off, ln), oracle.sha512_batch(a, off, ln, 4))


# ---- a C caller of the ABI (idiom of helpers/touch.go:20-57) ---------------

def build_c_harness(tmp_path):
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "c_abi_harness"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-o", str(exe), str(ROOT / "tests" / "c_abi_harness.c"),
                           "-L" + str(ROOT / "snappy_b200"), "-lsnapgpu", "-Wl,-rpath," + str(ROOT / "snappy_b200")])
    return exe


def test_c_program_writes_the_golden_hashes_yaml(gpu, golden_dir, tmp_path):
    """tests/c_abi_harness.c, linked against libsnapgpu.so like a cgo shim would be: it builds the tree
    of snappy/hashes_test.go:57-87 with libc, calls snapgpu_write_hashes and diffs DEBIAN/hashes.yaml with
    the reference's golden document; then Sha512sum / FilesAreEqual known answers and a pinned batch."""
    import subprocess
    exe = build_c_harness(tmp_path)
    p = subprocess.run([str(exe), str(tmp_path / "work"), str(golden_dir / "hashes_simple.yaml")], capture_output=True, text=True,
                       timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == "ok", (p.returncode, p.stdout, p.stderr)


def test_mixed_tree_every_size_class(gpu, oracle, tmp_path):
    """The tree of tests/test_hostsim.py (small / mid / chain size classes, wide and nested directories,
    symlinks, names yaml.v2 quotes, DEBIAN-prefixed names) through the real kernels, with a multi-piece
    archive whose chain runs beside the tree's batches; twice, with different packer fan-out."""
    from snappy_b200 import build
    from test_hostsim import make_mixed_tree
    rng = np.random.default_rng(7)
    tree = tmp_path / "tree"
    make_mixed_tree(tree, rng)
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(rng.integers(0, 256, size=(5 << 20) + 77, dtype=np.uint8).tobytes())
    want = oracle.write_hashes(str(tree), str(tar))
    for _ in range(2):
        assert build.hashes_yaml(str(tree), str(tar)) == want
    st = gpu.tree_stats()
    assert st["files_hashed"] > 1100 and st["batches"] >= 1 and st["yaml_bytes"] == len(want)


def test_concurrent_callers_overlap(gpu, oracle):
    """Two callers on one device use different pipes: a chain-bound call (one 3 MiB message: ~45 ms of
    serial chain, almost nothing to copy) and a copy-bound call (512 MiB of 64 KiB files: ~10 ms of PCIe)
    together take less than the two one after the other."""
    import threading
    import time
    from snappy_b200 import helpers
    rng = np.random.default_rng(77)
    chain = rng.integers(0, 256, size=3 << 20, dtype=np.uint8)
    c_off, c_len = np.array([0], np.uint64), np.array([len(chain)], np.uint64)
    n = 8192
    bulk = rng.integers(0, 256, size=n * 65536, dtype=np.uint8)
    b_off, b_len = np.arange(n, dtype=np.uint64) * np.uint64(65536), np.full(n, 65536, np.uint64)
    want_chain = oracle.sha512_batch(chain, c_off, c_len, 1)
    helpers.sha512_batch(chain, c_off, c_len)
    helpers.sha512_batch(bulk, b_off, b_len)                      # warm-up: staging buffers of both shapes
    best_serial, best_both = 1e9, 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        helpers.sha512_batch(chain, c_off, c_len)
        helpers.sha512_batch(bulk, b_off, b_len)
        best_serial = min(best_serial, time.perf_counter() - t0)
        out = {}
        th = [threading.Thread(target=lambda: out.__setitem__("c", helpers.sha512_batch(chain, c_off, c_len))),
              threading.Thread(target=lambda: out.__setitem__("b", helpers.sha512_batch(bulk, b_off, b_len)))]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        best_both = min(best_both, time.perf_counter() - t0)
        assert np.array_equal(out["c"], want_chain)
    spot = [0, 1, n // 2, n - 1]
    assert np.array_equal(out["b"][spot], oracle.sha512_batch(bulk, b_off[spot], b_len[spot], 1))
    assert best_both < 0.9 * best_serial, (best_both, best_serial)


def test_archive_hashed_while_written_then_write_hashes_with_the_digest(gpu, oracle, tmp_path):
    """The flow of INTEGRATION.md section 3b: data.tar.gz goes through an io.MultiWriter(file, hasher)
    while tar | gzip produces it (clickdeb/deb.go:261-285), and writeHashes gets the digest instead of
    the file name (snapgpu_hashes_yaml_digest): the document equals the one made from the finished file."""
    import subprocess
    from snappy_b200 import build, helpers
    from test_hostsim import make_mixed_tree
    rng = np.random.default_rng(17)
    tree = tmp_path / "tree"
    make_mixed_tree(tree, rng, nfiles=150)
    tar = tmp_path / "data.tar.gz"
    h = helpers.Sha512Stream()
    p = subprocess.Popen(["tar", "-C", str(tree), "-cz", "."], stdout=subprocess.PIPE)
    with open(tar, "wb") as f:
        while True:
            piece = p.stdout.read(1 << 16)
            if not piece:
                break
            f.write(piece)
            h.Write(piece)
    assert p.wait() == 0
    digest = h.Sum()
    assert digest.hex() == oracle.sha512sum(str(tar))
    want = oracle.write_hashes(str(tree), str(tar))
    assert build.hashes_yaml_digest(str(tree), digest) == want
    assert build.hashes_yaml(str(tree), str(tar)) == want


def test_warm_up_then_write_hashes(gpu, oracle, tmp_path):
    """snapgpu_warm (no counterpart in the reference: it only moves one-off allocation and kernel-load cost ahead
    of the first writeHashes) leaves results alone: the document of a tree hashed right after it is the oracle's,
    and it can be called again at any time."""
    from snappy_b200 import build
    build.warm()
    tree = tmp_path / "t"
    (tree / "bin").mkdir(parents=True)
    rng = np.random.default_rng(5)
    for i, n in enumerate([0, 1, 111, 112, 4096, 40_000, 70_000, 300_000]):
        (tree / "bin" / f"f{i}").write_bytes(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
    tar = tmp_path / "data.tar.gz"
    tar.write_bytes(rng.integers(0, 256, 50_000, dtype=np.uint8).tobytes())
    assert build.hashes_yaml(str(tree), str(tar)) == oracle.write_hashes(str(tree), str(tar))
    build.warm()
    assert build.hashes_yaml(str(tree), str(tar)) == oracle.write_hashes(str(tree), str(tar))
    # ... and from a second thread while writeHashes is already running (a build's goroutine may lose the race)
    import threading
    errors = []

    def warm():
        try:
            build.warm()
        except Exception as exc:                              # noqa: BLE001
            errors.append(exc)

    t = threading.Thread(target=warm)
    t.start()
    doc = build.hashes_yaml(str(tree), str(tar))
    t.join()
    assert not errors and doc == oracle.write_hashes(str(tree), str(tar))


def test_two_ended_claims_cover_every_unit_once(gpu, oracle):
    """Launches with several CTAs per SM claim units from both ends of the length-sorted plan (slow warps
    from the short end, sha512_kernels.cuh): on 400,000 files -- enough units for the mode to switch on at
    two and at three CTAs per SM -- every digest equals the oracle's, with the mode on and off."""
    import torch
    from snappy_b200 import device
    rng = np.random.default_rng(99)
    lengths = np.concatenate([rng.integers(0, 3000, 399_000), rng.integers(20_000, 70_000, 1000)]).astype(np.uint64)
    rng.shuffle(lengths)
    data, off, ln = pack(lengths, rng)
    want = oracle.sha512_batch(data, off, ln, 16, True)
    d = torch.from_numpy(data).cuda()
    for warps in (2, 3):
        gpu.set_option("sha_warps_per_sm", warps)
        for mode in (2, 1, 0):                    # always / when the lengths are spread wide (they are here) / never
            gpu.set_option("two_ended", mode)
            got = device.sha512_batch_device(d, off, ln).cpu().numpy()
            assert np.array_equal(got, want), (warps, mode)
    gpu.set_option("sha_warps_per_sm", 0)
    gpu.set_option("two_ended", 1)
