"""The C-ABI library loads, exports what include/snapgpu.h declares, and refuses to compute
without a GPU (no CPU fallback)."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "snapgpu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snapgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    syms = declared_symbols()
    assert len(syms) >= 25
    L = native.lib()
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/snapgpu.h but not exported"
    # and the binding table covers the header exactly
    assert sorted(native.SIGNATURES) == syms


def test_exports_are_plain_c(native):
    out = subprocess.run(["nm", "-D", "--defined-only", str(native.LIB_PATH)], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [s for s in exported if s.startswith("snapgpu_")]
    assert set(declared_symbols()) <= set(ours)


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "snapgpu.h"\nint (*fp)(void) = snapgpu_num_devices;\nint main(void){ return fp == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o",
                           str(tmp_path / "t.o")])


def test_no_cpu_fallback(native):
    """Without an initialised CUDA device every compute entry point fails loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU box")
    L = native.lib()
    assert L.snapgpu_init(None, 0) == native.ECUDA
    assert "no CPU fallback" in native.last_error()
    data = np.zeros(64, dtype=np.uint8)
    off = np.zeros(1, dtype=np.uint64)
    ln = np.full(1, 3, dtype=np.uint64)
    out = np.zeros(64, dtype=np.uint8)
    assert L.snapgpu_sha512_batch(data.ctypes.data, off.ctypes.data, ln.ctypes.data, 1, out.ctypes.data) == native.ENOINIT
    assert L.snapgpu_cmp_batch(data.ctypes.data, data.ctypes.data, off.ctypes.data, ln.ctypes.data, 1,
                               out.ctypes.data) == native.ENOINIT
    assert not out.any()
    hexbuf = ctypes.create_string_buffer(129)
    assert L.snapgpu_sha512sum_file(b"/etc/hostname", hexbuf) < 0
    assert hexbuf.value == b""
    # FilesAreEqual maps every failure to false, a dead GPU included
    assert L.snapgpu_files_are_equal(b"/etc/hostname", b"/etc/hostname") == 0


def test_python_mirror_raises_without_gpu(native, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from snappy_b200 import build, helpers
    p = tmp_path / "f"
    p.write_bytes(b"x")
    with pytest.raises(native.SnapGpuError):
        helpers.Sha512sum(str(p))
    with pytest.raises(native.SnapGpuError):
        build.writeHashes(str(tmp_path), str(p))
    with pytest.raises(native.SnapGpuError):
        helpers.sha512_batch(np.zeros(16, np.uint8), [0], [1])


def test_product_does_not_touch_the_oracle():
    """Nothing under snappy_b200/ may import, link or execute oracle/ (task section 3)."""
    for p in (ROOT / "snappy_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h") or p.name == "Makefile":
            text = p.read_text(errors="replace")
            assert "oracle" not in text.lower(), p
    assert "hashlib" not in "".join(p.read_text() for p in (ROOT / "snappy_b200").glob("*.py"))


def test_c_caller_links_and_the_library_refuses_without_cuda(native, tmp_path):
    """tests/c_abi_harness.c compiles as C11 with -Werror against include/snapgpu.h, links against
    libsnapgpu.so and runs; in this container (no GPU) its --no-gpu mode checks that init and the
    compute entry points fail instead of falling back to a CPU."""
    import subprocess
    from test_gpu_parity import build_c_harness
    exe = build_c_harness(tmp_path)
    p = subprocess.run([str(exe), "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.startswith("ok") or "nothing to check" in p.stdout, (p.stdout, p.stderr)
