import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def native():
    """libsnapgpu through ctypes; built on demand (nvcc cross-compiles without a GPU)."""
    from snappy_b200 import _native as N
    if not N.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    N.lib()
    return N


@pytest.fixture(scope="session")
def gpu(native):
    """Initialised library on cuda:0; a GPU test without a device is an error, not a skip."""
    native.init([int(os.environ.get("LOCAL_RANK", "0"))])
    return native


def make_reference_tree(root: Path):
    """The tree of TestBuildCreateDebianHashesSimple (snappy/hashes_test.go:57-87)."""
    (root / "DEBIAN").mkdir(mode=0o755)
    (root / "DEBIAN" / "bar").write_bytes(b"")
    (root / "foo").write_bytes(b"")
    os.chmod(root / "foo", 0o644)
    (root / "bin").mkdir(mode=0o755)
    os.chmod(root / "bin", 0o755)
    (root / "bin" / "bar").write_bytes(b"bar\n")
    os.chmod(root / "bin" / "bar", 0o644)
    os.symlink("/dsafdsafsadf", root / "broken-link")
