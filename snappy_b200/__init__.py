"""snappy_b200 -- B200-native hashing and compare path of Ubuntu's ``snappy`` package manager.

Only the one data-parallel hot path of the reference is here:

* ``helpers``  Sha512sum / FilesAreEqual / DirUpdated and their batch forms
* ``build``    writeHashes (DEBIAN/hashes.yaml)
* ``policy``   AppArmorDelta
* ``device``   device-resident batch calls used by the benchmark (kernel-only timing)
* ``synth``    the deterministic synthetic trees of the benchmark configs

Everything computes in hand-written sm_100a CUDA kernels behind the C ABI of
``include/snapgpu.h`` (``libsnapgpu.so``); there is no CPU fallback.
"""
from . import _native  # noqa: F401
from . import build, helpers, policy  # noqa: F401

__all__ = ["helpers", "build", "policy"]
