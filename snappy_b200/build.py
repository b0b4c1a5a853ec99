"""Mirror of the hashes writer of the reference's ``snappy`` package.

* ``writeHashes``   /root/reference/snappy/build.go:216-270
* ``hashes_yaml``   the same walk, returning the document instead of writing it
* ``copyToBuildDir`` /root/reference/snappy/build.go:362-418 (+ ``shouldExclude``, build.go:52-83)

The tree walk, the batching of every regular file into one GPU call and the yaml.v2-exact
emitter all live in libsnapgpu (csrc/host_path.cpp); this module is the binding.
"""
from __future__ import annotations

import ctypes

from . import _native as N


class UnknownFileMode(Exception):
    """yamlFileMode.MarshalYAML's "Unknown file mode" (snappy/hashes.go:44)."""


def _raise(rc: int):
    if rc == N.EIO:
        raise OSError(N.last_error())
    if rc == N.EMODE:
        raise UnknownFileMode(N.last_error())
    N.check(rc)


def writeHashes(buildDir: str, dataTar: str) -> None:
    """Write ``<buildDir>/DEBIAN/hashes.yaml``; the first error aborts, as in the reference."""
    _raise(N.lib().snapgpu_write_hashes(N.fs(buildDir), N.fs(dataTar)))


def copyToBuildDir(sourceDir: str, buildDir: str, no_link: bool = False) -> None:
    """copyToBuildDir (/root/reference/snappy/build.go:362-418).  Files that have to be copied are
    read once, for the copy and for the SHA-512 that the following writeHashes needs."""
    _raise(N.lib().snapgpu_copy_to_build_dir(N.fs(sourceDir), N.fs(buildDir), 1 if no_link else 0))


def warm() -> None:
    """snapgpu_warm: pin the packer's chunk pool, allocate the pipeline's buffers and launch every kernel once,
    ahead of the first writeHashes of the process (no counterpart in the reference)."""
    _raise(N.lib().snapgpu_warm())


def shouldExclude(basename: str) -> bool:
    """shouldExclude (/root/reference/snappy/build.go:52-83)."""
    return bool(N.lib().snapgpu_should_exclude(N.fs(basename)))


def verifyHashes(root: str, yamlPath: str, dataTar: str | None = None) -> list[str]:
    """Re-hash ``root`` and diff it against the hashes.yaml at ``yamlPath`` (no counterpart in the
    reference).  Returns the report lines; an empty list means it verifies."""
    from .helpers import _names
    ptr, count = ctypes.c_void_p(), ctypes.c_size_t()
    _raise(N.lib().snapgpu_verify_hashes(N.fs(root), N.fs(yamlPath), N.fs(dataTar) if dataTar else None,
                                         ctypes.byref(ptr), ctypes.byref(count)))
    return _names(ptr, count)


def readArchiveSha512(yamlPath: str) -> str:
    """``part.hash`` of NewSnapPartFromYaml (snappy/snapp.go:466-478): archive-sha512 of a hashes.yaml,
    after decoding every entry's mode like yaml.Unmarshal does (UnknownFileMode on a bad one)."""
    buf = ctypes.create_string_buffer(256)
    _raise(N.lib().snapgpu_read_archive_sha512(N.fs(yamlPath), buf, 256))
    return buf.value.decode()


def digest_cache_stats() -> tuple[int, int]:
    entries, hits = ctypes.c_size_t(), ctypes.c_uint64()
    N.lib().snapgpu_digest_cache_stats(ctypes.byref(entries), ctypes.byref(hits))
    return entries.value, hits.value


def hashes_yaml_digest(buildDir: str, archiveSha512: bytes) -> bytes:
    """writeHashes with archive-sha512 already known (streamed through ``helpers.Sha512Stream`` while
    data.tar.gz was being written, clickdeb/deb.go:360-366): the archive is not read again."""
    assert len(archiveSha512) == 64
    buf = ctypes.create_string_buffer(archiveSha512, 64)
    ptr, length = ctypes.c_void_p(), ctypes.c_size_t()
    _raise(N.lib().snapgpu_hashes_yaml_digest(N.fs(buildDir), ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ptr),
                                              ctypes.byref(length)))
    return N.take_string(ptr, length.value)


def hashes_yaml(buildDir: str, dataTar: str) -> bytes:
    ptr, length = ctypes.c_void_p(), ctypes.c_size_t()
    _raise(N.lib().snapgpu_hashes_yaml(N.fs(buildDir), N.fs(dataTar), ctypes.byref(ptr), ctypes.byref(length)))
    return N.take_string(ptr, length.value)
