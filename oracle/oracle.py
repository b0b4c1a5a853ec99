"""CPU ORACLE -- test infrastructure, NOT the product.

Python side of the oracle: a ctypes view of ``liboracle.so`` (the plain-C
restatement in ``sha512_oracle.c``) plus pure-Python restatements of the host
logic on the hot path:

* ``write_hashes``      /root/reference/snappy/build.go:216-270
* ``file_mode_string``  /root/reference/snappy/hashes.go:33-57
* ``marshal_hashes``    yaml.Marshal of hashesYaml/fileHash, snappy/hashes.go:93-110
* ``files_are_equal``   /root/reference/helpers/cmp.go:31-59
* ``streams_equal``     /root/reference/helpers/cmp.go:61-86
* ``dir_updated``       /root/reference/helpers/cmp.go:97-114
* ``apparmor_delta``    /root/reference/policy/policy.go:155-167
* ``copy_to_build_dir`` /root/reference/snappy/build.go:362-418 (``should_exclude``: build.go:52-83)
* ``verify_hashes``     no reference counterpart: fileHash / yamlFileMode semantics
                        of snappy/hashes.go:59-110 applied to a re-hash of the tree

Third-party code that is NOT under /root/reference and is restated here from its
published behaviour:

* Go stdlib ``path/filepath.Walk`` (pre-order, per-directory ``sort.Strings`` of the
  names, ``Lstat``; root visited first) and ``filepath.Glob`` (sorted names, dotfiles
  match ``*``).
* ``gopkg.in/yaml.v2`` @ 49c95bdc21843256fb6c4e0d370a05f24a0bf213 (2015-02-24,
  dependencies.tsv:7): struct encoder + the libyaml-derived emitter (scalar analysis,
  style selection, block sequence/mapping layout, width-80 folding).

Parity pinning: the golden document of snappy/hashes_test.go:89-103 and the fragment of
snappy/hashes_test.go:30-33 are reproduced byte for byte (tests/test_oracle.py).  Names
outside the plain-safe ASCII subset (quoting, folding, !!binary) follow the yaml.v2 rules
restated below but are PARITY UNPINNED: the reference holds no vector for them.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import base64
import ctypes
import fnmatch
import os
import re
import stat
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    """Compile liboracle.so (gcc) if it is missing or stale."""
    so = _HERE / "liboracle.so"
    src = _HERE / "sha512_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(_HERE), "liboracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        u8p = ctypes.c_void_p
        L.oracle_sha512.argtypes = [u8p, ctypes.c_size_t, u8p]
        L.oracle_sha512.restype = None
        L.oracle_sha512sum_file.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.oracle_sha512sum_file.restype = ctypes.c_int
        L.oracle_sha512_batch.argtypes = [u8p, u8p, u8p, ctypes.c_size_t, u8p, ctypes.c_int, ctypes.c_int]
        L.oracle_sha512_batch.restype = ctypes.c_int
        L.oracle_sha512sum_files.argtypes = [ctypes.c_char_p, u8p, ctypes.c_size_t, u8p, ctypes.c_int, ctypes.c_int]
        L.oracle_sha512sum_files.restype = ctypes.c_int
        L.oracle_cmp_batch.argtypes = [u8p, u8p, u8p, u8p, ctypes.c_size_t, u8p, ctypes.c_int]
        L.oracle_cmp_batch.restype = ctypes.c_int
        L.oracle_streams_equal.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.c_size_t]
        L.oracle_streams_equal.restype = ctypes.c_int
        L.oracle_files_are_equal.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.oracle_files_are_equal.restype = ctypes.c_int
        L.oracle_cmp_algorithmic_bytes.argtypes = [u8p, u8p, ctypes.c_uint64]
        L.oracle_cmp_algorithmic_bytes.restype = ctypes.c_uint64
        L.oracle_have_openssl.restype = ctypes.c_int
        L.oracle_now_seconds.restype = ctypes.c_double
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


# --------------------------------------------------------------------------- SHA-512

def sha512(data: bytes) -> bytes:
    out = ctypes.create_string_buffer(64)
    buf = ctypes.create_string_buffer(data, len(data)) if data else ctypes.create_string_buffer(1)
    lib().oracle_sha512(ctypes.cast(buf, ctypes.c_void_p), len(data), ctypes.cast(out, ctypes.c_void_p))
    return out.raw


def sha512sum(path: str) -> str:
    """helpers.Sha512sum: hex digest of a file; raises OSError like Go returns err."""
    out = ctypes.create_string_buffer(129)
    rc = lib().oracle_sha512sum_file(os.fsencode(path), out)
    if rc != 0:
        raise OSError(-rc, os.strerror(-rc), path)
    return out.value.decode("ascii")


def sha512_batch(data: np.ndarray, offsets: np.ndarray, lengths: np.ndarray,
                 nthreads: int = 1, use_openssl: bool = False) -> np.ndarray:
    """Digests (n, 64) uint8 of ``data[offsets[i]:offsets[i]+lengths[i]]``."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
    n = len(offsets)
    out = np.zeros((n, 64), dtype=np.uint8)
    if n:
        rc = lib().oracle_sha512_batch(_ptr(data), _ptr(offsets), _ptr(lengths), n, _ptr(out),
                                       int(nthreads), int(use_openssl))
        if rc != 0:
            raise RuntimeError(f"oracle_sha512_batch failed: {rc}")
    return out


def sha512sum_files(paths, sizes, nthreads: int = 1, use_openssl: bool = False) -> np.ndarray:
    """helpers.Sha512sum over a list of files -- the loop of writeHashes (snappy/build.go:240) --
    on ``nthreads`` threads (1 = what the reference does; more = the file list statically sharded,
    the "best CPU" comparator of SURVEY.md 8d).  Digests (n, 64) uint8; raises OSError."""
    blob = b"".join(os.fsencode(p) + b"\0" for p in paths)
    sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
    out = np.zeros((len(paths), 64), dtype=np.uint8)
    if len(paths):
        rc = lib().oracle_sha512sum_files(blob, _ptr(sizes), len(paths), _ptr(out), int(nthreads), int(use_openssl))
        if rc != 0:
            raise OSError(-rc, os.strerror(-rc))
    return out


# --------------------------------------------------------------------------- cmp

def streams_equal(a: bytes, b: bytes) -> bool:
    ba = ctypes.create_string_buffer(a, max(len(a), 1))
    bb = ctypes.create_string_buffer(b, max(len(b), 1))
    return bool(lib().oracle_streams_equal(ctypes.cast(ba, ctypes.c_void_p), len(a),
                                           ctypes.cast(bb, ctypes.c_void_p), len(b)))


def files_are_equal(a: str, b: str) -> bool:
    return bool(lib().oracle_files_are_equal(os.fsencode(a), os.fsencode(b)))


def cmp_batch(a: np.ndarray, b: np.ndarray, offsets: np.ndarray, lengths: np.ndarray,
              nthreads: int = 1) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
    n = len(offsets)
    out = np.zeros(n, dtype=np.uint8)
    if n:
        rc = lib().oracle_cmp_batch(_ptr(a), _ptr(b), _ptr(offsets), _ptr(lengths), n, _ptr(out), int(nthreads))
        if rc != 0:
            raise RuntimeError(f"oracle_cmp_batch failed: {rc}")
    return out


def cmp_algorithmic_bytes(a: np.ndarray, b: np.ndarray, offsets, lengths) -> int:
    """Bytes the reference's streamsEqual reads over a batch."""
    total = 0
    L = lib()
    for off, ln in zip(offsets, lengths):
        total += L.oracle_cmp_algorithmic_bytes(_ptr(a) + int(off), _ptr(b) + int(off), int(ln))
    return total


def _glob_star(directory: str) -> list[str]:
    """filepath.Glob(dir/*): sorted names; '*' also matches dotfiles in Go."""
    try:
        names = sorted(os.listdir(os.fsencode(directory)))
    except OSError:
        return []
    return [os.path.join(directory, os.fsdecode(n)) for n in names]


def dir_updated(dir_a: str, dir_b: str, pfx: str) -> dict[str, bool]:
    """helpers.DirUpdated (cmp.go:97-114)."""
    updated: dict[str, bool] = {}
    for file_a in _glob_star(dir_a):
        if os.path.isdir(file_a):           # IsDirectory: os.Stat, follows symlinks
            continue
        name = os.path.basename(file_a)
        file_b = os.path.join(dir_b, name)
        if os.path.exists(file_b) and not files_are_equal(file_a, file_b):   # FileExists: os.Stat
            updated[pfx + name] = True
    return updated


def apparmor_delta(old_path: str, new_path: str, prefix: str):
    """policy.AppArmorDelta (policy/policy.go:155-167)."""
    newaa = os.path.join(new_path, "meta", "framework-policy", "apparmor")
    oldaa = os.path.join(old_path, "meta", "framework-policy", "apparmor")
    return (dir_updated(os.path.join(oldaa, "policygroups"), os.path.join(newaa, "policygroups"), prefix),
            dir_updated(os.path.join(oldaa, "templates"), os.path.join(newaa, "templates"), prefix))


# --------------------------------------------------------------------------- mode strings

class UnknownFileMode(Exception):
    pass


def go_mode_string(st_mode: int) -> str:
    """Go's os.FileMode.String() for an lstat mode (used only in the error text)."""
    letters = ""
    if stat.S_ISDIR(st_mode):
        letters += "d"
    if stat.S_ISLNK(st_mode):
        letters += "L"
    if stat.S_ISBLK(st_mode):
        letters += "D"
    if stat.S_ISCHR(st_mode):
        letters += "Dc"
    if stat.S_ISFIFO(st_mode):
        letters += "p"
    if stat.S_ISSOCK(st_mode):
        letters += "S"
    if st_mode & stat.S_ISUID:
        letters += "u"
    if st_mode & stat.S_ISGID:
        letters += "g"
    if st_mode & stat.S_ISVTX:
        letters += "t"
    # Go prints the letters in the fixed order "dalTLDpSugct"
    order = "dalTLDpSugct"
    letters = "".join(sorted(set(letters), key=order.index))
    if not letters:
        letters = "-"
    rwx = "rwxrwxrwx"
    perm = "".join(c if st_mode & (1 << (8 - i)) else "-" for i, c in enumerate(rwx))
    return letters + perm


def file_mode_string(st_mode: int) -> str:
    """yamlFileMode.MarshalYAML (snappy/hashes.go:33-57) applied to an lstat mode."""
    if stat.S_ISDIR(st_mode):
        t = "d"
    elif stat.S_ISLNK(st_mode):
        t = "l"
    elif stat.S_ISREG(st_mode):
        t = "f"
    else:
        raise UnknownFileMode("Unknown file mode " + go_mode_string(st_mode))
    rwx = "rwxrwxrwx"
    return t + "".join(c if st_mode & (1 << (8 - i)) else "-" for i, c in enumerate(rwx))


def parse_mode_string(s: str) -> int:
    """yamlFileMode.UnmarshalYAML (snappy/hashes.go:59-88) -> st_mode style bits."""
    t = {"d": stat.S_IFDIR, "f": stat.S_IFREG, "l": stat.S_IFLNK}.get(s[0])
    if t is None:
        raise UnknownFileMode("Unknown file mode " + s)
    m = t
    for i, c in enumerate(s[1:10]):
        if c == "rwxrwxrwx"[i]:
            m |= 1 << (8 - i)
    return m


# --------------------------------------------------------------------------- yaml.v2 emitter restatement

_RESOLVE_MAP = set()
for _words in (("y", "Y", "yes", "Yes", "YES"), ("true", "True", "TRUE"), ("on", "On", "ON"),
               ("n", "N", "no", "No", "NO"), ("false", "False", "FALSE"), ("off", "Off", "OFF"),
               ("", "~", "null", "Null", "NULL"), (".nan", ".NaN", ".NAN"), (".inf", ".Inf", ".INF"),
               ("+.inf", "+.Inf", "+.INF"), ("-.inf", "-.Inf", "-.INF"), ("<<",)):
    _RESOLVE_MAP.update(w.encode() for w in _words)

_GO_INT = re.compile(rb"^[+-]?(0[xX][0-9a-fA-F]+|0[0-7]*|[1-9][0-9]*)$")
_GO_FLOAT = re.compile(rb"^[+-]?(([0-9]+\.?[0-9]*|\.[0-9]+)([eE][+-]?[0-9]+)?|inf|infinity|nan)$", re.I)
_BASE60 = re.compile(rb"^[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+(?:\.[0-9_]*)?$")


def _go_parse_int_ok(s: bytes) -> bool:
    """strconv.ParseInt/ParseUint(s, 0, 64) succeeds (Go 1.x of 2015: 0x hex, leading-0 octal)."""
    if not _GO_INT.match(s):
        return False
    body = s.lstrip(b"+-")
    neg = s.startswith(b"-")
    if body[:2].lower() == b"0x":
        v = int(body[2:], 16)
    elif len(body) > 1 and body[0:1] == b"0":
        v = int(body, 8)
    else:
        v = int(body)
    if neg:
        return v <= 1 << 63              # ParseInt range
    if s.startswith(b"+"):
        return v < 1 << 63               # ParseUint takes no sign
    return v < 1 << 64                   # ParseInt, else ParseUint


def _go_parse_float_ok(s: bytes) -> bool:
    if not _GO_FLOAT.match(s):
        return False
    body = s.lstrip(b"+-").lower()
    if body in (b"inf", b"infinity", b"nan"):
        return True
    try:
        v = float(body)
    except ValueError:
        return False
    return v != float("inf")          # ParseFloat reports ErrRange on overflow


def _resolves_to_non_string(s: bytes) -> bool:
    """yaml.v2 resolve("", s) returns a tag other than !!str (for valid UTF-8 input)."""
    if s == b"":
        return True                    # null
    c = s[0:1]
    if c in b"yYnNtTfFoO~<" or c == b".":
        if s in _RESOLVE_MAP:
            return True
        if c == b".":
            return _go_parse_float_ok(s)
        return False
    if c in b"+-0123456789":
        if s in _RESOLVE_MAP:
            return True
        plain = s.replace(b"_", b"")
        if _go_parse_int_ok(plain) or _go_parse_float_ok(plain):
            return True
        for pre in (b"0b", b"-0b"):
            if plain.startswith(pre):
                digits = plain[len(pre):]
                if digits and set(digits) <= set(b"01") and int(digits, 2) < (1 << 64 if pre == b"0b" else (1 << 63) + 1):
                    return True
        return False
    return False


def _is_base60_float(s: bytes) -> bool:
    if not s or s[0:1] not in b"+-0123456789" or b":" not in s:
        return False
    return bool(_BASE60.match(s))


def _utf8_valid(b: bytes) -> bool:
    try:
        b.decode("utf-8")              # strict: rejects surrogates, overlongs, > U+10FFFF
        return True
    except UnicodeDecodeError:
        return False


def _printable(cp: int) -> bool:
    """yaml.v2 yamlprivateh.go is_printable, on a decoded code point."""
    if cp == 0x0A or 0x20 <= cp <= 0x7E:
        return True
    if cp < 0xA0:
        return False                   # C0/C1 controls, DEL, NEL (0x85)
    if cp <= 0xD7FF:
        return True
    if 0xE000 <= cp <= 0xFFFD and cp != 0xFEFF:
        return True
    if cp >= 0x10000:
        return False                   # 4-byte sequences (lead byte 0xF0+) are not in the list
    return False


_BREAKS = (0x0D, 0x0A, 0x85, 0x2028, 0x2029)


class _Emitter:
    """Just enough of yaml.v2's emitterc.go for block mappings/sequences of scalars."""

    BEST_WIDTH = 80
    BEST_INDENT = 2

    def __init__(self):
        self.out: list[str] = []
        self.column = 0
        self.whitespace = True
        self.indention = True
        self.indent = 0

    # -- low level writers
    def put(self, ch: str):
        self.out.append(ch)
        self.column += 1

    def put_break(self):
        self.out.append("\n")
        self.column = 0

    def write_indent(self):
        indent = max(self.indent, 0)
        if not self.indention or self.column > indent or (self.column == indent and not self.whitespace):
            self.put_break()
        while self.column < indent:
            self.put(" ")
        self.whitespace = True
        self.indention = True

    def write_indicator(self, ind: str, need_ws: bool, is_ws: bool, is_ind: bool):
        if need_ws and not self.whitespace:
            self.put(" ")
        for ch in ind:
            self.put(ch)
        self.whitespace = is_ws
        self.indention = self.indention and is_ind

    # -- scalar analysis (yaml_emitter_analyze_scalar)
    @staticmethod
    def analyze(text: str) -> dict:
        r = dict(multiline=False, flow_plain=False, block_plain=True, single=True, block=False)
        if not text:
            return r
        cps = [ord(c) for c in text]
        n = len(cps)
        block_ind = flow_ind = line_breaks = special = False
        lead_sp = lead_br = trail_sp = trail_br = break_space = space_break = False
        prev_space = prev_break = False
        if text.startswith("---") or text.startswith("..."):
            block_ind = flow_ind = True
        preceded_ws = True
        for i, cp in enumerate(cps):
            followed_ws = i + 1 >= n or cps[i + 1] in (0x20, 0x09)
            ch = text[i]
            if i == 0:
                if ch in "#,[]{}&*!|>'\"%@`":
                    flow_ind = block_ind = True
                elif ch in "?:":
                    flow_ind = True
                    if followed_ws:
                        block_ind = True
                elif ch == "-" and followed_ws:
                    flow_ind = block_ind = True
            else:
                if ch in ",?[]{}":
                    flow_ind = True
                elif ch == ":":
                    flow_ind = True
                    if followed_ws:
                        block_ind = True
                elif ch == "#" and preceded_ws:
                    flow_ind = block_ind = True
            if not _printable(cp):
                special = True
            if cp == 0x20:
                if i == 0:
                    lead_sp = True
                if i == n - 1:
                    trail_sp = True
                if prev_break:
                    break_space = True
                prev_space, prev_break = True, False
            elif cp in _BREAKS:
                line_breaks = True
                if i == 0:
                    lead_br = True
                if i == n - 1:
                    trail_br = True
                if prev_space:
                    space_break = True
                prev_space, prev_break = False, True
            else:
                prev_space = prev_break = False
            preceded_ws = cp in (0x20, 0x09) or cp in _BREAKS or cp == 0
        r.update(multiline=line_breaks, flow_plain=True, block_plain=True, single=True, block=True)
        if lead_sp or lead_br or trail_sp or trail_br:
            r["flow_plain"] = r["block_plain"] = False
        if trail_sp:
            r["block"] = False
        if break_space:
            r["flow_plain"] = r["block_plain"] = r["single"] = False
        if space_break or special:
            r["flow_plain"] = r["block_plain"] = r["single"] = r["block"] = False
        if line_breaks:
            r["flow_plain"] = r["block_plain"] = False
        if flow_ind:
            r["flow_plain"] = False
        if block_ind:
            r["block_plain"] = False
        return r

    # -- scalar writers (block context, mapping value => allow_breaks is True)
    def write_plain(self, text: str):
        if not self.whitespace:
            self.put(" ")
        spaces = False
        n = len(text)
        for i, ch in enumerate(text):
            if ch == " ":
                if not spaces and self.column > self.BEST_WIDTH and not (i + 1 < n and text[i + 1] == " "):
                    self.write_indent()
                else:
                    self.put(ch)
                spaces = True
            else:
                self.put(ch)
                self.indention = False
                spaces = False
        self.whitespace = False
        self.indention = False

    def write_single(self, text: str):
        self.write_indicator("'", True, False, False)
        spaces = False
        n = len(text)
        for i, ch in enumerate(text):
            if ch == " ":
                if (not spaces and self.column > self.BEST_WIDTH and 0 < i < n - 1
                        and not (i + 1 < n and text[i + 1] == " ")):
                    self.write_indent()
                else:
                    self.put(ch)
                spaces = True
            else:
                if ch == "'":
                    self.put("'")
                self.put(ch)
                self.indention = False
                spaces = False
        self.write_indicator("'", False, False, False)

    _ESC = {0x00: "0", 0x07: "a", 0x08: "b", 0x09: "t", 0x0A: "n", 0x0B: "v", 0x0C: "f", 0x0D: "r",
            0x1B: "e", 0x22: '"', 0x5C: "\\", 0x85: "N", 0xA0: "_", 0x2028: "L", 0x2029: "P"}

    def write_double(self, text: str):
        self.write_indicator('"', True, False, False)
        spaces = False
        n = len(text)
        i = 0
        while i < n:
            ch = text[i]
            cp = ord(ch)
            if not _printable(cp) or cp == 0xFEFF or cp in _BREAKS or ch in '"\\':
                self.put("\\")
                if cp in self._ESC:
                    self.put(self._ESC[cp])
                elif cp <= 0xFF:
                    for c in "x%02X" % cp:
                        self.put(c)
                elif cp <= 0xFFFF:
                    for c in "u%04X" % cp:
                        self.put(c)
                else:
                    for c in "U%08X" % cp:
                        self.put(c)
                spaces = False
                i += 1
            elif ch == " ":
                if not spaces and self.column > self.BEST_WIDTH and 0 < i < n - 1:
                    self.write_indent()
                    i += 1
                    if i < n and text[i] == " ":
                        self.put("\\")
                else:
                    self.put(ch)
                    i += 1
                spaces = True
            else:
                self.put(ch)
                spaces = False
                i += 1
        self.write_indicator('"', False, False, False)

    def write_literal(self, text: str):
        self.write_indicator("|", True, False, False)
        # block scalar hints
        hint = ""
        if text and (text[0] == " " or ord(text[0]) in _BREAKS):
            hint += str(self.BEST_INDENT)
        if not text or ord(text[-1]) not in _BREAKS:
            hint += "-"
        elif len(text) == 1 or ord(text[-2]) in _BREAKS:
            hint += "+"
        if hint:
            self.write_indicator(hint, False, False, False)
        self.put_break()
        self.indention = True
        self.whitespace = True
        breaks = True
        for ch in text:
            if ord(ch) in _BREAKS:
                self.put_break()           # every break is written as "\n" except LS/PS/NEL ...
                self.indention = True
                breaks = True
            else:
                if breaks:
                    self.write_indent()
                self.put(ch)
                self.indention = False
                breaks = False

    # -- yaml.v2 encoder.stringv + emit_scalar
    def emit_string_value(self, raw: bytes):
        tag = ""
        if _utf8_valid(raw):
            text = raw.decode("utf-8")
            non_str = _resolves_to_non_string(raw)
        else:
            enc = base64.b64encode(raw).decode("ascii")
            lines = len(enc) // 70 + 1
            if lines > 1:
                enc = "".join(enc[i:i + 70] + "\n" for i in range(0, len(enc), 70))
            text, tag, non_str = enc, "!!binary", False
        if not tag and (non_str or _is_base60_float(raw)):
            style = "double"
        elif "\n" in text:
            style = "literal"
        else:
            style = "plain"
        a = self.analyze(text)
        # select_scalar_style, block context, not a simple key
        if style == "plain":
            if not a["block_plain"]:
                style = "single"
        if style == "single" and not a["single"]:
            style = "double"
        if style == "literal" and not a["block"]:
            style = "double"
        if tag:
            self.write_indicator(tag, True, False, False)
        saved = self.indent
        self.indent = self.indent + self.BEST_INDENT if self.indent >= 0 else self.BEST_INDENT
        {"plain": self.write_plain, "single": self.write_single,
         "double": self.write_double, "literal": self.write_literal}[style](text)
        self.indent = saved

    def emit_plain_value(self, text: str):
        """ints and other values yaml.v2 writes as implicit plain scalars."""
        saved = self.indent
        self.indent += self.BEST_INDENT
        self.write_plain(text)
        self.indent = saved

    def key(self, name: str, first_in_seq_item: bool = False):
        if not first_in_seq_item:
            self.write_indent()
        self.write_plain(name)           # keys here are short plain ASCII (simple keys)
        self.write_indicator(":", False, False, False)


def marshal_file_hash(em: _Emitter, entry: dict, in_sequence: bool):
    """One fileHash mapping: name, size (omitempty on nil), sha512 (omitempty on ""), mode."""
    first = in_sequence
    em.key("name", first_in_seq_item=first)
    em.emit_string_value(entry["name"])
    if entry.get("size") is not None:
        em.key("size")
        em.emit_plain_value(str(int(entry["size"])))
    if entry.get("sha512"):
        em.key("sha512")
        em.emit_string_value(entry["sha512"].encode("ascii"))
    em.key("mode")
    em.emit_string_value(entry["mode"].encode("ascii"))


def marshal_hashes(archive_sha512: str, files: list[dict]) -> bytes:
    """yaml.Marshal(hashesYaml{...}) (snappy/build.go:264)."""
    em = _Emitter()
    em.key("archive-sha512")
    em.emit_string_value(archive_sha512.encode("ascii"))
    em.key("files")
    if not files:
        em.write_indicator("[", True, True, False)
        em.write_indicator("]", False, False, False)
    else:
        for f in files:
            em.write_indent()                       # indentless sequence inside a mapping
            em.write_indicator("-", True, False, True)
            em.indent = 2
            marshal_file_hash(em, f, in_sequence=True)
            em.indent = 0
    em.put_break()
    return "".join(em.out).encode("utf-8")


def marshal_single_file_hash(entry: dict) -> bytes:
    """yaml.Marshal(&fileHash{...}) as in snappy/hashes_test.go:50-55."""
    em = _Emitter()
    marshal_file_hash(em, entry, in_sequence=False)
    em.put_break()
    return "".join(em.out).encode("utf-8")


# --------------------------------------------------------------------------- writeHashes

def _walk(root: bytes):
    """filepath.Walk order: (path, lstat) pre-order, children sorted bytewise.

    Mirrors Go's quirk that a directory whose listing fails is reported to the callback a
    second time (with the error, which writeHashes ignores: build.go:228,241)."""
    st = os.lstat(root)
    yield root, st
    if stat.S_ISDIR(st.st_mode):
        yield from _walk_children(root)


def _walk_children(path: bytes):
    try:
        names = sorted(os.listdir(path))
    except OSError:
        yield path, os.lstat(path)
        return
    for name in names:
        child = path + b"/" + name
        try:
            st = os.lstat(child)
        except OSError:
            continue                                  # Go would hand a nil FileInfo to the callback (panic)
        yield child, st
        if stat.S_ISDIR(st.st_mode):
            yield from _walk_children(child)


def collect_hashes(build_dir: str, data_tar: str, hasher=None):
    """The walk of writeHashes without the final marshal; returns (archive_sha512, entries).

    ``hasher`` maps a path to its hex digest (defaults to the oracle's Sha512sum), so that a
    test can feed GPU digests through the same host logic."""
    hasher = hasher or sha512sum
    root = os.fsencode(build_dir.rstrip("/") or "/")
    os.makedirs(os.path.join(build_dir, "DEBIAN"), mode=0o755, exist_ok=True)
    archive = hasher(data_tar)
    entries = []
    for path, st in _walk(root):
        rel = path[len(root):]
        if rel.startswith(b"/DEBIAN"):                 # build.go:229 -- prefix test, not a path test
            continue
        if path == root:
            continue
        e = {"name": rel[1:], "size": None, "sha512": "", "st_mode": st.st_mode}
        if stat.S_ISREG(st.st_mode):
            e["sha512"] = hasher(os.fsdecode(path))
            e["size"] = st.st_size
        entries.append(e)
    return archive, entries


def write_hashes(build_dir: str, data_tar: str, hasher=None) -> bytes:
    """writeHashes (snappy/build.go:216-270). Writes DEBIAN/hashes.yaml, returns its bytes."""
    archive, entries = collect_hashes(build_dir, data_tar, hasher)
    for e in entries:
        e["mode"] = file_mode_string(e["st_mode"])    # Marshal-time error for fifo/socket/device
    content = marshal_hashes(archive, entries)
    out = os.path.join(build_dir, "DEBIAN", "hashes.yaml")
    with open(out, "wb") as f:
        f.write(content)
    os.chmod(out, 0o644)
    return content


# ---------------------------------------------------------------------------------------------
# copyToBuildDir (/root/reference/snappy/build.go:362-418) -- SURVEY.md section 8f, row 2
# ---------------------------------------------------------------------------------------------

# snappy/build.go:52-83, joined with "|" exactly as the reference does.  Go's regexp differs from
# Python's in two ways that matter here: `$` matches only at the very end of the text (Python's
# also matches before a trailing newline) and `{arch}` is a literal; hence \Z and the escapes.
_SHOULD_EXCLUDE = re.compile("|".join([
    r"\.snap\Z", r"\.click\Z", r"^\..*\.sw.\Z", r"~\Z", r"^,,", r"^\.[#~]", r"^\.arch-ids\Z", r"^\.arch-inventory\Z",
    r"^\.bzr\Z", r"^\.bzr-builddeb\Z", r"^\.bzr\.backup\Z", r"^\.bzr\.tags\Z", r"^\.bzrignore\Z", r"^\.cvsignore\Z",
    r"^\.git\Z", r"^\.gitattributes\Z", r"^\.gitignore\Z", r"^\.gitmodules\Z", r"^\.hg\Z", r"^\.hgignore\Z",
    r"^\.hgsigs\Z", r"^\.hgtags\Z", r"^\.shelf\Z", r"^\.svn\Z", r"^CVS\Z", r"^DEADJOE\Z", r"^RCS\Z", r"^_MTN\Z",
    r"^_darcs\Z", r"^\{arch\}\Z",
]).encode())


def should_exclude(basename) -> bool:
    """shouldExclude (snappy/build.go:52-83) on a base name."""
    return _SHOULD_EXCLUDE.search(os.fsencode(basename)) is not None


def copy_to_build_dir(source_dir: str, build_dir: str, no_link: bool = False) -> None:
    """copyToBuildDir (snappy/build.go:362-418): Walk order, exclusions, Mkdir with the source's
    mode, hard link where possible, else open / O_EXCL create / io.Copy.  Raises OSError where Go
    returns the error."""
    source = os.fsencode(os.path.abspath(source_dir))
    build = os.fsencode(build_dir)
    try:
        try:
            os.unlink(build)                       # os.Remove: unlink, then rmdir
        except (IsADirectoryError, PermissionError):
            os.rmdir(build)
    except FileNotFoundError:
        pass

    def visit(path: bytes, st) -> None:
        if should_exclude(os.path.basename(path)):
            return                                  # SkipDir for a directory, skip for a file
        dest = build + path[len(source):]
        if stat.S_ISDIR(st.st_mode):
            os.mkdir(dest, stat.S_IMODE(st.st_mode))
            for name in sorted(os.listdir(path)):
                child = path + b"/" + name
                visit(child, os.lstat(child))
            return
        if not no_link:
            try:
                os.link(path, dest, follow_symlinks=False)
                return
            except OSError:
                pass
        with open(path, "rb") as fin:               # os.Open follows symlinks
            fd = os.open(dest, os.O_WRONLY | os.O_CREAT | os.O_EXCL, stat.S_IMODE(st.st_mode))
            with os.fdopen(fd, "wb") as fout:
                while True:
                    chunk = fin.read(1 << 20)
                    if not chunk:
                        break
                    fout.write(chunk)

    visit(source, os.lstat(source))


# ---------------------------------------------------------------------------------------------
# hashes.yaml verification -- SURVEY.md section 8f, row 4 (no function in the reference does this;
# the fields compared are those of fileHash, snappy/hashes.go:93-101)
# ---------------------------------------------------------------------------------------------

def verify_hashes(root: str, yaml_path: str, data_tar: str | None = None) -> list[str]:
    """Report lines in the order and spelling of snapgpu_verify_hashes, for documents whose names
    are plain scalars.  The old document is read with PyYAML (an independent parser), the tree is
    hashed with the oracle's own Sha512sum."""
    import yaml
    old = yaml.safe_load(open(yaml_path, "rb").read())
    base = os.fsencode(root.rstrip("/") or "/")
    fresh = {}
    order = []
    for path, st in _walk(base):
        rel = path[len(base):]
        if rel.startswith(b"/DEBIAN") or path == base:
            continue
        name = os.fsdecode(rel[1:])
        e = {"mode": file_mode_string(st.st_mode)}
        if stat.S_ISREG(st.st_mode):
            e["size"] = st.st_size
            e["sha512"] = sha512sum(os.fsdecode(path))
        fresh[name] = e
        order.append(name)
    report = []
    if data_tar is not None and old.get("archive-sha512") != sha512sum(data_tar):
        report.append("archive-sha512 differs")
    seen = set()
    for ent in old.get("files") or []:
        name = str(ent["name"])
        seen.add(name)
        if name not in fresh:
            report.append(f"missing: {name}")
            continue
        what = [k for k in ("size", "sha512", "mode") if str(ent.get(k, "")) != str(fresh[name].get(k, ""))]
        if what:
            report.append(f"changed: {name} ({' '.join(what)})")
    self_name = None
    ap, ar = os.path.abspath(yaml_path), os.path.abspath(root)
    if ap.startswith(ar + "/"):
        self_name = ap[len(ar) + 1:]
    for name in order:
        if name not in seen and name != self_name:
            report.append(f"extra: {name}")
    return report
