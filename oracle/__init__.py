"""ctypes binding of the CPU oracle (oracle/qd_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (quickdna_b200/) never imports this.

Parity status: pinned -- see the header of oracle/qd_oracle.h.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqd_oracle.so")

KIND_NAMES = {
    0: "Ok",
    1: "NonAsciiByte",
    2: "BadNucleotide",
    3: "UnexpectedAmbiguousNucleotide",
    4: "BadTranslationTable",
}


class OracleError(ValueError):
    """Mirrors the ValueError the reference's PyO3 layer raises (src/python_api.rs:15-19)."""

    def __init__(self, kind: int, byte: int, pos: int, message: str):
        super().__init__(message)
        self.kind = kind
        self.byte = byte
        self.pos = pos


class _Err(C.Structure):
    _fields_ = [("kind", C.c_int32), ("byte", C.c_uint8), ("pos", C.c_uint64)]


def _cpu_signature() -> str:
    """What `-march=native` depends on: the library built on one host must not be loaded on another."""
    import hashlib

    model, flags = "", ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name") and not model:
                    model = line.split(":", 1)[1].strip()
                elif line.startswith("flags") and not flags:
                    flags = " ".join(sorted(line.split(":", 1)[1].split()))
                if model and flags:
                    break
    except OSError:
        pass
    with open(os.path.join(_HERE, "Makefile"), "rb") as f:
        mk = f.read()
    return hashlib.sha256((model + "|" + flags).encode() + mk).hexdigest()[:16]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it.

    Rebuilt when a source is newer than the library or when the library was built for another CPU (the snapshot that
    travels to the GPU box carries the library built in the development container)."""
    src = os.path.join(_HERE, "qd_oracle.c")
    sig_path = os.path.join(_HERE, ".build_sig")
    sig = _cpu_signature()
    try:
        with open(sig_path) as f:
            built_for = f.read().strip()
    except OSError:
        built_for = ""
    stale = (not os.path.exists(_LIB_PATH)) or built_for != sig or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("qd_oracle.c", "qd_oracle.h", "ncbi_codes.h", "Makefile")
    )
    if force or stale:
        r = subprocess.run(["make", "-C", _HERE, "-B", "libqd_oracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/libqd_oracle.so failed:\n" + r.stdout + r.stderr)
        with open(sig_path, "w") as f:
            f.write(sig + "\n")
    assert os.path.exists(src)
    return _LIB_PATH


_lib: Optional[C.CDLL] = None

_u8p = C.POINTER(C.c_uint8)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.qdo_tables.restype = _u8p
        L.qdo_table_slot.argtypes = [C.c_int]
        L.qdo_err_message.argtypes = [C.POINTER(_Err), C.c_char_p, C.c_size_t]
        L.qdo_translate_bytes.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(_Err)]
        L.qdo_translate_masks.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.qdo_translate_masks.restype = None
        L.qdo_revcomp_bytes.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(_Err)]
        L.qdo_revcomp_masks.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.qdo_revcomp_masks.restype = None
        L.qdo_self_frames_masks.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        L.qdo_all_frames_masks.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t)]
        for f in (L.qdo_py_self_frames, L.qdo_py_all_frames):
            f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t),
                          C.POINTER(C.c_int), C.POINTER(_Err)]
        L.qdo_parse.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(_Err)]
        L.qdo_forward_canonical.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.qdo_forward_canonical.restype = None
        L.qdo_canonical.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.qdo_canonical.restype = None
        L.qdo_canonical_windows.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.qdo_canonical_windows.restype = C.c_size_t
        L.qdo_expansions_size_hint.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.qdo_expansions.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.qdo_expansions.restype = C.c_size_t
        L.qdo_windows.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.qdo_windows.restype = C.c_size_t
        L.qdo_windows_all.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(C.c_size_t)]
        L.qdo_windows_all.restype = C.c_size_t
        L.qdo_bench_all_frames_batch.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
        L.qdo_bench_all_frames_batch.restype = C.c_double
        L.qdo_bench_revcomp.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int]
        L.qdo_bench_revcomp.restype = C.c_double
        L.qdo_synth.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_char_p, C.c_size_t, C.c_void_p]
        L.qdo_synth.restype = None
        L.qdo_checksum64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.qdo_checksum64.restype = C.c_uint64
        L.qdo_check_frames_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.qdo_check_frames_uniform.restype = C.c_uint64
        L.qdo_check_revcomp.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        L.qdo_check_revcomp.restype = C.c_uint64
        L.qdo_check_windows.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        L.qdo_check_windows.restype = C.c_uint64
        L.qdo_check_expansions.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
        L.qdo_check_expansions.restype = C.c_uint64
        _lib = L
    return _lib


# --------------------------------------------------------------------------- helpers


def _as_u8(x) -> np.ndarray:
    if isinstance(x, str):
        x = x.encode("latin-1")
    if isinstance(x, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(x), dtype=np.uint8)
    return np.ascontiguousarray(x, dtype=np.uint8)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _raise(err: _Err):
    buf = C.create_string_buffer(128)
    lib().qdo_err_message(C.byref(err), buf, 128)
    raise OracleError(err.kind, err.byte, err.pos, buf.value.decode("ascii"))


ASCII_OF_MASK = b"?ATWCMYHGRKDSVBN"
_MASK_OF_ASCII = np.zeros(256, dtype=np.uint8)
for _m, _c in enumerate(ASCII_OF_MASK):
    if _m:
        _MASK_OF_ASCII[_c] = _m
        _MASK_OF_ASCII[ord(chr(_c).lower())] = _m


def masks_of(ascii_seq) -> np.ndarray:
    """Plain table decode of already-valid ASCII (test convenience, no whitespace handling)."""
    m = _MASK_OF_ASCII[_as_u8(ascii_seq)]
    assert (m != 0).all()
    return m


def ascii_of(masks) -> bytes:
    return bytes(np.frombuffer(ASCII_OF_MASK, dtype=np.uint8)[_as_u8(masks)])


# --------------------------------------------------------------------------- API


def tables() -> bytes:
    return C.string_at(lib().qdo_tables(), 27 * 4096)


def table_slot(ncbi_id: int) -> int:
    return lib().qdo_table_slot(int(ncbi_id))


def err_message(kind: int, byte: int) -> str:
    e = _Err(kind, byte, 0)
    buf = C.create_string_buffer(128)
    lib().qdo_err_message(C.byref(e), buf, 128)
    return buf.value.decode("ascii")


def translate_bytes(table: int, dna, strict: bool = False) -> bytes:
    a = _as_u8(dna)
    out = np.empty(len(a) // 3, dtype=np.uint8)
    err = _Err()
    if lib().qdo_translate_bytes(int(table), int(strict), _ptr(a), len(a), _ptr(out), C.byref(err)) != 0:
        _raise(err)
    return out.tobytes()


def translate_masks(slot: int, masks) -> bytes:
    a = _as_u8(masks)
    out = np.empty(len(a) // 3, dtype=np.uint8)
    lib().qdo_translate_masks(int(slot), _ptr(a), len(a), _ptr(out))
    return out.tobytes()


def revcomp_bytes(dna, strict: bool = False) -> bytes:
    a = _as_u8(dna)
    out = np.empty(len(a), dtype=np.uint8)
    err = _Err()
    if lib().qdo_revcomp_bytes(int(strict), _ptr(a), len(a), _ptr(out), C.byref(err)) != 0:
        _raise(err)
    return out.tobytes()


def revcomp_masks(masks) -> np.ndarray:
    a = _as_u8(masks)
    out = np.empty(len(a), dtype=np.uint8)
    lib().qdo_revcomp_masks(_ptr(a), len(a), _ptr(out))
    return out


def _split(out: np.ndarray, lens, n_frames_total: int) -> List[bytes]:
    frames, off = [], 0
    for ln in lens:
        if ln > 0:
            frames.append(out[off:off + ln].tobytes())
        off += ln
    assert len(frames) == n_frames_total
    return frames


def self_frames_masks(slot: int, masks) -> List[bytes]:
    a = _as_u8(masks)
    out = np.empty(len(a) + 8, dtype=np.uint8)
    lens = (C.c_size_t * 3)()
    nf = lib().qdo_self_frames_masks(int(slot), _ptr(a), len(a), _ptr(out), lens)
    return _split(out, list(lens), nf)


def all_frames_masks(slot: int, masks) -> List[bytes]:
    a = _as_u8(masks)
    out = np.empty(2 * len(a) + 8, dtype=np.uint8)
    lens = (C.c_size_t * 6)()
    nf = lib().qdo_all_frames_masks(int(slot), _ptr(a), len(a), _ptr(out), lens)
    return _split(out, list(lens), nf)


def py_self_frames(table: int, dna, strict: bool = False) -> List[bytes]:
    a = _as_u8(dna)
    out = np.empty(len(a) + 8, dtype=np.uint8)
    lens = (C.c_size_t * 3)()
    nf = C.c_int()
    err = _Err()
    if lib().qdo_py_self_frames(int(table), int(strict), _ptr(a), len(a), _ptr(out), lens, C.byref(nf), C.byref(err)) != 0:
        _raise(err)
    return _split(out, list(lens), nf.value)


def py_all_frames(table: int, dna, strict: bool = False) -> List[bytes]:
    a = _as_u8(dna)
    out = np.empty(2 * len(a) + 8, dtype=np.uint8)
    lens = (C.c_size_t * 6)()
    nf = C.c_int()
    err = _Err()
    if lib().qdo_py_all_frames(int(table), int(strict), _ptr(a), len(a), _ptr(out), lens, C.byref(nf), C.byref(err)) != 0:
        _raise(err)
    return _split(out, list(lens), nf.value)


def parse(ascii_seq, strict: bool = False) -> np.ndarray:
    a = _as_u8(ascii_seq)
    out = np.empty(len(a), dtype=np.uint8)
    n = C.c_size_t()
    err = _Err()
    if lib().qdo_parse(int(strict), _ptr(a), len(a), _ptr(out), C.byref(n), C.byref(err)) != 0:
        _raise(err)
    return out[: n.value].copy()


def forward_canonical(masks) -> np.ndarray:
    a = _as_u8(masks)
    out = np.empty(len(a), dtype=np.uint8)
    lib().qdo_forward_canonical(_ptr(a), len(a), _ptr(out))
    return out


def canonical(masks) -> np.ndarray:
    a = _as_u8(masks)
    out = np.empty(len(a), dtype=np.uint8)
    lib().qdo_canonical(_ptr(a), len(a), _ptr(out))
    return out


def canonical_windows(masks, w: int) -> np.ndarray:
    a = _as_u8(masks)
    count = max(0, len(a) - w + 1) if w > 0 else 0
    out = np.empty((count, w), dtype=np.uint8)
    got = lib().qdo_canonical_windows(_ptr(a), len(a), w, _ptr(out))
    assert got == count
    return out


def expansions_size_hint(masks) -> Optional[int]:
    a = _as_u8(masks)
    c = C.c_uint64()
    ok = lib().qdo_expansions_size_hint(_ptr(a), len(a), C.byref(c))
    return c.value if ok else None


def expansions(masks, max_out: int = 1 << 20) -> np.ndarray:
    a = _as_u8(masks)
    hint = expansions_size_hint(a)
    cap = max_out if hint is None else min(max_out, hint)
    out = np.empty((cap, len(a)), dtype=np.uint8)
    got = lib().qdo_expansions(_ptr(a), len(a), _ptr(out), cap)
    return out[:got]


def windows(seq, w: int) -> np.ndarray:
    a = _as_u8(seq)
    count = max(0, len(a) - w + 1) if w > 0 else 0
    out = np.empty((count, w), dtype=np.uint8)
    got = lib().qdo_windows(_ptr(a), len(a), w, _ptr(out))
    assert got == count
    return out


def windows_all(masks, w_dna: int = 42, w_aa: int = 20) -> Tuple[np.ndarray, np.ndarray, np.ndarray, List[int]]:
    """(dna windows [fwd then rc], aa windows [frames 0..5], origin indices, counts[8])."""
    a = _as_u8(masks)
    counts = (C.c_size_t * 8)()
    total = lib().qdo_windows_all(_ptr(a), len(a), w_dna, w_aa, None, None, None, counts)
    cs = list(counts)
    dna = np.empty((cs[0] + cs[1], w_dna), dtype=np.uint8)
    aa = np.empty((sum(cs[2:]), w_aa), dtype=np.uint8)
    origin = np.empty(total, dtype=np.uint64)
    lib().qdo_windows_all(_ptr(a), len(a), w_dna, w_aa, _ptr(dna), _ptr(aa), _ptr(origin), counts)
    return dna, aa, origin, cs


class _FastaRecord(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("header_begin", "header_end", "content_begin", "content_end", "line_begin", "line_end")]


class FastaError(OracleError):
    """Located<FastaParseError::ParseError(TranslationError)> (src/fasta.rs:413-423, src/errors.rs:8-15)."""

    def __init__(self, kind, byte, pos, line):
        super().__init__(kind, byte, pos, err_message(kind, byte))
        self.line = line

    def located(self) -> str:
        return f"on line {self.line}: error parsing record: {err_message(self.kind, self.byte)}"


def header_text(text: bytes, begin: int, end: int) -> str:
    """The header string of a record: its header lines without the leading '>' / ';', joined by newlines."""
    if begin == 2 ** 64 - 1:
        return ""
    lines = text[begin:end].split(b"\n")
    lines = [ln[:-1] if i + 1 < len(lines) and ln.endswith(b"\r") else ln for i, ln in enumerate(lines)]
    return "\n".join(ln[1:].decode("utf-8", "replace") for ln in lines)


def fasta_scan(text, strict: bool = False, concatenate_headers: bool = True, allow_preceding_comment: bool = False):
    """[(header, masks, (line_begin, line_end))] like FastaParser::<DnaSequence<T>>::parse_str (src/fasta.rs:425-480)."""
    raw = bytes(text) if not isinstance(text, str) else text.encode()
    a = _as_u8(raw)
    masks = np.empty(max(len(a), 1), dtype=np.uint8)
    cap = raw.count(b">") + raw.count(b";") + 2
    recs = (_FastaRecord * cap)()
    n_masks, n_recs, line = C.c_size_t(), C.c_size_t(), C.c_uint64()
    err = _Err()
    L = lib()
    L.qdo_fasta_scan.restype = C.c_int
    rc = L.qdo_fasta_scan(int(strict), int(concatenate_headers), int(allow_preceding_comment), _ptr(a), C.c_size_t(len(a)), _ptr(masks),
                          C.byref(n_masks), recs, C.c_size_t(cap), C.byref(n_recs), C.byref(err), C.byref(line))
    if rc != 0:
        raise FastaError(err.kind, err.byte, err.pos, line.value)
    out = []
    for r in recs[: n_recs.value]:
        out.append((header_text(raw, r.header_begin, r.header_end), masks[r.content_begin:r.content_end].copy(),
                    (int(r.line_begin), int(r.line_end))))
    return out


def bench_all_frames_batch(slot: int, dna: np.ndarray, n_reads: int, read_len: int, threads: int, reps: int = 1,
                           out: Optional[np.ndarray] = None) -> Tuple[float, np.ndarray]:
    a = _as_u8(dna)
    assert a.size == n_reads * read_len
    if out is None:
        out = np.empty(n_reads * 2 * max(read_len - 2, 0), dtype=np.uint8)
    secs = lib().qdo_bench_all_frames_batch(int(slot), _ptr(a), n_reads, read_len, _ptr(out), threads, reps)
    return secs, out


def bench_revcomp(dna: np.ndarray, threads: int, reps: int = 1) -> Tuple[float, np.ndarray]:
    a = _as_u8(dna)
    out = np.empty(a.size, dtype=np.uint8)
    secs = lib().qdo_bench_revcomp(_ptr(a), a.size, _ptr(out), threads, reps)
    return secs, out


# --------------------------------------------------------------------------- synthetic configs as checksums


def synth(seed: int, start: int, count: int, alphabet: bytes = b"ACGT") -> np.ndarray:
    """quickdna_b200/synth.py synth_np, in C (the whole-config checks generate their input with it)."""
    out = np.empty(count, dtype=np.uint8)
    lib().qdo_synth(seed, start, count, alphabet, len(alphabet), _ptr(out))
    return out


def checksum64(buf, base: int = 0) -> int:
    """The position-weighted linear checksum of qd_checksum64_dev: `buf` = bytes [base, base + len) of a stream."""
    a = _as_u8(buf)
    return int(lib().qdo_checksum64(_ptr(a), a.size, base))


def check_frames_uniform(seed: int, first_read: int, n_reads: int, read_len: int, slot: int = 0, frames: int = 6, threads: int = 1) -> int:
    return int(lib().qdo_check_frames_uniform(seed, first_read, n_reads, read_len, slot, frames, threads))


def check_revcomp(seed: int, start: int, n: int, strict: bool = False, threads: int = 1) -> int:
    return int(lib().qdo_check_revcomp(seed, start, n, int(strict), threads))


def check_canonical_windows(dna, n_seq: int, seq_len: int, w: int = 42, threads: int = 1) -> int:
    a = _as_u8(dna)
    assert a.size == n_seq * seq_len
    return int(lib().qdo_check_windows(0, _ptr(a), n_seq, seq_len, w, 0, threads))


def check_windows_all(dna, n_seq: int, seq_len: int, w_dna: int = 42, w_aa: int = 20, threads: int = 1) -> int:
    a = _as_u8(dna)
    assert a.size == n_seq * seq_len
    return int(lib().qdo_check_windows(1, _ptr(a), n_seq, seq_len, w_dna, w_aa, threads))


def check_expansions(windows, cap: int = 16, threads: int = 1) -> Tuple[int, int]:
    """(checksum of the expansion rows, number of rows) of a [n_win, w] uint8 array of ASCII windows."""
    a = _as_u8(windows)
    assert a.ndim == 2
    rows = C.c_uint64()
    s = lib().qdo_check_expansions(_ptr(a), a.shape[0], a.shape[1], cap, threads, C.byref(rows))
    return int(s), int(rows.value)
