"""Batch and device-resident entry points (numpy / torch plumbing over the C ABI).

This is the SURVEY 8(f) rank-4 "Python batch API": what a caller looping
`DnaSequence.translate_all_frames` over many reads does (quickdna/__init__.py:163-185), in one
native call.  torch is used only for device memory and streams.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native as N


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        return np.frombuffer(a, dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def frame_len(n: int, k: int) -> int:
    return (n - k) // 3 if n > k else 0


def runs(n: int) -> int:
    return (n + 47) // 48


def slot_frame_offset(n: int, frames: int, f: int) -> int:
    """Offset of frame f's first residue inside a sequence's slot block (include/quickdna_b200.h)."""
    slot = 16 * runs(n)
    if f < 3 or frames != N.FRAMES_ALL:
        return f * slot
    return (f + 1) * slot - frame_len(n, f - 3)


@dataclass
class FramesBatch:
    """Result of a batched frame translation in the slot layout."""

    data: np.ndarray  # uint8, slot layout
    out_offsets: np.ndarray  # uint64 [n_seq+1]
    lengths: np.ndarray  # uint64 [n_seq]
    frames: int

    def frame(self, i: int, f: int) -> bytes:
        n = int(self.lengths[i])
        ln = frame_len(n, f % 3)
        base = int(self.out_offsets[i]) + slot_frame_offset(n, self.frames, f)
        return self.data[base:base + ln].tobytes()

    def sequence_frames(self, i: int) -> List[bytes]:
        """The frames the reference returns for sequence i (empty frames omitted)."""
        n_slots = 1 if self.frames == N.FRAMES_ONE else self.frames
        out = [self.frame(i, f) for f in range(n_slots)]
        return [x for x in out if len(x)]


def translate_frames_batch(
    reads: Union[np.ndarray, Sequence[bytes], Tuple[np.ndarray, np.ndarray]],
    table: int = 1,
    strict: bool = False,
    frames: int = N.FRAMES_ALL,
    enc: int = N.ENC_ASCII,
    ctx: Optional[N.Context] = None,
    offsets=None,
    devices=None,
) -> FramesBatch:
    """Translate every read of a batch.

    reads: a 2-D uint8 array (n_reads x read_len, uniform batch), a sequence of bytes-like reads, or -- with
    `offsets=` -- one flat uint8 buffer holding all reads back to back (CSR, offsets uint64 [n+1]).  A 2-tuple
    (flat ndarray, integer ndarray) is still taken as the CSR pair; a tuple of reads is a sequence of reads.

    devices: a list of device indices (or of Contexts): the batch is partitioned by bytes over them
    (qd_translate_frames_batch_multi: one host thread and one set of streams per device, one output buffer).
    """
    L = N.lib()
    err = N.QdErr()
    ctxs = None
    if devices is not None:
        ctxs = [d if isinstance(d, N.Context) else N.default_context(int(d)) for d in devices]
        ctx = ctxs[0]
    ctx = ctx or N.default_context()
    if isinstance(reads, np.ndarray) and reads.ndim == 2:
        a = _u8(reads)
        n_seq, seq_len = a.shape
        per = 16 * runs(seq_len) * frames
        out = np.empty(n_seq * per, dtype=np.uint8)
        if ctxs is not None:
            rc = L.qd_translate_frames_batch_multi(N.context_array(ctxs), len(ctxs), table, enc, int(strict), frames, _p(a), None, n_seq, seq_len,
                                                   _p(out), None, C.byref(err))
        else:
            rc = L.qd_translate_frames_batch_uniform(ctx.handle, table, enc, int(strict), frames, _p(a), n_seq, seq_len, _p(out),
                                                     C.byref(err))
        ctx.check(rc, err)
        return FramesBatch(out, np.arange(n_seq + 1, dtype=np.uint64) * np.uint64(per),
                           np.full(n_seq, seq_len, dtype=np.uint64), frames)
    def _is_offsets(o) -> bool:
        return isinstance(o, np.ndarray) and o.ndim == 1 and np.issubdtype(o.dtype, np.integer) and o.dtype != np.uint8

    if offsets is None and isinstance(reads, tuple) and len(reads) == 2 and isinstance(reads[0], np.ndarray) and _is_offsets(reads[1]):
        reads, offsets = reads
    if offsets is not None:
        flat, offsets = _u8(reads), np.ascontiguousarray(offsets, dtype=np.uint64)
        if flat.ndim != 1 or offsets.ndim != 1 or offsets.size < 1 or int(offsets[0]) != 0 or int(offsets[-1]) != flat.size \
                or (offsets.size > 1 and bool((np.diff(offsets.astype(np.int64)) < 0).any())):
            raise ValueError("offsets must be a non-decreasing uint64 array [n_reads + 1] with offsets[0] == 0 and offsets[-1] == len(flat)")
    else:
        lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
        offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offsets[1:])
        flat = _u8(b"".join(bytes(r) for r in reads))
    n_seq = len(offsets) - 1
    total = int(L.qd_batch_out_bytes(_p(offsets), n_seq, frames))
    out = np.empty(max(total, 1), dtype=np.uint8)
    out_offsets = np.empty(n_seq + 1, dtype=np.uint64)
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.uint8)
    if ctxs is not None:
        rc = L.qd_translate_frames_batch_multi(N.context_array(ctxs), len(ctxs), table, enc, int(strict), frames, _p(flat), _p(offsets), n_seq, 0,
                                               _p(out), _p(out_offsets), C.byref(err))
    else:
        rc = L.qd_translate_frames_batch(ctx.handle, table, enc, int(strict), frames, _p(flat), _p(offsets), n_seq, _p(out),
                                         _p(out_offsets), C.byref(err))
    ctx.check(rc, err)
    return FramesBatch(out[:total], out_offsets, np.diff(offsets), frames)


def translate_frames(dna, table: int = 1, strict: bool = False, frames: int = N.FRAMES_ALL, enc: int = N.ENC_ASCII,
                     ctx: Optional[N.Context] = None, devices=None) -> List[bytes]:
    """The frames of ONE sequence (qd_translate_frames; with `devices`, qd_translate_frames_multi: the sequence is cut
    at multiples of 48 with a 2-nt halo and spread over the devices).  Returns the frames the reference returns."""
    L = N.lib()
    a = _u8(dna)
    n = a.size
    n_slots = 1 if frames == N.FRAMES_ONE else frames
    out = np.empty(max(1, sum(frame_len(n, f % 3) for f in range(n_slots))), dtype=np.uint8)
    lens = (C.c_size_t * 6)()
    nf = C.c_int()
    err = N.QdErr()
    src = _p(a if n else np.zeros(1, np.uint8))
    if devices is not None:
        ctxs = [d if isinstance(d, N.Context) else N.default_context(int(d)) for d in devices]
        ctx = ctxs[0]
        rc = L.qd_translate_frames_multi(N.context_array(ctxs), len(ctxs), table, enc, int(strict), frames, src, n, _p(out), lens, C.byref(nf),
                                         C.byref(err))
    else:
        ctx = ctx or N.default_context()
        rc = L.qd_translate_frames(ctx.handle, table, enc, int(strict), frames, src, n, _p(out), lens, C.byref(nf), C.byref(err))
    ctx.check(rc, err)
    res, off = [], 0
    for f in range(6):
        if lens[f]:
            res.append(out[off:off + lens[f]].tobytes())
        off += lens[f]
    assert len(res) == nf.value
    return res


def revcomp(dna, strict: bool = False, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None, devices=None, out=None) -> np.ndarray:
    """Reverse complement of ONE sequence (qd_revcomp; with `devices`, qd_revcomp_multi over the devices)."""
    L = N.lib()
    a = _u8(dna)
    if out is None:
        out = np.empty(max(a.size, 1), dtype=np.uint8)
    err = N.QdErr()
    src = _p(a if a.size else np.zeros(1, np.uint8))
    if devices is not None:
        ctxs = [d if isinstance(d, N.Context) else N.default_context(int(d)) for d in devices]
        ctx = ctxs[0]
        rc = L.qd_revcomp_multi(N.context_array(ctxs), len(ctxs), enc, int(strict), src, a.size, _p(out), C.byref(err))
    else:
        ctx = ctx or N.default_context()
        rc = L.qd_revcomp(ctx.handle, enc, int(strict), src, a.size, _p(out), C.byref(err))
    ctx.check(rc, err)
    return out[: a.size]


def parse(ascii_seq, strict: bool = False, ctx: Optional[N.Context] = None) -> np.ndarray:
    """ASCII -> masks, dropping ' ' and '\\t' (TryFrom<&[u8]> for DnaSequence<T>, src/rust_api.rs:310-322)."""
    ctx = ctx or N.default_context()
    a = _u8(ascii_seq)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    k = C.c_size_t()
    err = N.QdErr()
    rc = N.lib().qd_parse(ctx.handle, int(strict), _p(a if a.size else np.zeros(1, np.uint8)), a.size, _p(out), C.byref(k), C.byref(err))
    ctx.check(rc, err)
    return out[: k.value]


class FastaRecordC(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("header_begin", "header_end", "content_begin", "content_end", "line_begin", "line_end")]


class FastaScan:
    """Result of fasta_scan: all records' nucleotides as masks back to back + CSR offsets (what the batch entry points
    take), and per record its header text and line range (FastaRecord of src/fasta.rs:17-29)."""

    def __init__(self, text: bytes, masks: np.ndarray, records):
        self.masks = masks
        self.offsets = np.array([r.content_begin for r in records] + [records[-1].content_end if records else 0], dtype=np.uint64)
        self.line_ranges = [(int(r.line_begin), int(r.line_end)) for r in records]
        self.headers = [self._header(text, r) for r in records]

    @staticmethod
    def _header(text: bytes, r) -> str:
        if r.header_begin == 2 ** 64 - 1:
            return ""
        lines = text[r.header_begin:r.header_end].split(b"\n")
        lines = [ln[:-1] if i + 1 < len(lines) and ln.endswith(b"\r") else ln for i, ln in enumerate(lines)]
        return "\n".join(ln[1:].decode("utf-8", "replace") for ln in lines)

    def __len__(self) -> int:
        return len(self.headers)

    def record(self, i: int):
        """(header, masks, line_range) of record i."""
        return self.headers[i], self.masks[int(self.offsets[i]):int(self.offsets[i + 1])], self.line_ranges[i]


class FastaError(N.TranslationError):
    """Located<FastaParseError::ParseError(TranslationError)> (src/fasta.rs:413-423): `.line` is the 1-based line."""

    def located(self) -> str:
        return f"on line {self.line}: error parsing record: {self}"


def fasta_scan(text, strict: bool = False, concatenate_headers: bool = True, allow_preceding_comment: bool = False,
               ctx: Optional[N.Context] = None) -> FastaScan:
    """FastaParser::<DnaSequence<T>>::parse_str (src/fasta.rs:425-480) on the GPU."""
    ctx = ctx or N.default_context()
    raw = text.encode() if isinstance(text, str) else bytes(text)
    a = _u8(raw)
    L = N.lib()
    n_masks, n_records = C.c_size_t(), C.c_size_t()
    err = N.QdErr()
    args = (ctx.handle, int(strict), int(concatenate_headers), int(allow_preceding_comment), _p(a if a.size else np.zeros(1, np.uint8)), a.size)

    def call(masks, recs, cap):
        rc = L.qd_fasta_scan(*args, None if masks is None else _p(masks), C.byref(n_masks), recs, cap, C.byref(n_records), C.byref(err))
        if rc in (1, 2, 3):
            e = N.make_error(err)
            fe = FastaError(str(e), e.kind, e.byte, e.pos, 0)
            fe.line = int(err.seq_index)
            raise fe
        return rc

    # one call in the usual case: there are at most n masks, and a record of ordinary FASTA is longer than 64 bytes; only a
    # file of very short records needs the exact count (the failed call has reported it) and a second call
    cap = a.size // 64 + 16
    masks = np.empty(max(a.size, 1), dtype=np.uint8)
    rec_bytes = np.empty(cap * C.sizeof(FastaRecordC), dtype=np.uint8)  # untouched pages cost nothing
    rc = call(masks, C.c_void_p(rec_bytes.ctypes.data), cap)
    if rc == N.QD_ERR_INVALID_ARGUMENT and n_records.value > cap:
        cap = n_records.value
        rec_bytes = np.empty(cap * C.sizeof(FastaRecordC), dtype=np.uint8)
        rc = call(masks, C.c_void_p(rec_bytes.ctypes.data), cap)
    ctx.check(rc, err)
    recs = (FastaRecordC * max(n_records.value, 1)).from_buffer(rec_bytes)
    return FastaScan(raw, masks[: n_masks.value], list(recs[: n_records.value]))


def fasta_translate_frames(text, table: int = 1, frames: int = N.FRAMES_ALL, strict: bool = False, concatenate_headers: bool = True,
                           allow_preceding_comment: bool = False, ctx: Optional[N.Context] = None):
    """FASTA text in, the frames of every record out, in ONE call (qd_fasta_translate_frames): the scanner's masks and CSR
    offsets stay on the device and feed the frame kernels.  Returns (FastaScan without masks, FramesBatch): headers and
    line ranges of the records, and their frames in the slot layout -- `frames.sequence_frames(i)` are the frames the
    reference returns for record i (FastaParser::parse + translate_all_frames per record)."""
    ctx = ctx or N.default_context()
    raw = text.encode() if isinstance(text, str) else bytes(text)
    a = _u8(raw)
    L = N.lib()
    n_records, out_bytes = C.c_size_t(), C.c_size_t()
    err = N.QdErr()
    head = (ctx.handle, table, int(strict), frames, int(concatenate_headers), int(allow_preceding_comment), _p(a if a.size else np.zeros(1, np.uint8)), a.size)

    def call(recs, cap, out, out_cap, offs):
        rc = L.qd_fasta_translate_frames(*head, recs, cap, C.byref(n_records), None if out is None else _p(out), out_cap,
                                         None if offs is None else _p(offs), C.byref(out_bytes), C.byref(err))
        if rc in (1, 2, 3):
            e = N.make_error(err)
            fe = FastaError(str(e), e.kind, e.byte, e.pos, 0)
            fe.line = int(err.seq_index)
            raise fe
        return rc

    # one call in the usual case: generous guesses for both capacities; the failed call reports what is needed
    n_slots = 1 if frames == N.FRAMES_ONE else frames
    cap = a.size // 64 + 16
    out_cap = 16 * n_slots * (a.size // 48 + cap)
    out = np.empty(max(out_cap, 1), dtype=np.uint8)  # untouched pages cost nothing
    rec_bytes = np.empty(cap * C.sizeof(FastaRecordC), dtype=np.uint8)
    offs = np.empty(cap + 1, dtype=np.uint64)
    rc = call(C.c_void_p(rec_bytes.ctypes.data), cap, out, out_cap, offs)
    if rc == N.QD_ERR_INVALID_ARGUMENT and (n_records.value > cap or out_bytes.value > out_cap):
        cap, out_cap = max(cap, n_records.value), max(out_cap, out_bytes.value)
        out = np.empty(max(out_cap, 1), dtype=np.uint8)
        rec_bytes = np.empty(cap * C.sizeof(FastaRecordC), dtype=np.uint8)
        offs = np.empty(cap + 1, dtype=np.uint64)
        rc = call(C.c_void_p(rec_bytes.ctypes.data), cap, out, out_cap, offs)
    ctx.check(rc, err, table)
    k = n_records.value
    recs = list((FastaRecordC * max(k, 1)).from_buffer(rec_bytes)[:k])
    scan = FastaScan(raw, None, recs)
    lengths = np.array([r.content_end - r.content_begin for r in recs], dtype=np.uint64)
    out_offsets = offs[: k + 1].copy() if k else np.zeros(1, dtype=np.uint64)
    return scan, FramesBatch(out[: out_bytes.value], out_offsets, lengths, frames)


def canonical(dna, forward_only: bool = False, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None) -> np.ndarray:
    """DnaSequence::<Nucleotide>::canonical of one sequence of any length (src/rust_api.rs:374-377)."""
    ctx = ctx or N.default_context()
    a = _u8(dna)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    err = N.QdErr()
    rc = N.lib().qd_canonical(ctx.handle, enc, int(forward_only), _p(a if a.size else np.zeros(1, np.uint8)), a.size, _p(out), C.byref(err))
    ctx.check(rc, err)
    return out[: a.size]


def canonical_windows(dna, w: int = 42, forward_only: bool = False, enc: int = N.ENC_ASCII,
                      ctx: Optional[N.Context] = None) -> np.ndarray:
    """Every length-w window of a strict sequence, canonicalised (src/canonical.rs; benches/canonical.rs:29-34)."""
    ctx = ctx or N.default_context()
    a = _u8(dna)
    count = max(0, a.size - w + 1)
    out = np.empty((count, w), dtype=np.uint8)
    err = N.QdErr()
    rc = N.lib().qd_canonical_windows(ctx.handle, enc, int(forward_only), _p(a if a.size else np.zeros(1, np.uint8)), a.size, w,
                                      _p(out if count else np.zeros(1, np.uint8)), C.byref(err))
    ctx.check(rc, err)
    return out


def canonical_window_keys(dna, w: int = 42, kind: int = N.CANON_KEY_FP64, forward_only: bool = False, enc: int = N.ENC_ASCII,
                          ctx: Optional[N.Context] = None) -> np.ndarray:
    """One fixed-width key per canonical window of a strict sequence: kind CANON_KEY_FP64 ->
    uint64 [count]; CANON_KEY_PACKED -> uint32 [count, 2 * ceil(w / 32)], the canonical window as two bit planes (low bits of the codes, then high bits)."""
    ctx = ctx or N.default_context()
    a = _u8(dna)
    count = max(0, a.size - w + 1)
    words = N.lib().qd_canonical_key_bytes(kind, w) // 4
    if words == 0:
        raise ValueError("bad key kind or window length")
    out = np.zeros(max(count, 1) * words, dtype=np.uint32)
    err = N.QdErr()
    rc = N.lib().qd_canonical_window_keys(ctx.handle, enc, int(forward_only), kind, _p(a if a.size else np.zeros(1, np.uint8)), a.size, w,
                                          _p(out), C.byref(err))
    ctx.check(rc, err)
    out = out[: count * words]
    return out.view(np.uint64) if kind == N.CANON_KEY_FP64 else out.reshape(count, words)


def expansion_windows(windows, max_expansions: int = 16, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    """Expansions of a batch of equal-length ambiguous windows (src/expansions.rs; benches/expansions.rs:126-130).

    Returns (counts, out_index, rows): counts[i] = size_hint (2**64-1 on overflow), out_index[i] = first row of
    window i or 2**64-1 if it was skipped (more than max_expansions), rows = uint8 [total, w].
    """
    ctx = ctx or N.default_context()
    a = _u8(windows)
    assert a.ndim == 2
    n_win, w = a.shape
    counts = np.empty(n_win, dtype=np.uint64)
    out_index = np.empty(n_win + 1, dtype=np.uint64)
    err = N.QdErr()
    L = N.lib()
    if 0 < n_win and w <= 64 and max_expansions < 128 and n_win * max_expansions * w <= (1 << 19):
        # small batches: one call with room for the cap (the library runs its one-pass kernel and copies back once)
        room = np.empty((n_win * max_expansions, w), dtype=np.uint8)
        rc = L.qd_expansion_windows(ctx.handle, enc, _p(a), n_win, w, max_expansions, _p(counts), _p(out_index), _p(room), C.byref(err))
        ctx.check(rc, err)
        return counts, out_index, room[: int(out_index[n_win])]
    rc = L.qd_expansion_windows(ctx.handle, enc, _p(a), n_win, w, max_expansions, _p(counts), _p(out_index), None, C.byref(err))
    ctx.check(rc, err)
    total = int(out_index[n_win])
    rows = np.empty((total, w), dtype=np.uint8)
    if total:
        rc = L.qd_expansion_windows(ctx.handle, enc, _p(a), n_win, w, max_expansions, _p(counts), _p(out_index), _p(rows),
                                    C.byref(err))
        ctx.check(rc, err)
    return counts, out_index, rows


def windows(seq, w: int, ctx: Optional[N.Context] = None) -> np.ndarray:
    """DnaSequence::windows / ProteinSequence::windows (src/rust_api.rs:100-104, 277-279)."""
    ctx = ctx or N.default_context()
    a = _u8(seq)
    count = max(0, a.size - w + 1)
    out = np.empty((count, w), dtype=np.uint8)
    n_windows = C.c_size_t()
    rc = N.lib().qd_windows(ctx.handle, _p(a if a.size else np.zeros(1, np.uint8)), a.size, w,
                            _p(out if count else np.zeros(1, np.uint8)), C.byref(n_windows))
    ctx.check(rc)
    assert n_windows.value == count
    return out


def windows_all(dna, w_dna: int = 42, w_aa: int = 20, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    """benches/all_windows.rs SequenceWindows::from_dna + enumerate for one strict sequence.

    Returns (dna_windows [fwd then rc], aa_windows [returned frames in order], origin, counts[8]).
    """
    ctx = ctx or N.default_context()
    a = _u8(dna)
    n = a.size

    def nwin(length, w):
        return length + 1 - min(w, length + 1)

    n_dna = nwin(n, w_dna) if n else 0
    lens = [frame_len(n, f % 3) for f in range(6)]
    n_aa = sum(nwin(ln, w_aa) for ln in lens if ln)
    dna_out = np.empty((2 * n_dna, w_dna), dtype=np.uint8)
    aa_out = np.empty((n_aa, w_aa), dtype=np.uint8)
    origin = np.empty(2 * n_dna + n_aa, dtype=np.uint64)
    counts = (C.c_size_t * 8)()
    err = N.QdErr()
    dummy = np.zeros(1, np.uint8)
    rc = N.lib().qd_windows_all(ctx.handle, enc, _p(a if n else dummy), n, w_dna, w_aa, _p(dna_out if dna_out.size else dummy),
                                _p(aa_out if aa_out.size else dummy), _p(origin if origin.size else np.zeros(1, np.uint64)),
                                counts, C.byref(err))
    ctx.check(rc, err)
    return dna_out, aa_out, origin, list(counts)


# ----------------------------------------------------------------------------------------------
# device-resident (torch tensors on the context's GPU); asynchronous on torch's current stream
# ----------------------------------------------------------------------------------------------
def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def revcomp_dev(dna, out=None, strict: bool = False, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    if out is None:
        out = torch.empty_like(dna)
    rc = N.lib().qd_revcomp_dev(ctx.handle, enc, int(strict), C.c_void_p(dna.data_ptr()), dna.numel(),
                                C.c_void_p(out.data_ptr()), _stream_ptr(torch))
    ctx.check(rc)
    return out


def translate_frames_uniform_dev(dna, n_seq: int, seq_len: int, out=None, table: int = 1, strict: bool = False,
                                 frames: int = N.FRAMES_ALL, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    if out is None:
        out = torch.empty(n_seq * 16 * runs(seq_len) * frames, dtype=torch.uint8, device=dna.device)
    rc = N.lib().qd_translate_frames_batch_uniform_dev(ctx.handle, table, enc, int(strict), frames, C.c_void_p(dna.data_ptr()),
                                                       n_seq, seq_len, C.c_void_p(out.data_ptr()), _stream_ptr(torch))
    ctx.check(rc, table=table)
    return out


def translate_frames_csr_dev(dna, offsets, out=None, table: int = 1, strict: bool = False, frames: int = N.FRAMES_ALL,
                             enc: int = N.ENC_ASCII, total_runs: Optional[int] = None, ctx: Optional[N.Context] = None):
    """dna: uint8 device tensor; offsets: int64 device tensor [n_seq+1] starting at 0."""
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    n_seq = offsets.numel() - 1
    if out is None:
        if total_runs is None:
            lens = offsets[1:] - offsets[:-1]
            total_runs = int(((lens + 47) // 48).sum().item())
        out = torch.empty(total_runs * 16 * frames, dtype=torch.uint8, device=dna.device)
    rc = N.lib().qd_translate_frames_batch_dev(ctx.handle, table, enc, int(strict), frames, C.c_void_p(dna.data_ptr()),
                                               C.c_void_p(offsets.data_ptr()), n_seq, dna.numel(), C.c_void_p(out.data_ptr()),
                                               _stream_ptr(torch))
    ctx.check(rc, table=table)
    return out


def fasta_scan_dev(text, strict: bool = False, concatenate_headers: bool = True, allow_preceding_comment: bool = False,
                   max_records: Optional[int] = None, ctx: Optional[N.Context] = None):
    """qd_fasta_scan_dev on a torch uint8 CUDA tensor holding FASTA text: returns (masks [n_masks] uint8, records [n_records, 6]
    int64 -- header_begin, header_end, content_begin, content_end, line_begin, line_end --, offsets [n_records + 1] int64),
    all on the device: masks and offsets are what translate_frames_csr_dev(enc=ENC_MASK) takes."""
    import torch

    ctx = ctx or N.default_context(text.device.index or 0)
    n = text.numel()
    cap = max_records if max_records is not None else n // 64 + 16
    masks = torch.empty(max(n, 1), dtype=torch.uint8, device=text.device)
    n_masks, n_records = C.c_size_t(), C.c_size_t()
    err = N.QdErr()
    L = N.lib()

    def call(cap):
        recs = torch.empty((cap, 6), dtype=torch.int64, device=text.device)
        offs = torch.empty(cap + 1, dtype=torch.int64, device=text.device)
        rc = L.qd_fasta_scan_dev(ctx.handle, int(strict), int(concatenate_headers), int(allow_preceding_comment), C.c_void_p(text.data_ptr()), n,
                                 C.c_void_p(masks.data_ptr()), C.byref(n_masks), C.c_void_p(recs.data_ptr()), cap, C.byref(n_records),
                                 C.c_void_p(offs.data_ptr()), C.byref(err), _stream_ptr(torch))
        return rc, recs, offs

    rc, recs, offs = call(cap)
    if rc == N.QD_ERR_INVALID_ARGUMENT and n_records.value > cap:
        rc, recs, offs = call(n_records.value)
    if rc in (1, 2, 3):
        e = N.make_error(err)
        fe = FastaError(str(e), e.kind, e.byte, e.pos, 0)
        fe.line = int(err.seq_index)
        raise fe
    ctx.check(rc, err)
    k = n_records.value
    return masks[: n_masks.value], recs[:k], offs[: k + 1]


def parse_dev(ascii_seq, out=None, strict: bool = False, ctx: Optional[N.Context] = None):
    """qd_parse_dev on a torch uint8 CUDA tensor: returns (masks buffer, device int64[1] count)."""
    import torch

    ctx = ctx or N.default_context(ascii_seq.device.index or 0)
    n = ascii_seq.numel()
    if out is None:
        out = torch.empty(n + 16, dtype=torch.uint8, device=ascii_seq.device)
    count = torch.empty(1, dtype=torch.int64, device=ascii_seq.device)  # always written by the call (no fill kernel needed)
    rc = N.lib().qd_parse_dev(ctx.handle, int(strict), C.c_void_p(ascii_seq.data_ptr()), n, C.c_void_p(out.data_ptr()),
                              C.c_void_p(count.data_ptr()), _stream_ptr(torch))
    ctx.check(rc)
    return out, count


def canonical_windows_batch_dev(dna, n_seq: int, seq_len: int, w: int = 42, out=None, forward_only: bool = False,
                                enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    """Every length-w window of each of n_seq sequences of seq_len nucleotides (torch uint8 CUDA tensor, back to back);
    returns a uint8 tensor [n_seq * (seq_len - w + 1), w] (benches/canonical.rs shape, batched)."""
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    nw = max(0, seq_len - w + 1)
    if out is None:
        out = torch.empty(max(n_seq * nw * w, 16), dtype=torch.uint8, device=dna.device)
    rc = N.lib().qd_canonical_windows_batch_dev(ctx.handle, enc, int(forward_only), C.c_void_p(dna.data_ptr()), n_seq, seq_len, w,
                                                C.c_void_p(out.data_ptr()), _stream_ptr(torch))
    ctx.check(rc)
    return out[: n_seq * nw * w].view(n_seq * nw, w) if nw else out[:0].view(0, w)


def canonical_window_keys_dev(dna, n_seq: int, seq_len: int, w: int = 42, kind: int = N.CANON_KEY_FP64, out=None,
                              forward_only: bool = False, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    """qd_canonical_window_keys_dev on a torch uint8 CUDA tensor of n_seq x seq_len nucleotides: an int64 tensor
    [n_windows] (CANON_KEY_FP64; the bits are the unsigned key) or an int32 tensor [n_windows, 2 * ceil(w / 32)] (PACKED: bit planes)."""
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    nw = max(0, seq_len - w + 1)
    words = N.lib().qd_canonical_key_bytes(kind, w) // 4
    if words == 0:
        raise ValueError("bad key kind or window length")
    if out is None:
        out = torch.empty(max(n_seq * nw * words, 4), dtype=torch.int32, device=dna.device)
    rc = N.lib().qd_canonical_window_keys_dev(ctx.handle, enc, int(forward_only), kind, C.c_void_p(dna.data_ptr()), n_seq, seq_len, w,
                                              C.c_void_p(out.data_ptr()), _stream_ptr(torch))
    ctx.check(rc)
    flat = out.view(torch.int32)[: n_seq * nw * words]
    return flat.view(torch.int64) if kind == N.CANON_KEY_FP64 else flat.view(n_seq * nw, words)


def canonical_window_keys_csr_dev(dna, offsets, w: int = 42, kind: int = N.CANON_KEY_FP64, forward_only: bool = False,
                                  enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    """One key per canonical window of a VARIABLE-LENGTH batch (the masks + CSR offsets of fasta_scan_dev, say), by flat
    position: the batch is run as ONE sequence, so the key of window j of sequence i sits at index offsets[i] + j, and the
    w - 1 positions in front of every sequence end hold the keys of windows that span two sequences -- to be skipped.
    Returns (keys, valid): keys as canonical_window_keys_dev over total - w + 1 positions, valid a bool tensor of the same
    length (position p is valid iff p + w <= the end of its sequence)."""
    import torch

    total = dna.numel()
    keys = canonical_window_keys_dev(dna, 1, total, w, kind=kind, forward_only=forward_only, enc=enc, ctx=ctx)
    n_pos = max(0, total - w + 1)
    offs = offsets.to(torch.int64)
    # +1 at a sequence's start, -1 at the first start that no longer fits it: valid = the running sum is 1
    delta = torch.zeros(n_pos + w + 1, dtype=torch.int32, device=dna.device)
    starts, ends = offs[:-1], offs[1:]
    fits = (ends - starts) >= w
    delta.index_add_(0, starts[fits], torch.ones(int(fits.sum()), dtype=torch.int32, device=dna.device))
    delta.index_add_(0, (ends - w + 1)[fits], -torch.ones(int(fits.sum()), dtype=torch.int32, device=dna.device))
    valid = torch.cumsum(delta, 0)[:n_pos] > 0
    return keys, valid


def expansion_windows_dev(windows, max_expansions: int = 16, out=None, enc: int = N.ENC_ASCII, ctx: Optional[N.Context] = None):
    """qd_expansion_windows_dev on a torch uint8 CUDA tensor [n_win, w]: returns (counts, out_index, rows buffer); rows are
    written while they fit `out` (default: the worst case n_win * max_expansions rows)."""
    import torch

    ctx = ctx or N.default_context(windows.device.index or 0)
    n_win, w = windows.shape
    counts = torch.empty(n_win, dtype=torch.int64, device=windows.device)
    out_index = torch.empty(n_win + 1, dtype=torch.int64, device=windows.device)
    if out is None:
        out = torch.empty(max(n_win * max_expansions * w, 16), dtype=torch.uint8, device=windows.device)
    rc = N.lib().qd_expansion_windows_dev(ctx.handle, enc, C.c_void_p(windows.data_ptr()), n_win, w, max_expansions,
                                          C.c_void_p(counts.data_ptr()), C.c_void_p(out_index.data_ptr()), C.c_void_p(out.data_ptr()),
                                          out.numel() // w, _stream_ptr(torch))
    ctx.check(rc)
    return counts, out_index, out


def windows_batch_dev(seq, n_seq: int, seq_len: int, w: int, mode: int = 0, out=None, ctx: Optional[N.Context] = None):
    """qd_windows_batch_dev: every length-w window of each of n_seq sequences (torch uint8 CUDA tensor) -> [n_seq * count, w].
    mode: 0 copy, 1 strict DNA letters upper-cased, 2 windows of the reverse complement, 3 / 4 the same from masks."""
    import torch

    ctx = ctx or N.default_context(seq.device.index or 0)
    nw = max(0, seq_len - w + 1)
    if out is None:
        out = torch.empty(max(n_seq * nw * w, 16), dtype=torch.uint8, device=seq.device)
    rc = N.lib().qd_windows_batch_dev(ctx.handle, C.c_void_p(seq.data_ptr()), n_seq, seq_len, w, mode, C.c_void_p(out.data_ptr()),
                                      _stream_ptr(torch))
    ctx.check(rc)
    return out[: n_seq * nw * w].view(n_seq * nw, w) if nw else out[:0].view(0, w)


def windows_all_batch_dev(dna, n_seq: int, seq_len: int, w_dna: int = 42, w_aa: int = 20, enc: int = N.ENC_ASCII, dna_out=None,
                          aa_out=None, ctx: Optional[N.Context] = None):
    """benches/all_windows.rs SequenceWindows::from_dna for a batch (table Ncbi1): returns (counts[8], dna windows
    [forward block, reverse-complement block], aa windows [six frame blocks])."""
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    counts = (C.c_size_t * 8)()
    N.lib().qd_windows_all_batch_counts(seq_len, w_dna, w_aa, counts)
    cs = list(counts)
    if dna_out is None:
        dna_out = torch.empty(max(2 * n_seq * cs[0] * w_dna, 16), dtype=torch.uint8, device=dna.device)
    if aa_out is None:
        aa_out = torch.empty(max(n_seq * sum(cs[2:]) * w_aa, 16), dtype=torch.uint8, device=dna.device)
    rc = N.lib().qd_windows_all_batch_dev(ctx.handle, enc, C.c_void_p(dna.data_ptr()), n_seq, seq_len, w_dna, w_aa,
                                          C.c_void_p(dna_out.data_ptr()), C.c_void_p(aa_out.data_ptr()), _stream_ptr(torch))
    ctx.check(rc)
    return cs, dna_out, aa_out


def checksum64_dev(buf, base: int = 0, acc=None, ctx: Optional[N.Context] = None):
    """qd_checksum64_dev: adds the position-weighted linear checksum of a uint8 CUDA tensor (bytes [base, base + n) of a
    stream) to `acc` (int64 device tensor [1], created zeroed when None) and returns it; int(acc.item()) & (2**64 - 1)
    is the checksum."""
    import torch

    ctx = ctx or N.default_context(buf.device.index or 0)
    if acc is None:
        acc = torch.zeros(1, dtype=torch.int64, device=buf.device)
    rc = N.lib().qd_checksum64_dev(ctx.handle, C.c_void_p(buf.data_ptr()), buf.numel(), base, C.c_void_p(acc.data_ptr()), _stream_ptr(torch))
    ctx.check(rc)
    return acc


def u64(t) -> int:
    """The unsigned value of a one-element int64 tensor."""
    return int(t.item()) & 0xFFFFFFFFFFFFFFFF


def _consumer(consume):
    if consume is None:
        return None, None
    cb = N.CHUNK_FN(lambda user, chunk, first, n_units, slot, nbytes, n_rows_dev, stream: consume(chunk, first, n_units, slot, nbytes, n_rows_dev, stream))
    return cb, cb


def canonical_windows_chunked_dev(dna, n_seq: int, seq_len: int, w: int, ring, chunk_seqs: int, checksums: bool = True,
                                  forward_only: bool = False, enc: int = N.ENC_ASCII, consume=None, ctx: Optional[N.Context] = None):
    """qd_canonical_windows_chunked_dev: canonical windows of a batch whose output does not fit HBM, chunk by chunk through
    the uint8 CUDA tensor `ring`.  Returns the per-chunk checksums (int64 device tensor [n_chunks]; their sum mod 2**64 is
    the checksum of the whole output) or None.  consume(chunk, first_seq, n_seq, slot_ptr, bytes, n_rows_dev, stream) is
    called on the host after each chunk has been enqueued."""
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    L = N.lib()
    n_chunks = L.qd_chunk_count(n_seq, chunk_seqs)
    sums = torch.zeros(max(n_chunks, 1), dtype=torch.int64, device=dna.device) if checksums else None
    cb, keep = _consumer(consume)
    rc = L.qd_canonical_windows_chunked_dev(ctx.handle, enc, int(forward_only), C.c_void_p(dna.data_ptr()), n_seq, seq_len, w,
                                            C.c_void_p(ring.data_ptr()), ring.numel(), chunk_seqs,
                                            None if sums is None else C.c_void_p(sums.data_ptr()), C.cast(cb, C.c_void_p) if cb else None, None,
                                            _stream_ptr(torch))
    ctx.check(rc)
    del keep
    return None if sums is None else sums[:n_chunks]


def windows_all_chunked_dev(dna, n_seq: int, seq_len: int, ring, chunk_seqs: int, w_dna: int = 42, w_aa: int = 20, checksums: bool = True,
                            enc: int = N.ENC_ASCII, consume=None, ctx: Optional[N.Context] = None):
    """qd_windows_all_chunked_dev: the all_windows shape of a batch, chunk by chunk through `ring`; returns the
    chunk-relative per-chunk checksums or None."""
    import torch

    ctx = ctx or N.default_context(dna.device.index or 0)
    L = N.lib()
    n_chunks = L.qd_chunk_count(n_seq, chunk_seqs)
    sums = torch.zeros(max(n_chunks, 1), dtype=torch.int64, device=dna.device) if checksums else None
    cb, keep = _consumer(consume)
    rc = L.qd_windows_all_chunked_dev(ctx.handle, enc, C.c_void_p(dna.data_ptr()), n_seq, seq_len, w_dna, w_aa, C.c_void_p(ring.data_ptr()),
                                      ring.numel(), chunk_seqs, None if sums is None else C.c_void_p(sums.data_ptr()),
                                      C.cast(cb, C.c_void_p) if cb else None, None, _stream_ptr(torch))
    ctx.check(rc)
    del keep
    return None if sums is None else sums[:n_chunks]


def expansion_windows_chunked_dev(windows, ring, chunk_windows: int, max_expansions: int = 16, checksums: bool = True, counts=None,
                                  enc: int = N.ENC_ASCII, consume=None, ctx: Optional[N.Context] = None):
    """qd_expansion_windows_chunked_dev on a [n_win, w] uint8 CUDA tensor: returns (rows per chunk, chunk-relative
    checksums or None), both int64 device tensors [n_chunks]."""
    import torch

    ctx = ctx or N.default_context(windows.device.index or 0)
    L = N.lib()
    n_win, w = windows.shape
    n_chunks = L.qd_chunk_count(n_win, chunk_windows)
    rows = torch.zeros(max(n_chunks, 1), dtype=torch.int64, device=windows.device)
    sums = torch.zeros(max(n_chunks, 1), dtype=torch.int64, device=windows.device) if checksums else None
    cb, keep = _consumer(consume)
    rc = L.qd_expansion_windows_chunked_dev(ctx.handle, enc, C.c_void_p(windows.data_ptr()), n_win, w, max_expansions,
                                            C.c_void_p(ring.data_ptr()), ring.numel(), chunk_windows,
                                            None if counts is None else C.c_void_p(counts.data_ptr()), C.c_void_p(rows.data_ptr()),
                                            None if sums is None else C.c_void_p(sums.data_ptr()), C.cast(cb, C.c_void_p) if cb else None,
                                            None, _stream_ptr(torch))
    ctx.check(rc)
    del keep
    return rows[:n_chunks], (None if sums is None else sums[:n_chunks])


def dev_error_reset(ctx: Optional[N.Context] = None):
    import torch

    ctx = ctx or N.default_context()
    ctx.check(N.lib().qd_dev_error_reset(ctx.handle, _stream_ptr(torch)))


def dev_error_fetch(dna, offsets=None, n_seq: int = 0, uniform_len: int = 0, strict: bool = False, enc: int = N.ENC_ASCII,
                    ctx: Optional[N.Context] = None) -> Optional[N.TranslationError]:
    """Synchronises; returns None when the launches since dev_error_reset saw only valid input."""
    import torch

    ctx = ctx or N.default_context()
    err = N.QdErr()
    rc = N.lib().qd_dev_error_fetch(ctx.handle, _stream_ptr(torch), enc, int(strict), C.c_void_p(dna.data_ptr()),
                                    None if offsets is None else C.c_void_p(offsets.data_ptr()), n_seq, uniform_len, C.byref(err))
    if rc in (N.QD_ERR_CUDA, N.QD_ERR_INVALID_ARGUMENT):
        ctx.check(rc)
    return None if rc == N.QD_OK else N.make_error(err)


class _PinnedOwner:
    """Keeps a qd_host_alloc block alive for the numpy array that views it."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes

    def __del__(self):
        try:
            N.lib().qd_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8) -> np.ndarray:
    """A numpy array in pinned (page-locked) host memory (qd_host_alloc): the host entry points copy from and to it directly,
    without the staging that pageable arrays go through.  The memory is released when the array (and its views) are gone."""
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(1, count * dt.itemsize)
    ptr = N.lib().qd_host_alloc(nbytes)
    if not ptr:
        raise MemoryError(f"qd_host_alloc({nbytes}) failed")
    owner = _PinnedOwner(int(ptr), nbytes)
    buf = (C.c_uint8 * nbytes).from_address(owner.ptr)
    buf._owner = owner  # the ctypes buffer is the array's base: the block lives as long as any view of it
    return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)
