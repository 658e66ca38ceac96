"""qd_fasta_translate_frames: FASTA text in, the frames of every record out, masks never leave the device.
Against the oracle: FastaParser::parse (src/fasta.rs:425-480) then translate_all_frames per record.  Needs a B200: -m gpu."""

import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from quickdna_b200 import batch

    return batch


@pytest.fixture(scope="module")
def N():
    from quickdna_b200 import _native

    return _native


def _fasta(rng, n_rec, crlf=False, lo=0, hi=900, alphabet=b"ACGTacgtNRYn"):
    eol = b"\r\n" if crlf else b"\n"
    parts = []
    for i in range(n_rec):
        parts.append(b">seq%d some description" % i + eol)
        n = int(rng.integers(lo, hi))
        body = bytes(rng.choice(np.frombuffer(alphabet, dtype=np.uint8), n))
        for k in range(0, n, 70):
            parts.append(body[k:k + 70] + eol)
    return b"".join(parts)


def _frames_oracle(masks, frames, table):
    dna = oracle.ascii_of(masks)
    if frames == 1:
        out = oracle.translate_bytes(table, dna)
        return [out] if out else []
    return oracle.py_self_frames(table, dna) if frames == 3 else oracle.py_all_frames(table, dna)


@pytest.mark.parametrize("crlf", [False, True])
@pytest.mark.parametrize("frames,table", [(6, 1), (3, 11), (1, 2)])
def test_fasta_translate_frames_equals_scan_then_translate(B, frames, table, crlf):
    rng = np.random.default_rng(17 + frames)
    text = _fasta(rng, 301, crlf=crlf)
    scan, fb = B.fasta_translate_frames(text, table=table, frames=frames)
    want = oracle.fasta_scan(text)
    assert len(scan) == len(want) >= 290  # (an empty record makes two header lines one record: concatenate_headers)
    for i, (header, masks, lines) in enumerate(want):
        assert scan.headers[i] == header and scan.line_ranges[i] == tuple(lines)
        assert int(fb.lengths[i]) == len(masks)
        assert fb.sequence_frames(i) == _frames_oracle(masks, frames, table), i
    # the same as the two separate calls of this library
    two = B.fasta_scan(text)
    fb2 = B.translate_frames_batch((two.masks, two.offsets), table=table, frames=frames, enc=1)
    assert np.array_equal(fb.data, fb2.data) and np.array_equal(fb.out_offsets, fb2.out_offsets)


def test_fasta_translate_frames_edges(B, N):
    scan, fb = B.fasta_translate_frames(b"")
    assert len(scan) == 0 and fb.data.size == 0
    scan, fb = B.fasta_translate_frames(b">only a header\n")
    assert len(scan) == 1 and fb.sequence_frames(0) == []
    scan, fb = B.fasta_translate_frames(b"ACGTAC\n>x\nAC\n>y\n\n>z\nAAAGGGAAA\n")  # headerless leading record, short and empty records
    assert [fb.sequence_frames(i) for i in range(len(scan))] == [oracle.py_all_frames(1, s) for s in (b"ACGTAC", b"AC", b"", b"AAAGGGAAA")]
    with pytest.raises(ValueError) as ei:
        B.fasta_translate_frames(b">a\nACGT\nAC!T\n")
    assert ei.value.line == 3 and ei.value.kind == 2
    with pytest.raises(ValueError) as ei:
        B.fasta_translate_frames(b">a\nACGN\n", strict=True)
    assert ei.value.kind == 3
    with pytest.raises(ValueError):
        B.fasta_translate_frames(b">a\nACGT\n", table=17)  # table id first
    # the sizing call of the C ABI
    L, ctx = N.lib(), N.default_context()
    text = b">a\nAAAGGGAAA\n>b\nACG\n"
    a = np.frombuffer(text, dtype=np.uint8)
    n_rec, out_bytes, err = C.c_size_t(), C.c_size_t(), N.QdErr()
    rc = L.qd_fasta_translate_frames(ctx.handle, 1, 0, 6, 1, 0, a.ctypes.data_as(C.c_void_p), a.size, None, 0, C.byref(n_rec), None, 0, None,
                                     C.byref(out_bytes), C.byref(err))
    assert rc == 0 and n_rec.value == 2 and out_bytes.value == 2 * 96


def test_fasta_translate_frames_large(B):
    rng = np.random.default_rng(5)
    text = _fasta(rng, 60_000, lo=300, hi=1500, alphabet=b"ACGT")  # ~ 55 MB of text: staged copies, several tiles per record
    scan, fb = B.fasta_translate_frames(text)
    two = B.fasta_scan(text)
    assert len(scan) == len(two) == 60_000
    for i in (0, 1, 29_999, 59_999):
        assert fb.sequence_frames(i) == oracle.py_all_frames(1, oracle.ascii_of(two.record(i)[1]))
    fb2 = B.translate_frames_batch((two.masks, two.offsets), frames=6, enc=1)
    assert np.array_equal(fb.data, fb2.data)


def test_fasta_scan_dev_chains_into_the_frame_kernels(B, N):
    """Text resident on the device -> masks, records, CSR offsets on the device -> frames, nothing crossing PCIe in between."""
    import torch

    rng = np.random.default_rng(23)
    text = _fasta(rng, 2_000, crlf=True, lo=0, hi=3000)
    d = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).cuda()
    masks, recs, offs = B.fasta_scan_dev(d)
    host = B.fasta_scan(text)
    assert np.array_equal(masks.cpu().numpy(), host.masks)
    assert np.array_equal(offs.cpu().numpy().astype(np.uint64), host.offsets)
    r = recs.cpu().numpy()
    assert [(int(a), int(b)) for a, b in r[:, 4:6]] == host.line_ranges
    out = B.translate_frames_csr_dev(masks, offs, table=1, frames=6, enc=N.ENC_MASK)
    want = B.translate_frames_batch((host.masks, host.offsets), frames=6, enc=1)
    assert np.array_equal(out.cpu().numpy(), want.data)
    # too small a record capacity: the call reports the count, the wrapper retries
    masks2, recs2, offs2 = B.fasta_scan_dev(d, max_records=10)
    assert torch.equal(recs2, recs) and torch.equal(offs2, offs)
    bad = bytearray(text)
    bad[len(bad) // 2] = ord("!") if bad[len(bad) // 2] not in b">\r\n" else bad[len(bad) // 2]
    if bytes(bad) != text:
        with pytest.raises(ValueError) as ei:
            B.fasta_scan_dev(torch.from_numpy(np.frombuffer(bytes(bad), dtype=np.uint8).copy()).cuda())
        with pytest.raises(ValueError) as eh:
            B.fasta_scan(bytes(bad))
        assert (ei.value.line, ei.value.pos, ei.value.kind) == (eh.value.line, eh.value.pos, eh.value.kind)


def test_fasta_to_canonical_window_keys_on_the_device(B, N):
    """FASTA text -> masks + offsets -> one packed key per canonical 42-nt window of every record, nothing leaving the device:
    the variable-length batch runs as one sequence, keys are indexed by flat position, seam positions are flagged invalid."""
    import torch

    rng = np.random.default_rng(31)
    text = _fasta(rng, 300, lo=0, hi=400, alphabet=b"ACGTacgt")
    d = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).cuda()
    masks, recs, offs = B.fasta_scan_dev(d, strict=True)
    keys, valid = B.canonical_window_keys_csr_dev(masks, offs, 42, kind=N.CANON_KEY_PACKED, enc=N.ENC_MASK)
    keys, valid, o = keys.cpu().numpy(), valid.cpu().numpy(), offs.cpu().numpy()
    m = masks.cpu().numpy()
    n_valid = 0
    for i in range(len(o) - 1):
        seq = m[o[i]:o[i + 1]]
        nw = max(0, len(seq) - 41)
        assert valid[o[i]:o[i] + nw].all() and not valid[o[i] + nw:min(o[i + 1], len(valid))].any(), i
        n_valid += nw
        if nw and i % 7 == 0:
            alone = B.canonical_window_keys_dev(torch.from_numpy(seq.copy()).cuda(), 1, len(seq), 42, kind=N.CANON_KEY_PACKED, enc=N.ENC_MASK)
            assert np.array_equal(alone.cpu().numpy(), keys[o[i]:o[i] + nw]), i
    assert int(valid.sum()) == n_valid
