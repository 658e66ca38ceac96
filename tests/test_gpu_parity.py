"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: -m gpu.

Bit-exact bar: every output byte, every frame count, every error (kind, byte, position).
"""

import ctypes as C
import hashlib
import itertools

import numpy as np
import pytest

import oracle
from tests import vectors as V

pytestmark = pytest.mark.gpu

IUPAC = b"ACGTacgtWMRYSKBVDHNwmryskbvdhn"


@pytest.fixture(scope="module")
def qd():
    import quickdna_b200

    return quickdna_b200


@pytest.fixture(scope="module")
def N():
    from quickdna_b200 import _native

    return _native


@pytest.fixture(scope="module")
def B():
    from quickdna_b200 import batch

    return batch


def _rand(rng, n, alphabet=b"ACGT"):
    return bytes(np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), n)])


# ------------------------------------------------------------------ golden vectors, Python surface


@pytest.mark.parametrize("table,dna,protein", V.TRANSLATE)
def test_translate_vectors(qd, table, dna, protein):
    assert qd.DnaSequence(dna).translate(table=table) == qd.ProteinSequence(protein)
    if all(c in "ACGTacgt" for c in dna):
        assert qd.DnaSequence(dna).translate(table=table, strict=True) == qd.ProteinSequence(protein)


@pytest.mark.parametrize("dna,frames", V.SELF_FRAMES)
def test_self_frames_vectors(qd, dna, frames):
    want = [qd.ProteinSequence(f) for f in frames]
    assert qd.DnaSequence(dna).translate_self_frames(table=1) == want
    assert qd.DnaSequence(dna).translate_self_frames(table=1, strict=True) == want


@pytest.mark.parametrize("dna,frames", V.ALL_FRAMES)
def test_all_frames_vectors(qd, dna, frames):
    want = [qd.ProteinSequence(f) for f in frames]
    assert qd.DnaSequence(dna).translate_all_frames(table=1) == want
    assert qd.DnaSequence(dna).translate_all_frames(table=1, strict=True) == want


@pytest.mark.parametrize("dna,rc", V.REVCOMP)
def test_revcomp_vectors(qd, dna, rc):
    assert qd.DnaSequence(dna).reverse_complement() == qd.DnaSequence(rc)


def test_wrapper_behaviour_like_reference(qd):
    # tests/test_wrapper.py of the reference, restated
    D, P = qd.DnaSequence, qd.ProteinSequence
    assert D("aaa") == D("aaa") and D("aaa") != P("aaa") and hash(D("aaa")) != hash(P("aaa"))
    d = D("aaa")
    assert d[1] == ord("a") and d[:] == D("aaa") and d[1:] == D("aa") and d[1:] != P("aa")
    assert len(D("a" * 20)) == 20
    assert D("at") + D("cg") == D("atcg") and P("KG") + P("GK") == P("KGGK")
    with pytest.raises(TypeError):
        D("at") + P("KG")
    assert D("a") * 20 == D("a" * 20)
    assert list(D("atcg" * 20)) == list("atcg" * 20)
    with pytest.raises(ValueError):
        D("AAABBBGGG").reverse_complement(strict=True)
    assert D("AAAGGGAAA").translate(table=1, strict=True) == P("KGK")
    with pytest.raises(ValueError):
        D("AAABBBAAA").translate(table=1, strict=True)
    assert repr(D("A" * 40)) == "DnaSequence(seq='" + "A" * 30 + "...')"


@pytest.mark.parametrize("fn,dna,kind,byte,msg", V.ERRORS)
def test_error_vectors(qd, fn, dna, kind, byte, msg):
    from quickdna_b200 import quickdna as native

    raw = dna.encode("latin-1")
    call = {
        "translate": lambda: native._translate(1, raw),
        "translate_strict": lambda: native._translate_strict(1, raw),
        "revcomp": lambda: native._reverse_complement(raw),
        "revcomp_strict": lambda: native._reverse_complement_strict(raw),
    }[fn]
    with pytest.raises(ValueError) as ei:
        call()
    assert (ei.value.kind, ei.value.byte, str(ei.value)) == (kind, byte, msg)
    # and the oracle agrees on the position
    ocall = {
        "translate": lambda: oracle.translate_bytes(1, raw),
        "translate_strict": lambda: oracle.translate_bytes(1, raw, strict=True),
        "revcomp": lambda: oracle.revcomp_bytes(raw),
        "revcomp_strict": lambda: oracle.revcomp_bytes(raw, strict=True),
    }[fn]
    with pytest.raises(oracle.OracleError) as oi:
        ocall()
    assert ei.value.pos == oi.value.pos


def test_table_errors(qd):
    from quickdna_b200 import quickdna as native

    with pytest.raises(ValueError) as ei:
        native._translate(17, b"QQQ")  # table is checked before the DNA
    assert str(ei.value) == V.BAD_TABLE_MESSAGE
    for t in (0, 17, 18, 19, 20, 34, 200):
        with pytest.raises(ValueError):
            native._check_table(t)
        with pytest.raises(ValueError):
            native._translate_strict(t, b"")
    for t in V.ALL_ACCEPTED_TABLES:
        native._check_table(t)


def test_all_frames_short_input_quirk(qd):
    # quickdna/__init__.py:182: len < 3 still validates through the NON-strict reverse complement
    with pytest.raises(ValueError):
        qd.DnaSequence("Q").translate_all_frames()
    assert qd.DnaSequence("NN").translate_all_frames(strict=True) == []
    assert qd.DnaSequence("QQ").translate_self_frames() == []


# ------------------------------------------------------------------ exhaustive / random parity


def test_exhaustive_codons_all_tables(qd):
    # tests/test_exhaustive.py shape, against the oracle (whose LUT is sha-pinned to tables.dat)
    letters = "ATCGWMRYSKBVDHN"
    dna = "".join(a + b + c for a, b, c in itertools.product(letters, repeat=3))
    for table in V.ALL_ACCEPTED_TABLES:
        got = qd.DnaSequence(dna).translate(table=table).seq
        assert got == oracle.translate_bytes(table, dna), table
        assert qd.DnaSequence(dna.lower()).translate(table=table).seq == got


def test_exhaustive_codons_mask_encoding_all_tables_all_frame_modes(N):
    """All 15^3 codons, every table, typed (mask) input through the C ABI: one frame (direct codon table), three and
    six frames (pair decode + folded LUT) must agree with the oracle and with each other."""
    ctx = N.default_context()
    L = N.lib()
    masks = np.array([m for a in range(1, 16) for b in range(1, 16) for c in range(1, 16) for m in (a, b, c)], dtype=np.uint8)
    n = masks.size
    for table in V.ALL_ACCEPTED_TABLES:
        slot = oracle.table_slot(table)
        want1 = oracle.translate_masks(slot, masks)
        want6 = oracle.all_frames_masks(slot, masks)
        for frames, want in ((1, [want1]), (3, want6[:3]), (6, want6)):
            out = np.zeros(2 * n + 16, dtype=np.uint8)
            lens = (C.c_size_t * 6)()
            nf = C.c_int()
            err = N.QdErr()
            rc = L.qd_translate_frames(ctx.handle, table, 1, 0, frames, masks.ctypes.data_as(C.c_void_p), n,
                                       out.ctypes.data_as(C.c_void_p), lens, C.byref(nf), C.byref(err))
            assert rc == 0
            got, off = [], 0
            for ln in lens:
                if ln:
                    got.append(out[off:off + ln].tobytes())
                off += ln
            assert got == want, (table, frames)
        # strict: the first ambiguous mask is reported, kind UnexpectedAmbiguous
        err = N.QdErr()
        out = np.zeros(n, dtype=np.uint8)
        rc = L.qd_translate(ctx.handle, table, 1, 1, masks.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), C.byref(err))
        first_amb = int(np.argmax(np.array([bin(int(m)).count("1") > 1 for m in masks])))
        assert (rc, err.kind, err.pos) == (3, 3, first_amb)


def test_device_lut_image_matches_tables_dat(N):
    img = C.string_at(N.lib().qd_table_data(0), 27 * 4096)
    assert hashlib.sha256(img).hexdigest() == V.TABLES_DAT_SHA256
    assert img == oracle.tables()


def test_every_length_0_to_400(qd, tiles):
    rng = np.random.default_rng(1)
    for n in range(0, 401):
        dna = _rand(rng, n, IUPAC)
        d = qd.DnaSequence(dna)
        assert d.translate(table=1 + n % 6).seq == oracle.translate_bytes(1 + n % 6, dna), n
        assert d.reverse_complement().seq == oracle.revcomp_bytes(dna), n
        assert [f.seq for f in d.translate_all_frames(table=11)] == oracle.py_all_frames(11, dna), n
        assert [f.seq for f in d.translate_self_frames(table=2)] == oracle.py_self_frames(2, dna), n


def test_revcomp_alignments_and_sizes(N):
    # every (source misalignment, length mod 16) combination of the shifted-granule path
    rng = np.random.default_rng(2)
    ctx = N.default_context()
    L = N.lib()
    base = np.frombuffer(_rand(rng, 5100, IUPAC), dtype=np.uint8)
    for n in list(range(0, 70)) + [255, 256, 257, 1023, 4095, 4096, 4097, 4999]:
        for shift in (0, 1, 3, 7, 8, 15):
            src = base[shift:shift + n].copy()  # host alignment does not matter; device copy is aligned
            out = np.empty(max(n, 1), dtype=np.uint8)
            err = N.QdErr()
            rc = L.qd_revcomp(ctx.handle, 0, 0, src.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p), C.byref(err))
            assert rc == 0
            assert out[:n].tobytes() == oracle.revcomp_bytes(src.tobytes())


def test_mask_encoding(N):
    # typed Rust API: &[Nucleotide] / &[NucleotideAmbiguous] reinterpreted as bytes
    rng = np.random.default_rng(3)
    ctx = N.default_context()
    L = N.lib()
    for n in (0, 1, 2, 3, 4, 5, 47, 48, 49, 50, 51, 95, 96, 97, 1000, 5003):
        for ambiguous in (False, True):
            m = (rng.integers(1, 16, n) if ambiguous else np.array([1, 2, 4, 8])[rng.integers(0, 4, n)]).astype(np.uint8)
            src = m if n else np.zeros(1, np.uint8)
            out = np.zeros(2 * n + 16, dtype=np.uint8)
            lens = (C.c_size_t * 6)()
            nf = C.c_int()
            err = N.QdErr()
            rc = L.qd_translate_frames(ctx.handle, 1, 1, int(not ambiguous), 6, src.ctypes.data_as(C.c_void_p), n,
                                       out.ctypes.data_as(C.c_void_p), lens, C.byref(nf), C.byref(err))
            assert rc == 0, (n, ambiguous, err.kind, err.pos)
            want = oracle.all_frames_masks(0, m)
            got, off = [], 0
            for ln in lens:
                if ln:
                    got.append(out[off:off + ln].tobytes())
                off += ln
            assert got == want and nf.value == len(want)
            rcout = np.zeros(max(n, 1), dtype=np.uint8)
            rc = L.qd_revcomp(ctx.handle, 1, int(not ambiguous), src.ctypes.data_as(C.c_void_p), n,
                              rcout.ctypes.data_as(C.c_void_p), C.byref(err))
            assert rc == 0
            assert (rcout[:n] == oracle.revcomp_masks(m)).all()
    # invalid mask bytes and strictness
    bad = np.array([1, 2, 0, 4], dtype=np.uint8)
    err = N.QdErr()
    out = np.zeros(16, dtype=np.uint8)
    rc = L.qd_revcomp(ctx.handle, 1, 0, bad.ctypes.data_as(C.c_void_p), 4, out.ctypes.data_as(C.c_void_p), C.byref(err))
    assert (rc, err.kind, err.pos, err.byte) == (5, 5, 2, 0)
    amb = np.array([1, 2, 4, 15, 3], dtype=np.uint8)
    rc = L.qd_revcomp(ctx.handle, 1, 1, amb.ctypes.data_as(C.c_void_p), 5, out.ctypes.data_as(C.c_void_p), C.byref(err))
    assert (rc, err.kind, err.pos, err.byte) == (3, 3, 3, ord("N"))


def test_first_error_position_random(qd, tiles):
    from quickdna_b200 import quickdna as native

    rng = np.random.default_rng(4)
    for trial in range(120):
        n = int(rng.integers(1, 3000))
        dna = bytearray(_rand(rng, n, b"ACGTacgt"))
        for _ in range(int(rng.integers(1, 4))):
            dna[int(rng.integers(0, n))] = int(rng.choice(list(b"NRYqZ-\n \x00\xc4\xff*")))
        dna = bytes(dna)
        for strict in (False, True):
            for name, gpu, cpu in (
                ("translate", lambda: (native._translate_strict if strict else native._translate)(1, dna),
                 lambda: oracle.translate_bytes(1, dna, strict=strict)),
                ("revcomp", lambda: (native._reverse_complement_strict if strict else native._reverse_complement)(dna),
                 lambda: oracle.revcomp_bytes(dna, strict=strict)),
                ("all_frames", lambda: [f.seq for f in qd.DnaSequence(dna).translate_all_frames(1, strict)],
                 lambda: oracle.py_all_frames(1, dna, strict=strict)),
            ):
                try:
                    want = cpu()
                    werr = None
                except oracle.OracleError as e:
                    want, werr = None, e
                if werr is None:
                    assert gpu() == want, (name, strict, trial)
                else:
                    with pytest.raises(ValueError) as ei:
                        gpu()
                    assert (ei.value.kind, ei.value.byte, ei.value.pos, str(ei.value)) == (
                        werr.kind, werr.byte, werr.pos, str(werr)), (name, strict, trial)


# ------------------------------------------------------------------ batches


@pytest.fixture(params=["default-tiles", "big-tiles"])
def tiles(request, N):
    """Frame launches pick the 256- or the 1024-thread tile by size; force the big one onto small inputs too."""
    ctx = N.default_context()
    if request.param == "big-tiles":
        assert N.lib().qd_ctx_set_big_tile_min_runs(ctx.handle, 0) == 0
    yield request.param
    N.lib().qd_ctx_set_big_tile_min_runs(ctx.handle, 2 * 1024 * N.lib().qd_ctx_sm_count(ctx.handle))


def _check_batch(B, reads, table, strict, frames):
    res = B.translate_frames_batch(reads, table=table, strict=strict, frames=frames)
    slot = oracle.table_slot(table)
    for i, r in enumerate(reads):
        r = bytes(r)
        if frames == 6:
            want = oracle.py_all_frames(table, r, strict=strict) if len(r) >= 3 else []
        elif frames == 3:
            want = oracle.py_self_frames(table, r, strict=strict)
        else:
            want = [oracle.translate_bytes(table, r, strict=strict)] if len(r) >= 3 else []
        assert res.sequence_frames(i) == want, (i, len(r), frames)
    # bytes of a slot outside its frame are zero
    total_frames = sum(len(x) for i in range(len(reads)) for x in res.sequence_frames(i))
    assert int(np.count_nonzero(res.data)) <= total_frames
    return res


def test_batch_variable_lengths(B, tiles):
    rng = np.random.default_rng(5)
    lens = [0, 1, 2, 3, 4, 5, 47, 48, 49, 50, 0, 0, 96, 97, 1000, 12287, 12288, 12289, 30000, 7, 0]
    lens += [int(x) for x in rng.integers(0, 2500, 300)]
    reads = [_rand(rng, n, IUPAC) for n in lens]
    for frames in (6, 3, 1):
        _check_batch(B, reads, 1, False, frames)
    _check_batch(B, [_rand(rng, n) for n in lens], 4, True, 6)


def test_batch_long_sequences_among_reads(B, tiles):
    """The 32-run map of a variable-length batch is written per sequence: by its thread (up to 512 runs of 48 nt), by a warp
    (up to 65536 runs) or by a CTA (longer: a chromosome among reads) of a second kernel.  Lengths on every boundary."""
    rng = np.random.default_rng(15)
    lens = [5, 0, 48 * 65536 + 1, 17, 0, 120_000, 1000, 48 * 512 + 1, 48 * 512, 0, 48 * 65536, 3, 48 * 65536 + 4800, 64, 48 * 513, 999]
    lens += [int(x) for x in rng.integers(20_000, 60_000, 70)]  # more than one warp's worth of list entries
    reads = [_rand(rng, n) for n in lens]
    for frames in (6, 1):
        res = B.translate_frames_batch(reads, table=1, frames=frames)
        for i, r in enumerate(reads):
            want = (oracle.py_all_frames(1, r) if frames == 6 else [oracle.translate_bytes(1, r)]) if len(r) >= 3 else []
            assert res.sequence_frames(i) == want, (i, len(r), frames)
    bad = list(reads)
    bad[12] = bad[12][:3_000_000] + b"!" + bad[12][3_000_001:]
    with pytest.raises(ValueError) as ei:
        B.translate_frames_batch(bad, table=1, frames=6)
    assert (ei.value.seq_index, ei.value.pos) == (12, 3_000_000)


def test_batch_bad_offsets_are_refused(B, N):
    """CSR offsets are caller data.  Host call: decreasing offsets -> invalid argument before anything runs.  Device call: the
    scan skips sequences whose offsets decrease or leave the buffer (nothing outside it is read), the error fetch names the
    first one; the context stays usable."""
    import torch

    ctx = N.default_context()
    rng = np.random.default_rng(16)
    reads = [_rand(rng, n) for n in (100, 200, 300, 400)]
    flat = np.frombuffer(b"".join(reads), dtype=np.uint8)
    good = np.array([0, 100, 300, 600, 1000], dtype=np.uint64)
    for bad, first in ((np.array([0, 100, 50, 600, 1000], dtype=np.uint64), 1),
                       (np.array([0, 100, 300, 600, 2**40], dtype=np.uint64), 3),
                       (np.array([0, 100, 300, 2**63 + 5, 1000], dtype=np.uint64), 2)):
        d_flat, d_offs = torch.from_numpy(flat.copy()).cuda(), torch.from_numpy(bad.view(np.int64).copy()).cuda()
        out = torch.zeros(int(N.lib().qd_batch_out_bytes(good.ctypes.data_as(C.POINTER(C.c_uint64)), 4, 6)) + 4096, dtype=torch.uint8, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        L = N.lib()
        assert L.qd_dev_error_reset(ctx.handle, s) == 0
        rc = L.qd_translate_frames_batch_dev(ctx.handle, 1, N.ENC_ASCII, 0, 6, d_flat.data_ptr(), d_offs.data_ptr(), 4, flat.size, out.data_ptr(), s)
        assert rc == 0
        err = N.QdErr()
        rc = L.qd_dev_error_fetch(ctx.handle, s, N.ENC_ASCII, 0, d_flat.data_ptr(), d_offs.data_ptr(), 4, 0, C.byref(err))
        assert rc == N.QD_ERR_INVALID_ARGUMENT and err.seq_index == first, (rc, err.seq_index, first)
    # decreasing offsets on the host path
    with pytest.raises(ValueError):
        B.translate_frames_batch((flat, np.array([0, 100, 50, 600, 1000], dtype=np.uint64)), table=1, frames=6)
    # a nucleotide error takes precedence over bad offsets; then a clean call works
    dirty = flat.copy()
    dirty[42] = ord("!")
    d_flat, d_offs = torch.from_numpy(dirty).cuda(), torch.from_numpy(np.array([0, 100, 300, 200, 1000], dtype=np.int64)).cuda()
    out = torch.zeros(1 << 16, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    L = N.lib()
    assert L.qd_dev_error_reset(ctx.handle, s) == 0
    assert L.qd_translate_frames_batch_dev(ctx.handle, 1, N.ENC_ASCII, 0, 6, d_flat.data_ptr(), d_offs.data_ptr(), 4, dirty.size, out.data_ptr(), s) == 0
    err = N.QdErr()
    rc = L.qd_dev_error_fetch(ctx.handle, s, N.ENC_ASCII, 0, d_flat.data_ptr(), d_offs.data_ptr(), 4, 0, C.byref(err))
    assert rc == N.QD_ERR_BAD_NUCLEOTIDE and (err.seq_index, err.pos) == (0, 42)
    _check_batch(B, reads, 1, False, 6)


def test_batch_many_tiny_reads(B, tiles):
    # more sequences starting inside one tile than the shared-memory search window holds
    rng = np.random.default_rng(6)
    reads = [_rand(rng, int(n), IUPAC) for n in rng.integers(0, 6, 5000)]
    _check_batch(B, reads, 1, False, 6)


def test_batch_uniform(B, tiles):
    rng = np.random.default_rng(7)
    for n_reads, read_len in ((1, 1000), (37, 1000), (300, 999), (513, 150), (64, 48), (64, 49), (9, 2), (5, 3), (3, 20000)):
        a = np.frombuffer(_rand(rng, n_reads * read_len, IUPAC), dtype=np.uint8).reshape(n_reads, read_len)
        for frames in (6, 3, 1):
            res = B.translate_frames_batch(a, table=11, frames=frames)
            for i in range(n_reads):
                r = a[i].tobytes()
                want = {6: oracle.py_all_frames(11, r) if read_len >= 3 else [],
                        3: oracle.py_self_frames(11, r),
                        1: [oracle.translate_bytes(11, r)] if read_len >= 3 else []}[frames]
                assert res.sequence_frames(i) == want, (n_reads, read_len, frames, i)


def test_batch_chunked_pipeline_and_error_index(B, N, tiles):
    rng = np.random.default_rng(8)
    ctx = N.default_context()
    assert N.lib().qd_ctx_set_chunk_bytes(ctx.handle, 1 << 16) == 0  # force many chunks
    try:
        a = np.frombuffer(_rand(rng, 700 * 1000), dtype=np.uint8).reshape(700, 1000).copy()
        res = B.translate_frames_batch(a, table=1, frames=6)
        for i in (0, 65, 66, 131, 699):
            assert res.sequence_frames(i) == oracle.py_all_frames(1, a[i].tobytes())
        reads = [a[i, : 400 + i].tobytes() for i in range(600)]
        _check_batch(B, reads, 1, True, 6)
        a[501, 777] = ord("n")
        a[650, 3] = ord("Q")
        with pytest.raises(ValueError) as ei:
            B.translate_frames_batch(a, table=1, strict=True, frames=6)
        assert (ei.value.seq_index, ei.value.pos, ei.value.kind, ei.value.byte) == (501, 777, 3, ord("N"))
        with pytest.raises(ValueError) as ei:
            B.translate_frames_batch(a, table=1, strict=False, frames=6)
        assert (ei.value.seq_index, ei.value.pos, ei.value.kind, ei.value.byte) == (650, 3, 2, ord("Q"))
        reads[123] = reads[123][:50] + b"-" + reads[123][51:]
        with pytest.raises(ValueError) as ei:
            B.translate_frames_batch(reads, table=1, frames=6)
        assert (ei.value.seq_index, ei.value.pos, str(ei.value)) == (123, 50, "bad nucleotide: '-'")
    finally:
        N.lib().qd_ctx_set_chunk_bytes(ctx.handle, 64 << 20)


# ------------------------------------------------------------------ parse (src/rust_api.rs:310-322)


def test_parse_vectors(B):
    # accepted alphabet, case folding, space / tab skipping (src/rust_api.rs:421-457)
    for text in ("", " ", "\t \t", "ACGT", "a c\tg t", "aAtTcCgGmMrRwWsSyYkKvVhHdDbBnN", " A", "A ", "AC GT" * 100):
        got = B.parse(text.encode())
        assert got.tobytes() == oracle.parse(text.encode()).tobytes(), text
    assert B.parse(b"acgt ACGT", strict=True).tobytes() == bytes([1, 4, 8, 2, 1, 4, 8, 2])
    for text, strict, kind, byte, pos in (("ABCD", True, 3, ord("B"), 1), ("AAAelephant", False, 2, ord("e"), 3),
                                          ("AA\xc4\x8dCCG", False, 1, 0xC4, 2), ("AC\nGT", False, 2, 10, 2),
                                          ("  \t- ", False, 2, ord("-"), 3)):
        with pytest.raises(ValueError) as ei:
            B.parse(text.encode("latin-1"), strict=strict)
        assert (ei.value.kind, ei.value.byte, ei.value.pos) == (kind, byte, pos), text


def test_parse_random(B):
    rng = np.random.default_rng(21)
    for trial in range(60):
        n = int(rng.integers(0, 40000))
        alphabet = b"ACGTacgt" if trial % 3 == 0 else IUPAC
        a = np.frombuffer(_rand(rng, n, alphabet + b" \t" * int(rng.integers(0, 4))), dtype=np.uint8).copy()
        strict = trial % 3 == 0
        lead = int(rng.integers(0, 16))  # exercise every input alignment
        buf = np.zeros(n + lead, dtype=np.uint8)
        buf[lead:] = a
        assert B.parse(buf[lead:], strict=strict).tobytes() == oracle.parse(a.tobytes(), strict=strict).tobytes(), trial
        if n and trial % 2:
            a[int(rng.integers(0, n))] = int(rng.choice(list(b"qZ-\n\x00\xff*")))
            try:
                oracle.parse(a.tobytes(), strict=strict)
                raise AssertionError("oracle accepted a bad byte")
            except oracle.OracleError as e:
                with pytest.raises(ValueError) as ei:
                    B.parse(a, strict=strict)
                assert (ei.value.kind, ei.value.byte, ei.value.pos, str(ei.value)) == (e.kind, e.byte, e.pos, str(e))


def test_parse_large_device(B, N):
    import torch

    n = 64_000_001
    rng = np.random.default_rng(22)
    a = np.frombuffer(b"ACGTNRY \t", dtype=np.uint8)[rng.integers(0, 9, n)]
    x = torch.from_numpy(a).cuda()
    B.dev_error_reset()
    out, count = B.parse_dev(x)
    assert B.dev_error_fetch(x) is None
    want = oracle.parse(a.tobytes())
    k = int(count.item())
    assert k == want.size
    assert out[:k].cpu().numpy().tobytes() == want.tobytes()


def test_parse_dev_clean_blocks_then_dropped_bytes(B):
    """The count pass stores clean 16-byte granules in place and the write pass skips the 4 KiB blocks that the scan
    confirms: every mix of confirmed, shifted and compacted blocks, and inputs / outputs of different 16-byte phases."""
    import torch

    rng = np.random.default_rng(23)
    n = 9 * 4096 + 123
    for first_blank in (None, 0, 5, 4095, 4096, 3 * 4096 + 17, 16383, 16384, 2 * 16384 + 17, n - 1):
        for in_lead, out_lead in ((0, 0), (0, 16), (3, 3), (3, 0), (16, 5), (7, 23)):
            a = np.frombuffer(_rand(rng, n, b"ACGTacgtNRY"), dtype=np.uint8).copy()
            if first_blank is not None:
                a[first_blank] = 0x20
                a[first_blank + 1:][rng.random(n - first_blank - 1) < 0.002] = 0x09
            src = torch.zeros(n + 32, dtype=torch.uint8, device="cuda")
            src[in_lead:in_lead + n] = torch.from_numpy(a).cuda()
            dst = torch.full((n + 64,), 0xEE, dtype=torch.uint8, device="cuda")
            B.dev_error_reset()
            out, count = B.parse_dev(src[in_lead:in_lead + n], out=dst[out_lead:out_lead + n])
            assert B.dev_error_fetch(src[in_lead:in_lead + n]) is None
            want = oracle.parse(a.tobytes())
            k = int(count.item())
            assert k == want.size, (first_blank, in_lead, out_lead)
            assert out[:k].cpu().numpy().tobytes() == want.tobytes(), (first_blank, in_lead, out_lead)
            assert (dst[:out_lead] == 0xEE).all() and (dst[out_lead + n:] == 0xEE).all()  # nothing outside the caller's n bytes


# ------------------------------------------------------------------ windows / canonical / expansions


@pytest.mark.parametrize("dna,canon", V.CANONICAL[:-1])
def test_canonical_vectors(B, dna, canon):
    got = B.canonical_windows(dna.encode(), w=len(dna))
    assert got.shape == (1, len(dna)) and got[0].tobytes() == canon.encode()


def test_canonical_whole_sequence(B):
    # all inputs of length <= 3 (src/canonical.rs:216-404 enumerates them) and random longer ones, both variants
    for n in range(0, 4):
        for tup in itertools.product(b"ATCG", repeat=n):
            s = bytes(tup)
            m = oracle.masks_of(s) if n else np.zeros(0, np.uint8)
            assert B.canonical(s).tobytes() == (oracle.ascii_of(oracle.canonical(m)) if n else b"")
            assert B.canonical(s, forward_only=True).tobytes() == (oracle.ascii_of(oracle.forward_canonical(m)) if n else b"")
    rng = np.random.default_rng(31)
    for trial in range(80):
        n = int(rng.integers(1, 200)) if trial < 50 else int(rng.integers(1000, 300000))
        alphabet = [b"ACGT", b"AC", b"A", b"ACGTacgt", b"GT"][trial % 5]
        s = _rand(rng, n, alphabet)
        if trial % 7 == 0:  # palindromic under relabelling: the difference search runs to the end
            s = s + s[::-1]
        m = oracle.masks_of(s)
        assert B.canonical(s).tobytes() == oracle.ascii_of(oracle.canonical(m)), trial
        assert B.canonical(m, enc=1).tobytes() == oracle.canonical(m).tobytes(), trial
        assert B.canonical(s, forward_only=True).tobytes() == oracle.ascii_of(oracle.forward_canonical(m)), trial
    with pytest.raises(ValueError) as ei:
        B.canonical(b"ACGTNACGT")
    assert (ei.value.kind, ei.value.byte, ei.value.pos) == (3, ord("N"), 4)


def test_canonical_windows_batch(B):
    """Batched form: windows never span two sequences; tiles of 64 windows cross sequence seams."""
    import torch

    rng = np.random.default_rng(33)
    for n_seq, seq_len, w in ((1, 5000, 42), (7, 1000, 42), (300, 50, 42), (500, 42, 42), (64, 43, 42), (33, 99, 17),
                              (10, 700, 64), (20, 333, 33), (5, 2000, 1), (3, 100, 16), (9, 257, 48), (2, 99999, 42)):
        a = np.frombuffer(_rand(rng, n_seq * seq_len, b"ACGTacgt"), dtype=np.uint8).reshape(n_seq, seq_len)
        x = torch.from_numpy(a.copy()).cuda()
        B.dev_error_reset()
        got = B.canonical_windows_batch_dev(x, n_seq, seq_len, w).cpu().numpy()
        assert B.dev_error_fetch(x, uniform_len=seq_len, strict=True) is None
        nw = seq_len - w + 1
        assert got.shape == (n_seq * nw, w)
        for s in range(n_seq):
            want = oracle.ascii_of(oracle.canonical_windows(oracle.masks_of(a[s]), w).reshape(-1))
            assert got[s * nw:(s + 1) * nw].tobytes() == want, (n_seq, seq_len, w, s)
    # an invalid base anywhere in the batch is reported with its sequence and position
    a = np.frombuffer(_rand(rng, 40 * 500, b"ACGT"), dtype=np.uint8).reshape(40, 500).copy()
    a[17, 321] = ord("N")
    x = torch.from_numpy(a).cuda()
    B.dev_error_reset()
    B.canonical_windows_batch_dev(x, 40, 500, 42)
    err = B.dev_error_fetch(x, uniform_len=500, strict=True)
    assert err is not None and (err.seq_index, err.pos, err.kind, err.byte) == (17, 321, 3, ord("N"))


def test_canonical_windows_random(B):
    rng = np.random.default_rng(9)
    for n, w in ((10042 - 1, 42), (500, 1), (500, 7), (3000, 33), (1000, 64), (41, 42), (42, 42), (300, 20)):
        dna = _rand(rng, n, b"ACGTacgt")
        m = oracle.masks_of(dna)
        for fwd_only in (False, True):
            got = B.canonical_windows(dna, w=w, forward_only=fwd_only)
            if fwd_only:
                want = np.stack([oracle.forward_canonical(m[i:i + w]) for i in range(max(0, n - w + 1))]) if n >= w else np.empty((0, w), np.uint8)
            else:
                want = oracle.canonical_windows(m, w)
            assert got.shape == want.shape
            assert got.tobytes() == oracle.ascii_of(want.reshape(-1)), (n, w, fwd_only)
        gm = B.canonical_windows(m, w=w, enc=1)
        assert (gm == oracle.canonical_windows(m, w)).all()
    with pytest.raises(ValueError) as ei:
        B.canonical_windows(b"ACGTNACGT", w=4)
    assert (ei.value.kind, ei.value.pos, ei.value.byte) == (3, 4, ord("N"))


def _pack_windows(win_masks: np.ndarray) -> np.ndarray:
    """[count, w] masks {1, 2, 4, 8} -> [count, 2 * ceil(w / 32)] uint32: the QD_CANON_KEY_PACKED layout restated with numpy --
    the low bits of the codes (A, T, C, G = 0..3; nucleotide j at bit j % 32 of word j / 32), then the high bits."""
    count, w = win_masks.shape
    code = np.zeros(16, np.uint64)
    code[[1, 2, 4, 8]] = [0, 1, 2, 3]
    pw = (w + 31) // 32
    c = np.zeros((count, pw * 32), np.uint64)
    c[:, :w] = code[win_masks]
    shifts = np.arange(32, dtype=np.uint64)[None, None, :]
    lo = ((c & np.uint64(1)).reshape(count, pw, 32) << shifts).sum(axis=2, dtype=np.uint64)
    hi = ((c >> np.uint64(1)).reshape(count, pw, 32) << shifts).sum(axis=2, dtype=np.uint64)
    return np.concatenate([lo, hi], axis=1).astype(np.uint32)


def _fingerprint(packed: np.ndarray, w: int) -> np.ndarray:
    """The QD_CANON_KEY_FP64 definition of include/quickdna_b200.h restated with numpy uint64 arithmetic."""
    def mix(z):
        z = z ^ (z >> np.uint64(30))
        z = z * np.uint64(0xbf58476d1ce4e5b9)
        z = z ^ (z >> np.uint64(27))
        z = z * np.uint64(0x94d049bb133111eb)
        return z ^ (z >> np.uint64(31))

    count, words = packed.shape
    pw = words // 2
    p = packed.astype(np.uint64)
    with np.errstate(over="ignore"):
        h = np.full(count, (0x9e3779b97f4a7c15 * w) & (2**64 - 1), dtype=np.uint64)
        for k in range(pw):
            h = mix(h ^ (p[:, k] | (p[:, pw + k] << np.uint64(32))))
    return h


def _want_windows(m: np.ndarray, w: int, fwd_only: bool) -> np.ndarray:
    if not fwd_only:
        return oracle.canonical_windows(m, w)
    n = len(m)
    return np.stack([oracle.forward_canonical(m[i:i + w]) for i in range(n - w + 1)]) if n >= w else np.empty((0, w), np.uint8)


def test_canonical_window_keys_single_sequence(B, N):
    """packed keys ARE the oracle's canonical windows at 2 bits per nucleotide; the fingerprint is
    the documented function of them.  Low-complexity alphabets force the full-width decision path."""
    rng = np.random.default_rng(77)
    cases = [(3000, 42, b"ACGTacgt"), (700, 42, b"A"), (900, 42, b"AC"), (900, 42, b"ACG"), (500, 1, b"ACGT"), (500, 7, b"ACGT"),
             (800, 16, b"ACGT"), (800, 17, b"GT"), (1000, 32, b"ACGT"), (1000, 33, b"ACGT"), (1200, 48, b"AACC"), (1500, 64, b"ACGT"),
             (41, 42, b"ACGT"), (42, 42, b"ACGT"), (64, 64, b"TG"), (300, 20, b"ACGT")]
    for n, w, alphabet in cases:
        dna = _rand(rng, n, alphabet)
        if alphabet == b"AACC":  # long runs, then a palindrome under relabelling
            dna = dna[: n // 2] + dna[: n - n // 2][::-1]
        m = oracle.masks_of(dna)
        for fwd_only in (False, True):
            want = _pack_windows(_want_windows(m, w, fwd_only))
            got = B.canonical_window_keys(dna, w=w, kind=N.CANON_KEY_PACKED, forward_only=fwd_only)
            assert got.shape == want.shape and (got == want).all(), (n, w, alphabet, fwd_only)
            fp = B.canonical_window_keys(dna, w=w, kind=N.CANON_KEY_FP64, forward_only=fwd_only)
            assert fp.dtype == np.uint64 and (fp == _fingerprint(want, w)).all(), (n, w, alphabet, fwd_only)
        gm = B.canonical_window_keys(m, w=w, kind=N.CANON_KEY_PACKED, enc=1)
        assert (gm == _pack_windows(oracle.canonical_windows(m, w))).all()
    # the reference's own vectors (src/canonical.rs:200-214, 406-412; src/rust_api.rs:365-367) through the packed key
    for dna, canon in V.CANONICAL[:-1]:
        got = B.canonical_window_keys(dna.encode(), w=len(dna), kind=N.CANON_KEY_PACKED)
        assert (got == _pack_windows(oracle.masks_of(canon.encode())[None, :])).all(), dna
    with pytest.raises(ValueError) as ei:
        B.canonical_window_keys(b"ACGTNACGT", w=4)
    assert (ei.value.kind, ei.value.pos, ei.value.byte) == (3, 4, ord("N"))
    with pytest.raises(ValueError):
        B.canonical_window_keys(b"ACGT" * 40, w=65)


def test_canonical_window_keys_batch(B, N):
    """Batched device form: 256-window warp tiles cross sequence seams; keys equal the materialised canonical windows."""
    import torch

    rng = np.random.default_rng(78)
    for n_seq, seq_len, w in ((1, 5000, 42), (7, 1000, 42), (300, 50, 42), (500, 42, 42), (64, 43, 42), (33, 99, 17), (10, 700, 64),
                              (20, 333, 33), (5, 2000, 1), (3, 100, 16), (9, 257, 48), (2, 99999, 42), (1000, 297, 42), (4, 298, 42)):
        a = np.frombuffer(_rand(rng, n_seq * seq_len, b"ACGTacgt"), dtype=np.uint8).reshape(n_seq, seq_len)
        x = torch.from_numpy(a.copy()).cuda()
        B.dev_error_reset()
        packed = B.canonical_window_keys_dev(x, n_seq, seq_len, w, kind=N.CANON_KEY_PACKED).cpu().numpy().view(np.uint32)
        fp = B.canonical_window_keys_dev(x, n_seq, seq_len, w, kind=N.CANON_KEY_FP64).cpu().numpy().view(np.uint64)
        rows = B.canonical_windows_batch_dev(x, n_seq, seq_len, w).cpu().numpy()
        assert B.dev_error_fetch(x, uniform_len=seq_len, strict=True) is None
        nw = seq_len - w + 1
        want = _pack_windows(oracle.masks_of(rows.reshape(-1)).reshape(n_seq * nw, w))  # rows are oracle-checked in test_canonical_windows_batch
        assert packed.shape == want.shape and (packed == want).all(), (n_seq, seq_len, w)
        assert (fp == _fingerprint(want, w)).all(), (n_seq, seq_len, w)
        for s in range(0, n_seq, max(1, n_seq // 5)):  # and directly against the oracle on a few sequences
            assert (packed[s * nw:(s + 1) * nw] == _pack_windows(oracle.canonical_windows(oracle.masks_of(a[s]), w))).all()
    a = np.frombuffer(_rand(rng, 40 * 500, b"ACGT"), dtype=np.uint8).reshape(40, 500).copy()
    a[17, 321] = ord("N")
    x = torch.from_numpy(a).cuda()
    B.dev_error_reset()
    B.canonical_window_keys_dev(x, 40, 500, 42)
    err = B.dev_error_fetch(x, uniform_len=500, strict=True)
    assert err is not None and (err.seq_index, err.pos, err.kind, err.byte) == (17, 321, 3, ord("N"))


def test_canonical_window_keys_full_size_properties(B, N):
    """src/canonical.rs:414-437 at full size, through the keys: the canonical form does not change when the sequence is
    reversed or when its bases are renamed, so window i of x and window nw - 1 - i of reverse(x) share their key, and so
    do x and any relabelling of x."""
    import torch

    from quickdna_b200.synth import synth_torch

    n, w = 20_000_003, 42
    x = synth_torch(9, 0, n, torch.device("cuda:0"))
    B.dev_error_reset()
    k = B.canonical_window_keys_dev(x, 1, n, w)
    kr = B.canonical_window_keys_dev(torch.flip(x, dims=[0]).contiguous(), 1, n, w)
    assert torch.equal(k, torch.flip(kr, dims=[0]))
    table = torch.arange(256, dtype=torch.uint8, device=x.device)
    for a, b in zip(b"ACGT", b"GTAC"):
        table[a] = b
    kp = B.canonical_window_keys_dev(table[x.long()], 1, n, w)
    assert torch.equal(k, kp)
    assert B.dev_error_fetch(x, uniform_len=n, strict=True) is None
    # the packed key of the forward-only form starts with code 0 and is a restricted-growth string: its first nucleotide is A
    pf = B.canonical_window_keys_dev(x, 1, n, w, kind=N.CANON_KEY_PACKED, forward_only=True)
    assert int((pf[:, 0] & 1).sum()) == 0 and int((pf[:, 2] & 1).sum()) == 0


@pytest.mark.parametrize("dna,exp", V.EXPANSIONS)
def test_expansion_vectors(B, dna, exp):
    a = np.frombuffer(dna.encode(), dtype=np.uint8).reshape(1, -1)
    counts, out_index, rows = B.expansion_windows(a, max_expansions=64)
    assert int(counts[0]) == len(exp) and int(out_index[1]) == len(exp)
    assert [r.tobytes().decode() for r in rows] == exp


def test_expansion_windows_random(B):
    # benches/expansions.rs:68-102 shape: 42-nt windows, 2 ambiguities from the 11-code list, cap 16
    rng = np.random.default_rng(10)
    n_win, w = 1000, 42
    a = np.frombuffer(_rand(rng, n_win * w), dtype=np.uint8).reshape(n_win, w).copy()
    codes = np.frombuffer(b"WMYHRKDSVBN", dtype=np.uint8)
    for i in range(n_win):
        for j in rng.choice(w, size=int(rng.integers(0, 4)), replace=False):
            a[i, j] = codes[rng.integers(0, 11)]
    counts, out_index, rows = B.expansion_windows(a, max_expansions=16)
    at = 0
    for i in range(n_win):
        m = oracle.masks_of(a[i])
        hint = oracle.expansions_size_hint(m)
        assert int(counts[i]) == hint
        if hint <= 16:
            assert int(out_index[i]) == at
            want = oracle.expansions(m)
            assert rows[at:at + hint].tobytes() == oracle.ascii_of(want.reshape(-1))
            at += hint
        else:
            assert int(out_index[i]) == 2 ** 64 - 1
    assert at == len(rows) == int(out_index[n_win])
    big = np.frombuffer((V.SIZE_HINT_OVERFLOW + "A" * 10).encode(), dtype=np.uint8).reshape(1, -1)
    counts, _, rows = B.expansion_windows(big)
    assert int(counts[0]) == 2 ** 64 - 1 and len(rows) == 0
    big = np.frombuffer(V.SIZE_HINT_BIG[0].encode(), dtype=np.uint8).reshape(1, -1)
    assert int(B.expansion_windows(big)[0][0]) == V.SIZE_HINT_BIG[1]


def test_expansion_windows_paths(B):
    """Tiled fast path (w <= 64, cap < 128), generic path (larger caps / longer windows), both encodings, device API."""
    import torch

    rng = np.random.default_rng(12)
    codes = np.frombuffer(b"WMYHRKDSVBNwmyhrkdsvbn", dtype=np.uint8)
    for n_win, w, cap, max_amb in ((257, 42, 16, 4), (100, 41, 100, 5), (64, 64, 127, 7), (50, 7, 16, 3), (33, 1, 4, 1),
                                   (70, 42, 1000, 6), (20, 100, 64, 4), (300, 42, 1, 2)):
        a = np.frombuffer(_rand(rng, n_win * w, b"ACGTacgt"), dtype=np.uint8).reshape(n_win, w).copy()
        for i in range(n_win):
            k = int(rng.integers(0, min(max_amb, w) + 1))
            for j in rng.choice(w, size=k, replace=False):
                a[i, j] = codes[rng.integers(0, len(codes))]
        masks = np.stack([oracle.masks_of(a[i]) for i in range(n_win)])
        for enc, arr in ((0, a), (1, masks)):
            counts, out_index, rows = B.expansion_windows(arr, max_expansions=cap, enc=enc)
            at = 0
            for i in range(n_win):
                hint = oracle.expansions_size_hint(masks[i])
                assert int(counts[i]) == hint, (n_win, w, cap, enc, i)
                if hint <= cap:
                    assert int(out_index[i]) == at
                    want = oracle.expansions(masks[i])
                    want = oracle.ascii_of(want.reshape(-1)) if enc == 0 else want.tobytes()
                    assert rows[at:at + hint].tobytes() == want, (n_win, w, cap, enc, i)
                    at += hint
                else:
                    assert int(out_index[i]) == 2 ** 64 - 1
            assert at == len(rows) == int(out_index[n_win])
        if w <= 64 and cap < 128:
            x = torch.from_numpy(a).cuda()
            B.dev_error_reset()
            c_d, oi_d, out_d = B.expansion_windows_dev(x, max_expansions=cap)
            assert B.dev_error_fetch(x.view(-1)) is None
            counts, out_index, rows = B.expansion_windows(a, max_expansions=cap)
            total = int(oi_d[n_win].item())
            assert total == len(rows)
            assert out_d[: total * w].cpu().numpy().tobytes() == rows.tobytes()
            assert (c_d.cpu().numpy().astype(np.uint64) == counts).all()
    with pytest.raises(ValueError) as ei:
        B.expansion_windows(np.frombuffer(b"ACGTXACG", dtype=np.uint8).reshape(2, 4))
    assert (ei.value.seq_index, ei.value.pos, ei.value.kind) == (1, 0, 2)


def test_windows(B):
    seq, wins = V.WINDOWS_10
    got = B.windows(seq.encode(), 10)
    assert [w.tobytes().decode() for w in got] == wins
    assert len(B.windows(b"antg", 10)) == 0
    rng = np.random.default_rng(11)
    s = _rand(rng, 3001, b"ACDEFGHIKLMNPQRSTVWY*")
    for w in (1, 20, 42, 127):
        assert (B.windows(s, w) == oracle.windows(s, w)).all()


def test_windows_all(B):
    rng = np.random.default_rng(12)
    for n in (0, 1, 4, 5, 41, 42, 43, 59, 60, 61, 62, 63, 64, 65, 200, 999, 99999):
        dna = _rand(rng, n, b"ACGTacgt")
        m = oracle.masks_of(dna) if n else np.zeros(0, np.uint8)
        d_want, a_want, o_want, c_want = oracle.windows_all(m, 42, 20)
        d, a, o, c = B.windows_all(dna, 42, 20)
        assert c == c_want, n
        assert d.tobytes() == d_want.tobytes() and a.tobytes() == a_want.tobytes()
        assert (o == o_want).all()
        if n:
            d2, a2, o2, c2 = B.windows_all(m, 42, 20, enc=1)
            assert d2.tobytes() == d_want.tobytes() and a2.tobytes() == a_want.tobytes() and c2 == c_want


# ------------------------------------------------------------------ device-resident API at scale (properties)


def test_device_api_revcomp_large_properties(B, N):
    import torch

    from quickdna_b200.synth import synth_np, synth_torch

    dev = torch.device("cuda:0")
    for n in (250_000_000, 250_000_001, 249_999_999):
        x = synth_torch(2, 0, n, dev)
        B.dev_error_reset()
        y = B.revcomp_dev(x)
        z = B.revcomp_dev(y)
        assert B.dev_error_fetch(x) is None
        assert torch.equal(z, x)  # involution on upper-case ACGT
        # sampled slices against the oracle
        for s0 in (0, 12345, n // 2 - 7, n - 4099):
            sl = x[s0:s0 + 4099].cpu().numpy()
            assert sl.tobytes() == synth_np(2, s0, 4099).tobytes()
            want = oracle.revcomp_bytes(sl.tobytes())
            got = y[n - s0 - 4099:n - s0].cpu().numpy().tobytes()
            assert got == want
        # composition: reversed complement histogram
        hx = torch.bincount(x[: 1 << 24].to(torch.int64), minlength=128)
        hy = torch.bincount(y[n - (1 << 24):].to(torch.int64), minlength=128)
        assert hx[ord("A")] == hy[ord("T")] and hx[ord("C")] == hy[ord("G")] and hx[ord("G")] == hy[ord("C")]
        del x, y, z
    # strict failure at a known position
    x = synth_torch(2, 0, 1_000_003, dev)
    x[777_777] = ord("N")
    B.dev_error_reset()
    B.revcomp_dev(x, strict=True)
    e = B.dev_error_fetch(x, strict=True)
    assert e is not None and (e.kind, e.pos, e.byte) == (3, 777_777, ord("N"))


def test_device_api_six_frames_large_properties(B, N):
    import torch

    from quickdna_b200.synth import synth_torch

    dev = torch.device("cuda:0")
    n_reads, L = 1_000_000, 1000
    x = synth_torch(3, 0, n_reads * L, dev)
    B.dev_error_reset()
    out = B.translate_frames_uniform_dev(x, n_reads, L)
    assert B.dev_error_fetch(x, uniform_len=L) is None
    per = 16 * B.runs(L) * 6
    o = out.view(n_reads, per)
    # sampled reads against the oracle
    for i in (0, 1, 4242, 500_000, n_reads - 1):
        r = x[i * L:(i + 1) * L].cpu().numpy().tobytes()
        want = oracle.py_all_frames(1, r)
        row = o[i].cpu().numpy()
        got = [row[B.slot_frame_offset(L, 6, f):B.slot_frame_offset(L, 6, f) + B.frame_len(L, f % 3)].tobytes() for f in range(6)]
        assert got == want
    # property over the whole batch: reverse frames of x == forward frames of revcomp(x), read by read
    xr = torch.empty_like(x)
    for i0 in range(0, n_reads, 250_000):  # revcomp each read: flip within rows, complement by LUT
        blk = x[i0 * L:(i0 + 250_000) * L].view(-1, L).flip(1)
        comp = torch.zeros(256, dtype=torch.uint8, device=dev)
        for a, b in zip(b"ACGT", b"TGCA"):
            comp[a] = b
        xr[i0 * L:(i0 + 250_000) * L] = comp[blk.reshape(-1).to(torch.int64)]
    out_r = B.translate_frames_uniform_dev(xr, n_reads, L).view(n_reads, per)
    for f in range(3):
        ln = B.frame_len(L, f)
        a0, b0 = B.slot_frame_offset(L, 6, 3 + f), B.slot_frame_offset(L, 6, f)
        assert torch.equal(o[:, a0:a0 + ln], out_r[:, b0:b0 + ln])
    # same data through the CSR path gives identical bytes
    offsets = torch.arange(0, (n_reads + 1) * L, L, dtype=torch.int64, device=dev)
    out_csr = B.translate_frames_csr_dev(x, offsets, total_runs=n_reads * B.runs(L))
    assert torch.equal(out_csr, out)
    # an error deep inside the batch is located exactly
    x[123_456 * L + 789] = ord("@")
    B.dev_error_reset()
    B.translate_frames_uniform_dev(x, n_reads, L, out=out)
    e = B.dev_error_fetch(x, uniform_len=L)
    assert e is not None and (e.seq_index, e.pos, str(e)) == (123_456, 789, "bad nucleotide: '@'")


# ------------------------------------------------------------------ FASTA record scanner (src/fasta.rs)


def _fasta_same(B, text, strict, concat, allow):
    try:
        want = oracle.fasta_scan(text, strict=strict, concatenate_headers=concat, allow_preceding_comment=allow)
        werr = None
    except oracle.FastaError as e:
        want, werr = None, e
    if werr is not None:
        with pytest.raises(ValueError) as ei:
            B.fasta_scan(text, strict=strict, concatenate_headers=concat, allow_preceding_comment=allow)
        assert (ei.value.kind, ei.value.byte, ei.value.pos, ei.value.line, ei.value.located()) == (
            werr.kind, werr.byte, werr.pos, werr.line, werr.located())
        return
    got = B.fasta_scan(text, strict=strict, concatenate_headers=concat, allow_preceding_comment=allow)
    assert len(got) == len(want)
    for i, (h, m, lr) in enumerate(want):
        gh, gm, glr = got.record(i)
        assert (gh, gm.tobytes(), glr) == (h, m.tobytes(), lr), (i, concat, allow, strict)


@pytest.mark.parametrize("text,settings,want", V.FASTA)
def test_fasta_vectors(B, text, settings, want):
    for concat, allow in settings:
        got = B.fasta_scan(text, concatenate_headers=concat, allow_preceding_comment=allow)
        assert [(got.record(i)[0], oracle.ascii_of(got.record(i)[1]).decode(), got.record(i)[2]) for i in range(len(got))] == want
    for concat in (False, True):
        for allow in (False, True):
            _fasta_same(B, text, False, concat, allow)
            _fasta_same(B, text, True, concat, allow)


@pytest.mark.parametrize("text,strict,line,kind,byte,message", V.FASTA_ERRORS)
def test_fasta_error_vectors(B, text, strict, line, kind, byte, message):
    with pytest.raises(ValueError) as ei:
        B.fasta_scan(text, strict=strict)
    assert (ei.value.line, ei.value.kind, ei.value.byte, ei.value.located()) == (line, kind, byte, message)


def test_fasta_random(B):
    """Random files: line lengths 0..200 (plus a few multi-block lines), LF and CRLF, headers in runs, blanks, and an
    occasional bad byte; every setting combination against the oracle."""
    rng = np.random.default_rng(41)
    for trial in range(40):
        parts = []
        n_lines = int(rng.integers(0, 400))
        crlf = trial % 3 == 0
        for _ in range(n_lines):
            r = rng.random()
            if r < 0.15:
                line = (b">" if rng.random() < 0.7 else b";") + _rand(rng, int(rng.integers(0, 30)), b"abc xyz>;\t01")
            elif r < 0.2:
                line = b" \t"[: int(rng.integers(0, 3))]
            else:
                ln = int(rng.integers(0, 200)) if rng.random() < 0.97 else int(rng.integers(4000, 20000))
                line = _rand(rng, ln, b"ACGTacgtNRY \t" if trial % 2 else b"ACGTacgt ")
            parts.append(line + (b"\r\n" if crlf else b"\n"))
        text = b"".join(parts)
        if trial % 4 == 1 and text.endswith(b"\n"):
            text = text[:-2] if crlf else text[:-1]  # no trailing line break
        if trial % 5 == 2 and len(text) > 10:
            t = bytearray(text)
            t[int(rng.integers(0, len(t)))] = int(rng.choice(list(b"xQ-\r\x00\xc4")))
            text = bytes(t)
        lead = int(rng.integers(0, 16))  # input alignment
        buf = np.frombuffer(b"\n" * lead + text, dtype=np.uint8)[lead:]
        for concat in (False, True):
            for allow in (False, True):
                _fasta_same(B, bytes(buf), trial % 2 == 0 and b"N" not in text, concat, allow)


def test_fasta_many_tiny_records(B):
    """More records than the wrapper's first guess (one per 64 bytes of text): the second, exactly sized call."""
    text = b"".join(b">%d\nAC\n" % i for i in range(500))
    for concat in (False, True):
        _fasta_same(B, text, False, concat, False)
    scan = B.fasta_scan(text)
    assert len(scan) == 500 and scan.headers[499] == "499" and scan.record(3)[1].tobytes() == bytes([1, 4])


def test_fasta_feeds_batch_translate(B):
    """The scanner's CSR offsets are what the batch entry points consume."""
    rng = np.random.default_rng(42)
    seqs = [_rand(rng, int(n), b"ACGTN") for n in rng.integers(0, 3000, 50)]
    text = b"".join(b">read%d some text\n" % i + b"\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + b"\n" for i, s in enumerate(seqs))
    scan = B.fasta_scan(text)
    assert len(scan) == 50 and scan.headers[7] == "read7 some text"
    res = B.translate_frames_batch([scan.record(i)[1] for i in range(50)], table=1, frames=6, enc=1)
    for i, s in enumerate(seqs):
        want = oracle.py_all_frames(1, s) if len(s) >= 3 else []
        assert res.sequence_frames(i) == want


# ------------------------------------------------------------------ threading contract


def test_contexts_are_independent_across_threads(N):
    """The reference's functions are pure and re-entrant; here: one context per thread, used concurrently, and one
    shared context hammered from several threads (its entry points serialise on a mutex)."""
    import threading

    rng = np.random.default_rng(51)
    inputs = [_rand(rng, int(n), IUPAC) for n in rng.integers(1000, 200000, 8)]
    want = [(oracle.translate_bytes(11, d), oracle.revcomp_bytes(d)) for d in inputs]
    L = N.lib()
    errors = []

    def work(ctx, idx, rounds):
        try:
            for _ in range(rounds):
                d = inputs[idx]
                out = C.create_string_buffer(max(len(d) // 3, 1))
                rc_out = C.create_string_buffer(len(d))
                err = N.QdErr()
                assert L.qd_translate(ctx.handle, 11, 0, 0, d, len(d), out, C.byref(err)) == 0
                assert L.qd_revcomp(ctx.handle, 0, 0, d, len(d), rc_out, C.byref(err)) == 0
                assert out.raw[: len(d) // 3] == want[idx][0] and rc_out.raw == want[idx][1]
        except BaseException as e:  # surfaced in the main thread
            errors.append(e)

    own = [N.Context(0) for _ in range(4)]
    threads = [threading.Thread(target=work, args=(own[i], i, 10)) for i in range(4)]
    shared = N.default_context()
    threads += [threading.Thread(target=work, args=(shared, 4 + i, 10)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for c in own:
        c.close()
    assert not errors, errors[0]


# ------------------------------------------------------------------ one long sequence cut into shards


def test_chromosome_shards_stitch_to_the_unsharded_result(qd):
    """Each shard is translated as an ordinary sequence (its nucleotides + a 2-nt halo); host-side placement alone
    (quickdna_b200.shard) rebuilds the six frames and the reverse complement of the whole sequence."""
    from quickdna_b200 import shard

    rng = np.random.default_rng(61)
    for n, world in ((100_003, 2), (1_000_001, 4), (48 * 7, 8), (50, 2), (999_999, 3), (5, 4), (3, 2), (97, 3), (49, 2)):
        seq = _rand(rng, n, IUPAC)
        def six(compact, length):  # the reference omits empty frames: put the returned ones back at their frame numbers
            present = [k for k in range(3) if length > k and (length - k) // 3 > 0]
            assert len(compact) == 2 * len(present)
            full = [b""] * 6
            for j, k in enumerate(present):
                full[k], full[3 + k] = compact[j], compact[len(present) + j]
            return full

        whole = six([f.seq for f in qd.DnaSequence(seq).translate_all_frames(11)], n)
        frames = [bytearray(len(f)) for f in whole]
        seen = [0] * 6
        rc = bytearray(n)
        for s, e, read_end in shard.partition_chromosome(n, world):
            if e <= s:
                continue
            piece = seq[s:read_end]
            local = six([f.seq for f in qd.DnaSequence(piece).translate_all_frames(11)], len(piece))
            for lf, (gf, first, ln) in enumerate(shard.chromosome_shard_placement(n, s, e)):
                assert ln == len(local[lf])
                frames[gf][first:first + ln] = local[lf]
                seen[gf] += ln
            rc[n - e:n - s] = qd.DnaSequence(seq[s:e]).reverse_complement().seq
        assert [bytes(f) for f in frames] == whole, (n, world)
        assert seen == [len(f) for f in whole]
        assert bytes(rc) == qd.DnaSequence(seq).reverse_complement().seq


# ------------------------------------------------------------------ beyond 2^32 nucleotides (64-bit offsets)


def test_offsets_beyond_4g_nucleotides(B, N):
    """A 4.4-G-nt buffer: as ONE sequence (reverse complement, six frames) and as a uniform batch; outputs are sampled
    near the start, across 2^32 and at the very end, and compared with the oracle on the matching slices."""
    import torch

    from quickdna_b200.synth import synth_torch

    dev = torch.device("cuda:0")
    n = (1 << 32) + 100_000_007
    x = synth_torch(13, 0, n, dev)
    B.dev_error_reset()
    # ---- reverse complement of one sequence: out[n-1-i] = comp(in[i])
    out = torch.empty(n + 16, dtype=torch.uint8, device=dev)
    B.revcomp_dev(x, out=out[:n])
    for lo in (0, (1 << 32) - 50, n - 1000):
        hi = min(n, lo + 1000)
        want = oracle.revcomp_bytes(x[lo:hi].cpu().numpy().tobytes())
        assert out[n - hi:n - lo].cpu().numpy().tobytes() == want, lo
    del out
    # ---- six frames of one sequence (slot layout): frame k, residues [a, b) <-> codon starts 3a+k ...
    runs = B.runs(n)
    out = torch.empty(max(runs * 16 * 6, (n // 1000) * 16 * B.runs(1000) * 6), dtype=torch.uint8, device=dev)
    L = N.lib()
    ctx = N.default_context()
    rc = L.qd_translate_frames_dev(ctx.handle, 1, 0, 0, 6, C.c_void_p(x.data_ptr()), n, C.c_void_p(out.data_ptr()), B._stream_ptr(torch))
    ctx.check(rc)
    assert B.dev_error_fetch(x) is None
    for p0 in (0, ((1 << 32) - 48) // 3 * 3, n - 3000 - (n - 3000) % 3):  # p0 multiple of 3
        piece = x[p0:min(n, p0 + 999)].cpu().numpy().tobytes()
        for k in range(3):
            want = oracle.translate_bytes(1, piece[k:])
            off = B.slot_frame_offset(n, 6, k) + p0 // 3
            assert out[off:off + len(want)].cpu().numpy().tobytes() == want, (p0, k)
    tail = x[n - 3000:].cpu().numpy().tobytes()
    rc_tail = oracle.revcomp_bytes(tail)  # the reverse strand STARTS with the end of the sequence
    for k in range(3):
        want = oracle.translate_bytes(1, rc_tail[k:])[:900]
        off = B.slot_frame_offset(n, 6, 3 + k)
        assert out[off:off + len(want)].cpu().numpy().tobytes() == want, k
    # ---- the same bytes as a uniform batch of 1000-nt reads: the last reads live beyond offset 2^32
    n_reads = n // 1000
    per = 16 * B.runs(1000) * 6
    B.translate_frames_uniform_dev(x, n_reads, 1000, out=out[: n_reads * per])
    assert B.dev_error_fetch(x, uniform_len=1000) is None
    o = out[: n_reads * per].view(n_reads, per)
    for i in (0, (1 << 32) // 1000, (1 << 32) // 1000 + 1, n_reads - 1):
        r = x[i * 1000:(i + 1) * 1000].cpu().numpy().tobytes()
        want = oracle.py_all_frames(1, r)
        row = o[i].cpu().numpy()
        got = [row[B.slot_frame_offset(1000, 6, f):B.slot_frame_offset(1000, 6, f) + B.frame_len(1000, f % 3)].tobytes() for f in range(6)]
        assert got == want, i


# ------------------------------------------------------------------ batched windows / all_windows shape


def test_windows_batch_and_windows_all_batch(B):
    """Batched windows (copy / strict DNA / reverse-complement windows taken from the forward strand) and the
    benches/all_windows.rs window sets for a batch, against the oracle sequence by sequence."""
    import torch

    rng = np.random.default_rng(71)
    for n_seq, L, w_dna, w_aa in ((3, 999, 42, 20), (40, 100, 42, 20), (7, 42, 42, 20), (5, 300, 17, 5), (2, 5000, 64, 33), (9, 61, 1, 1),
                                  (4, 130, 63, 64)):
        a = np.frombuffer(_rand(rng, n_seq * L, b"ACGTacgt"), dtype=np.uint8).reshape(n_seq, L).copy()
        x = torch.from_numpy(a).cuda()
        B.dev_error_reset()
        # plain copy windows (ProteinSequence::windows shape)
        got = B.windows_batch_dev(x, n_seq, L, w_aa, mode=0).cpu().numpy()
        nw = L - w_aa + 1
        for s in range(n_seq):
            assert got[s * nw:(s + 1) * nw].tobytes() == oracle.windows(a[s], w_aa).tobytes(), (n_seq, L, s)
        counts, dna_w, aa_w = B.windows_all_batch_dev(x, n_seq, L, w_dna, w_aa)
        assert B.dev_error_fetch(x, uniform_len=L, strict=True) is None
        dna_w, aa_w = dna_w.cpu().numpy(), aa_w.cpu().numpy()
        nd = counts[0]
        aa_base = 0
        want_all = [oracle.windows_all(oracle.masks_of(a[s]), w_dna, w_aa) for s in range(n_seq)]
        for s in range(n_seq):
            wd, wa, _origin, cs = want_all[s]
            assert cs[:2] == counts[:2] and [c for c in cs[2:] if c] == [c for c in counts[2:] if c]
            assert dna_w[(s * nd) * w_dna:((s + 1) * nd) * w_dna].tobytes() == wd[:nd].tobytes(), ("fwd", n_seq, L, s)
            assert dna_w[(n_seq * nd + s * nd) * w_dna:(n_seq * nd + (s + 1) * nd) * w_dna].tobytes() == wd[nd:].tobytes(), ("rc", n_seq, L, s)
        for f in range(6):
            cf = counts[2 + f]
            if not cf:
                continue
            for s in range(n_seq):
                wa, cs = want_all[s][1], want_all[s][3]
                present = [c for c in cs[2:]]
                lo = sum(present[:f])
                got_f = aa_w[(aa_base + s * cf) * w_aa:(aa_base + (s + 1) * cf) * w_aa]
                assert got_f.tobytes() == wa[lo:lo + cf].tobytes(), ("aa", f, n_seq, L, s)
            aa_base += n_seq * cf
    # masks in, and an ambiguity code is an error with its sequence and position
    a = np.frombuffer(_rand(rng, 6 * 200, b"ACGT"), dtype=np.uint8).reshape(6, 200).copy()
    m = np.stack([oracle.masks_of(a[s]) for s in range(6)])
    xm = torch.from_numpy(m).cuda()
    got = B.windows_batch_dev(xm, 6, 200, 42, mode=4).cpu().numpy()
    for s in range(6):
        rc = np.frombuffer(oracle.revcomp_bytes(a[s].tobytes()), dtype=np.uint8)
        assert got[s * 159:(s + 1) * 159].tobytes() == oracle.windows(rc, 42).tobytes()
    a[4, 77] = ord("N")
    x = torch.from_numpy(a).cuda()
    B.dev_error_reset()
    B.windows_batch_dev(x, 6, 200, 42, mode=1)
    e = B.dev_error_fetch(x, uniform_len=200, strict=True)
    assert e is not None and (e.seq_index, e.pos, e.kind, e.byte) == (4, 77, 3, ord("N"))
