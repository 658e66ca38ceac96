"""Pins the CPU oracle (oracle/qd_oracle.c) to the reference's golden artefacts.

Runs without a GPU.  The container-only tests additionally read /root/reference to compare
against src/tables.dat byte-for-byte, replay the asserts in the reference's own Rust test
modules, and check the covid.txt pins.
"""

import hashlib
import itertools
import os
import re

import numpy as np
import pytest

import oracle
from tests import vectors as V


def test_lut_sha256_matches_tables_dat():
    t = oracle.tables()
    assert len(t) == 27 * 4096
    assert hashlib.sha256(t).hexdigest() == V.TABLES_DAT_SHA256


def test_lut_bytes_equal_reference_file(reference_dir):
    with open(os.path.join(reference_dir, "src", "tables.dat"), "rb") as f:
        assert f.read() == oracle.tables()


def test_table_slots():
    # src/trans_table.rs:114-146, 230-267
    assert oracle.table_slot(1) == 0 and oracle.table_slot(8) == 0
    assert oracle.table_slot(4) == 3 and oracle.table_slot(7) == 3
    assert oracle.table_slot(11) == 8 and oracle.table_slot(33) == 26
    for bad in (0, 17, 18, 19, 20, 34, 255):
        assert oracle.table_slot(bad) == -1
    slots = sorted({oracle.table_slot(t) for t in V.ALL_ACCEPTED_TABLES})
    assert slots == list(range(27))
    t = oracle.tables()
    assert t[0:4096] == t[8 * 4096:9 * 4096]  # NCBI 1 == NCBI 11 (start codons ignored)
    assert t[20 * 4096:21 * 4096] == t[21 * 4096:22 * 4096]


@pytest.mark.parametrize("table,dna,protein", V.TRANSLATE)
def test_translate_vectors(table, dna, protein):
    assert oracle.translate_bytes(table, dna) == protein.encode()
    if all(c in "ACGTacgt" for c in dna):
        assert oracle.translate_bytes(table, dna, strict=True) == protein.encode()
    if "Q" not in dna:
        m = oracle.parse(dna)
        assert oracle.translate_masks(oracle.table_slot(table), m) == protein.encode()


@pytest.mark.parametrize("dna,frames", V.SELF_FRAMES)
def test_self_frames_vectors(dna, frames):
    want = [f.encode() for f in frames]
    assert oracle.self_frames_masks(0, oracle.parse(dna)) == want
    assert oracle.py_self_frames(1, dna) == want
    assert oracle.py_self_frames(1, dna, strict=True) == want


@pytest.mark.parametrize("dna,frames", V.ALL_FRAMES)
def test_all_frames_vectors(dna, frames):
    want = [f.encode() for f in frames]
    assert oracle.all_frames_masks(0, oracle.parse(dna)) == want
    assert oracle.py_all_frames(1, dna) == want
    assert oracle.py_all_frames(1, dna, strict=True) == want


def test_lazy_frame_structure():
    # src/iter.rs:21-74 -- codon lists of the six frames of CGATCGAT and of its first 4 nt
    def codon_frames(dna):
        n = len(dna)
        rc = oracle.revcomp_bytes(dna).decode()
        out = []
        for s in (dna, rc):
            for k in range(3):
                cods = [s[i:i + 3] for i in range(k, n - 2, 3)]
                if cods:
                    out.append(cods)
        return out

    assert codon_frames(V.ITER_FRAMES_DNA) == V.ITER_FRAMES_CODONS
    assert codon_frames(V.ITER_FRAMES_DNA[:4]) == V.ITER_FRAMES_DNA4_CODONS
    assert codon_frames(V.ITER_FRAMES_DNA[:2]) == []
    # and the aa frames agree with translating those codons one by one
    frames = oracle.all_frames_masks(0, oracle.parse(V.ITER_FRAMES_DNA))
    for fr, cods in zip(frames, V.ITER_FRAMES_CODONS):
        assert fr == b"".join(oracle.translate_bytes(1, c) for c in cods)


@pytest.mark.parametrize("dna,rc", V.REVCOMP)
def test_revcomp_vectors(dna, rc):
    assert oracle.revcomp_bytes(dna) == rc.encode()
    assert oracle.ascii_of(oracle.revcomp_masks(oracle.parse(dna))) == rc.encode()


@pytest.mark.parametrize("dna,aa", V.RC_FRAME0)
def test_rc_frame0(dna, aa):
    assert oracle.translate_bytes(1, oracle.revcomp_bytes(dna)) == aa.encode()
    assert oracle.all_frames_masks(0, oracle.parse(dna))[3] == aa.encode()


@pytest.mark.parametrize("fn,dna,kind,byte,msg", V.ERRORS)
def test_error_vectors(fn, dna, kind, byte, msg):
    raw = dna.encode("latin-1")
    call = {
        "translate": lambda: oracle.translate_bytes(1, raw),
        "translate_strict": lambda: oracle.translate_bytes(1, raw, strict=True),
        "revcomp": lambda: oracle.revcomp_bytes(raw),
        "revcomp_strict": lambda: oracle.revcomp_bytes(raw, strict=True),
    }[fn]
    with pytest.raises(oracle.OracleError) as ei:
        call()
    assert (ei.value.kind, ei.value.byte, str(ei.value)) == (kind, byte, msg)


def test_bad_table_checked_before_dna():
    # src/python_api.rs:35, 48: table error first, even for invalid DNA
    with pytest.raises(oracle.OracleError) as ei:
        oracle.translate_bytes(17, b"QQQ")
    assert ei.value.kind == 4 and str(ei.value) == V.BAD_TABLE_MESSAGE
    for t in (0, 17, 18, 19, 20, 34, 200):
        with pytest.raises(oracle.OracleError):
            oracle.translate_bytes(t, b"AAA")
    for t in V.ALL_ACCEPTED_TABLES:
        assert len(oracle.translate_bytes(t, b"AAA")) == 1


def test_error_message_formats():
    # src/errors.rs:19-28; Rust {:?} escaping of chars, {:x?} of bytes
    assert oracle.err_message(1, 0xC4) == "non-ascii byte: c4"
    assert oracle.err_message(1, 0x80) == "non-ascii byte: 80"
    assert oracle.err_message(2, ord("x")) == "bad nucleotide: 'x'"
    assert oracle.err_message(2, ord("\n")) == "bad nucleotide: '\\n'"
    assert oracle.err_message(2, ord("\t")) == "bad nucleotide: '\\t'"
    assert oracle.err_message(2, ord("'")) == "bad nucleotide: '\\''"
    assert oracle.err_message(2, ord("\\")) == "bad nucleotide: '\\\\'"
    assert oracle.err_message(2, 0) == "bad nucleotide: '\\0'"
    assert oracle.err_message(2, 1) == "bad nucleotide: '\\u{1}'"
    assert oracle.err_message(2, 0x7F) == "bad nucleotide: '\\u{7f}'"
    assert oracle.err_message(2, ord('"')) == "bad nucleotide: '\"'"
    assert oracle.err_message(3, ord("N")) == "unexpected ambiguous nucleotide: 'N'"
    assert oracle.err_message(4, 17) == "not a ncbi translation table: 17"


def test_accepted_alphabet():
    # src/rust_api.rs:421-457: single-character parses over all of ASCII, both alphabets
    for c in range(128):
        ch = bytes([c])
        ok_amb = chr(c) in V.ACCEPTED_AMBIGUOUS + " \t"
        ok_strict = chr(c) in V.ACCEPTED_STRICT + " \t"
        for strict, ok in ((False, ok_amb), (True, ok_strict)):
            if ok:
                oracle.parse(ch, strict=strict)
            else:
                with pytest.raises(oracle.OracleError):
                    oracle.parse(ch, strict=strict)
        # the bytes path (python_api) accepts no whitespace at all
        if chr(c) in " \t":
            with pytest.raises(oracle.OracleError):
                oracle.revcomp_bytes(ch)
    for c in range(128, 256):
        with pytest.raises(oracle.OracleError) as ei:
            oracle.revcomp_bytes(bytes([c]))
        assert ei.value.kind == 1 and ei.value.byte == c


def test_parse_skips_space_and_tab_only():
    # src/rust_api.rs:771-785
    assert oracle.ascii_of(oracle.parse("  gcantac\tctaangtnattag ")) == b"GCANTACCTAANGTNATTAG"
    assert oracle.ascii_of(oracle.parse(" gca ctac ctaacgtcattag \t", strict=True)) == b"GCACTACCTAACGTCATTAG"
    with pytest.raises(oracle.OracleError):
        oracle.parse("AC\nGT")


def test_first_error_position_follows_loop_order():
    # derived from the loop order of src/trans_table.rs:191-194 and :273-276
    with pytest.raises(oracle.OracleError) as ei:
        oracle.translate_bytes(1, b"ACGTNQ", strict=True)
    assert (ei.value.pos, ei.value.kind, ei.value.byte) == (4, 3, ord("N"))
    with pytest.raises(oracle.OracleError) as ei:
        oracle.translate_bytes(1, b"ACGTNQ")
    assert (ei.value.pos, ei.value.kind) == (5, 2)
    assert oracle.translate_bytes(1, b"ACGTNNQ"[:7][:6] + b"Q") == b"TX"  # 7th byte is in the ignored tail
    with pytest.raises(oracle.OracleError) as ei:
        oracle.revcomp_bytes(b"ACGTACGTQ")
    assert ei.value.pos == 8
    # python all-frames: lowest failing index of [0, n) for n >= 3
    with pytest.raises(oracle.OracleError) as ei:
        oracle.py_all_frames(1, b"ACGTACGQ")
    assert ei.value.pos == 7 and ei.value.byte == ord("Q")
    # n < 3: only the NON-strict reverse complement looks at the bytes (quickdna/__init__.py:182)
    with pytest.raises(oracle.OracleError):
        oracle.py_all_frames(1, b"Q")
    assert oracle.py_all_frames(1, b"NN", strict=True) == []
    assert oracle.py_self_frames(1, b"QQ") == []


def test_exhaustive_codons_match_lut_rule():
    # tests/test_exhaustive.py shape: all 15^3 codons x every accepted table go through the
    # bytes path and equal the LUT entry; the LUT itself is pinned by sha256 above.
    letters = "ATCGWMRYSKBVDHN"
    t = oracle.tables()
    for table in V.ALL_ACCEPTED_TABLES:
        slot = oracle.table_slot(table)
        dna = "".join(a + b + c for a, b, c in itertools.product(letters, repeat=3))
        got = oracle.translate_bytes(table, dna)
        m = oracle.masks_of(dna).reshape(-1, 3).astype(np.int64)
        idx = (m[:, 0] << 8) | (m[:, 1] << 4) | m[:, 2]
        want = bytes(t[slot * 4096 + i] for i in idx)
        assert got == want
        assert set(got) <= set(b"*ABCDEFGHIJKLMNPQRSTVWXYZ")
        assert oracle.translate_bytes(table, dna.lower()) == got


@pytest.mark.parametrize("dna,canon", V.CANONICAL)
def test_canonical_vectors(dna, canon):
    assert oracle.ascii_of(oracle.canonical(oracle.parse(dna, strict=True))) == canon.encode()


@pytest.mark.parametrize("dna,canon", V.FORWARD_CANONICAL)
def test_forward_canonical_vectors(dna, canon):
    assert oracle.ascii_of(oracle.forward_canonical(oracle.parse(dna, strict=True))) == canon.encode()


def test_small_canonical_exhaustive():
    for n in (1, 2, 3):
        inputs = ["".join(p) for p in itertools.product("ATCG", repeat=n)]
        fw, cn = V.SMALL_FW[n].split(), V.SMALL_CANON[n].split()
        assert len(inputs) == len(fw) == len(cn)
        for s, f, c in zip(inputs, fw, cn):
            m = oracle.parse(s, strict=True)
            assert oracle.ascii_of(oracle.forward_canonical(m)) == f.encode(), s
            assert oracle.ascii_of(oracle.canonical(m)) == c.encode(), s


def test_canonical_asserts_in_reference_source(reference_dir):
    src = open(os.path.join(reference_dir, "src", "canonical.rs")).read()
    pairs_fw = re.findall(r'assert_eq!\(\s*fw_canon\("([ATCG]*)"\),\s*"([ATCG]*)"\s*\)', src)
    pairs_cn = re.findall(r'assert_eq!\(\s*canon\("([ATCG]*)"\),\s*"([ATCG]*)"\s*\)', src)
    assert len(pairs_fw) >= 86 and len(pairs_cn) >= 88
    for s, want in pairs_fw:
        assert oracle.ascii_of(oracle.forward_canonical(oracle.parse(s, strict=True))) == want.encode(), s
    for s, want in pairs_cn:
        assert oracle.ascii_of(oracle.canonical(oracle.parse(s, strict=True))) == want.encode(), s


def test_canonical_properties():
    # src/canonical.rs:414-437 (quickcheck properties) on seeded random strict DNA
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(0, 80))
        m = np.array([1, 2, 4, 8], dtype=np.uint8)[rng.integers(0, 4, n)]
        fw = oracle.forward_canonical(m)
        assert (oracle.forward_canonical(fw) == fw).all()
        cn = oracle.canonical(m)
        assert (oracle.canonical(cn) == cn).all()
        assert (oracle.canonical(oracle.revcomp_masks(m)) == cn).all()
        # LexicalMin == Vec::min under A<T<C<G, i.e. numeric order of the masks
        rev_fw = oracle.forward_canonical(m[::-1].copy())
        assert bytes(cn) == min(bytes(fw), bytes(rev_fw))


@pytest.mark.parametrize("dna,exp", V.EXPANSIONS)
def test_expansion_vectors(dna, exp):
    m = oracle.parse(dna)
    got = [oracle.ascii_of(r).decode() for r in oracle.expansions(m)]
    assert got == exp
    assert oracle.expansions_size_hint(m) == len(exp)


def test_expansion_size_hints():
    for dna, n in V.SIZE_HINTS:
        m = oracle.parse(dna)
        assert oracle.expansions_size_hint(m) == n
        assert len(oracle.expansions(m)) == n
    assert oracle.expansions_size_hint(oracle.parse(V.SIZE_HINT_BIG[0])) == V.SIZE_HINT_BIG[1]
    assert oracle.expansions_size_hint(oracle.parse(V.SIZE_HINT_OVERFLOW)) is None


def test_expansions_are_lexicographic():
    rng = np.random.default_rng(11)
    for _ in range(100):
        n = int(rng.integers(1, 12))
        m = rng.integers(1, 16, n).astype(np.uint8)
        hint = oracle.expansions_size_hint(m)
        if hint > 4096:
            continue
        ex = oracle.expansions(m)
        assert len(ex) == hint
        rows = [bytes(r) for r in ex]
        assert rows == sorted(set(rows))  # strictly increasing under A<T<C<G (numeric mask order)
        for r in ex:
            assert ((r & m) == r).all() and all(bin(x).count("1") == 1 for x in r)


def test_windows_vectors():
    seq, wins = V.WINDOWS_10
    got = [bytes(w).decode() for w in oracle.windows(seq, 10)]
    assert got == wins
    assert len(oracle.windows("antg", 10)) == 0


def test_windows_all_matches_iterator_formulation():
    # benches/all_windows.rs:221-226 self-check: SequenceWindows == IterBasedSequenceWindows.
    rng = np.random.default_rng(3)
    for n in (0, 1, 4, 5, 41, 42, 43, 59, 60, 61, 62, 63, 64, 65, 200, 999):
        m = np.array([1, 2, 4, 8], dtype=np.uint8)[rng.integers(0, 4, n)]
        dna, aa, origin, counts = oracle.windows_all(m, 42, 20)
        asc = oracle.ascii_of(m)
        rc = oracle.revcomp_bytes(asc)
        frames = oracle.all_frames_masks(0, m)

        def nwin(length, w):
            return length + 1 - min(w, length + 1)

        exp = [(i, asc[i:i + 42]) for i in range(nwin(n, 42))]
        first = nwin(n, 42) - 1
        exp += [(first - i, rc[i:i + 42]) for i in range(nwin(n, 42))]
        for f, fr in enumerate(frames):
            for i in range(nwin(len(fr), 20)):
                idx = n - ((i + 20) * 3 + f % 3) if f >= 3 else 3 * i + f % 3
                exp.append((idx, fr[i:i + 20]))
        got = [(int(o), bytes(w)) for o, w in zip(origin[: len(dna)], dna)]
        got += [(int(o), bytes(w)) for o, w in zip(origin[len(dna):], aa)]
        assert got == exp
    # window census at the bench length
    m = np.array([1, 2, 4, 8], dtype=np.uint8)[rng.integers(0, 4, 99999)]
    _, _, origin, counts = oracle.windows_all(m, 42, 20)
    assert counts[:2] == [99958, 99958] and sum(counts[2:]) == 199880 and len(origin) == 399796


def test_covid_pins(reference_dir):
    raw = open(os.path.join(reference_dir, "benchmarks", "covid.txt"), "rb").read().replace(b"\n", b"")
    P = V.COVID
    assert len(raw) == P["len"] and hashlib.sha256(raw).hexdigest().startswith(P["sha256_prefix"])
    t1 = oracle.translate_bytes(1, raw)
    assert hashlib.sha256(t1).hexdigest() == P["translate_t1_sha256"]
    assert t1.decode().startswith(P["translate_t1_head"])
    assert oracle.translate_bytes(11, raw) == t1
    assert hashlib.sha256(oracle.translate_bytes(2, raw)).hexdigest().startswith(P["translate_t2_sha256_prefix"])
    assert hashlib.sha256(oracle.translate_bytes(22, raw)).hexdigest().startswith(P["translate_t22_sha256_prefix"])
    assert hashlib.sha256(oracle.revcomp_bytes(raw)).hexdigest() == P["revcomp_sha256"]
    frames = oracle.py_all_frames(1, raw)
    assert [len(f) for f in frames] == P["frame_lens"]
    assert hashlib.sha256(b"\n".join(frames)).hexdigest() == P["frames_joined_sha256"]
    assert frames == oracle.all_frames_masks(0, oracle.parse(raw))
