"""Property tests with `hypothesis`, after the reference's tests/test_hypothesis.py (random strings over
aAtTcCgGnN, all 23 tables Biopython knows) and the quickcheck properties of src/canonical.rs:414-437.

CPU: properties of the oracle itself.  GPU (-m gpu): the CUDA path, through the Python drop-in surface, against the
oracle on every generated example."""

import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle

# the reference's profiles (tests/conftest.py:8-11): ci = 2000 examples, dev = 200
EXAMPLES = {"ci": 2000, "dev": 200}.get(os.environ.get("HYPOTHESIS_PROFILE", "dev"), 200)
COMMON = dict(max_examples=EXAMPLES, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])

st_dna = st.text(alphabet="aAtTcCgGnN", max_size=600)          # tests/test_hypothesis.py:19
st_strict = st.text(alphabet="ATCG", min_size=0, max_size=200)  # DnaSequenceStrict of src/quickcheck.rs
st_iupac = st.text(alphabet="ACGTacgtWMRYSKBVDHNwmryskbvdhn", max_size=600)


def valid_tables():  # tests/utils.py:7-12
    return [t for t in range(1, 34) if t not in (7, 8, 17, 18, 19, 20)]


# ------------------------------------------------------------------ oracle properties (CPU)


@settings(**COMMON)
@given(st_strict)
def test_forward_canonicalization_is_idempotent(dna):
    m = oracle.masks_of(dna.encode()) if dna else np.zeros(0, np.uint8)
    c = oracle.forward_canonical(m)
    assert oracle.forward_canonical(c).tobytes() == c.tobytes()


@settings(**COMMON)
@given(st_strict)
def test_canonicalization_is_idempotent_and_strand_symmetric(dna):
    m = oracle.masks_of(dna.encode()) if dna else np.zeros(0, np.uint8)
    c = oracle.canonical(m)
    assert oracle.canonical(c).tobytes() == c.tobytes()
    assert oracle.canonical(oracle.revcomp_masks(m)).tobytes() == c.tobytes()
    # LexicalMin == min of the two relabelled sequences (A < T < C < G is the order of the mask values 1, 2, 4, 8)
    fw, rv = oracle.forward_canonical(m), oracle.forward_canonical(m[::-1].copy())
    assert c.tobytes() == min(fw.tobytes(), rv.tobytes())


@settings(**COMMON)
@given(st_dna)
def test_oracle_frames_are_consistent(dna):
    raw = dna.encode()
    frames = oracle.py_all_frames(1, raw) if len(raw) >= 3 else []
    present = [k for k in range(3) if len(raw) > k and (len(raw) - k) // 3 > 0]
    assert len(frames) == 2 * len(present)
    rc = oracle.revcomp_bytes(raw)
    for j, k in enumerate(present):
        assert frames[j] == oracle.translate_bytes(1, raw[k:])
        assert frames[len(present) + j] == oracle.translate_bytes(1, rc[k:])
    assert oracle.revcomp_bytes(rc) == raw.upper()


# ------------------------------------------------------------------ CUDA path vs oracle (GPU)


@pytest.mark.gpu
@settings(**COMMON)
@given(st_dna)
def test_translate_like_reference_hypothesis(dna):
    import quickdna_b200 as qd

    raw = dna.encode()
    for table in valid_tables():
        assert qd.DnaSequence(dna).translate(table).seq == oracle.translate_bytes(table, raw)


@pytest.mark.gpu
@settings(**COMMON)
@given(st_dna)
def test_reverse_complement_like_reference_hypothesis(dna):
    import quickdna_b200 as qd

    assert qd.DnaSequence(dna).reverse_complement().seq == oracle.revcomp_bytes(dna.encode())


@pytest.mark.gpu
@settings(**COMMON)
@given(st_iupac, st.sampled_from([1, 2, 11, 22, 33]), st.booleans())
def test_all_frames_and_strictness(dna, table, strict):
    import quickdna_b200 as qd

    raw = dna.encode()
    try:
        if len(raw) >= 3:
            want = oracle.py_all_frames(table, raw, strict=strict)
        else:  # the wrapper still validates through reverse_complement(), non-strictly (quickdna/__init__.py:182)
            oracle.revcomp_bytes(raw)
            want = []
        werr = None
    except oracle.OracleError as e:
        want, werr = None, e
    if werr is None:
        assert [f.seq for f in qd.DnaSequence(dna).translate_all_frames(table, strict)] == want
    else:
        with pytest.raises(ValueError) as ei:
            qd.DnaSequence(dna).translate_all_frames(table, strict)
        assert (ei.value.kind, ei.value.byte, str(ei.value)) == (werr.kind, werr.byte, str(werr))


@pytest.mark.gpu
@settings(**COMMON)
@given(st_strict)
def test_canonical_gpu_properties(dna):
    from quickdna_b200 import batch as B

    raw = dna.encode()
    m = oracle.masks_of(raw) if raw else np.zeros(0, np.uint8)
    got = B.canonical(raw).tobytes()
    assert got == (oracle.ascii_of(oracle.canonical(m)) if raw else b"")
    assert B.canonical(got).tobytes() == got  # idempotent
    assert B.canonical(oracle.revcomp_bytes(raw)).tobytes() == got  # strand symmetric
    if len(raw) >= 20:
        w = 20
        assert B.canonical_windows(raw, w).tobytes() == oracle.ascii_of(oracle.canonical_windows(m, w).reshape(-1))


@pytest.mark.gpu
@settings(**COMMON)
@given(st.text(alphabet="ATCG", min_size=1, max_size=300), st.integers(min_value=1, max_value=64), st.booleans())
def test_canonical_window_keys_hypothesis(dna, w, forward_only):
    """One key per canonical window: unpacking the bit planes gives the oracle's canonical windows,
    for every window length the kernel takes; equal windows have equal fingerprints and (w <= 32) vice versa."""
    from quickdna_b200 import _native as N
    from quickdna_b200 import batch as B

    raw = dna.encode()
    m = oracle.masks_of(raw)
    count = max(0, len(raw) - w + 1)
    packed = B.canonical_window_keys(raw, w=w, kind=N.CANON_KEY_PACKED, forward_only=forward_only)
    fp = B.canonical_window_keys(raw, w=w, kind=N.CANON_KEY_FP64, forward_only=forward_only)
    pw = (w + 31) // 32
    assert packed.shape == (count, 2 * pw) and fp.shape == (count,)
    if count == 0:
        return
    want = (oracle.canonical_windows(m, w) if not forward_only
            else np.stack([oracle.forward_canonical(m[i:i + w]) for i in range(count)]))
    bits = (packed[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1   # [count, 2 pw, 32]
    lo = bits[:, :pw].reshape(count, -1)[:, :w]
    hi = bits[:, pw:].reshape(count, -1)[:, :w]
    mask_of_code = np.array([1, 2, 4, 8], dtype=np.uint8)  # A, T, C, G
    assert (mask_of_code[lo + 2 * hi] == want).all()
    assert (bits[:, :pw].reshape(count, -1)[:, w:] == 0).all() and (bits[:, pw:].reshape(count, -1)[:, w:] == 0).all()
    same_window = (want[:, None, :] == want[None, :, :]).all(axis=2)
    same_key = fp[:, None] == fp[None, :]
    assert (same_key[same_window]).all()
    if w <= 32:
        assert (same_window == same_key).all()
