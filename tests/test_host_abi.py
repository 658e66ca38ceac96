"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header
declares, host-only helpers agree with the oracle, and compute calls fail loudly without a GPU."""

import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import oracle
from tests import vectors as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def N():
    import __graft_entry__ as g

    g.build()
    from quickdna_b200 import _native

    return _native


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(N):
    header = open(os.path.join(ROOT, "include", "quickdna_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    L = N.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    assert L.qd_abi_version() == 1


def test_no_oracle_in_product():
    # the product must never route through oracle/ (or any CPU fallback)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quickdna_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "qd_oracle" not in text and "libqd_oracle" not in text, f


def test_host_lut_image_is_tables_dat(N):
    L = N.lib()
    img = C.string_at(L.qd_table_data(0), 27 * 4096)
    assert hashlib.sha256(img).hexdigest() == V.TABLES_DAT_SHA256
    assert img == oracle.tables()  # product and oracle build it with different algorithms
    assert L.qd_table_data(27) is None and L.qd_table_data(-1) is None


def test_table_slots_match_oracle(N):
    for t in range(-3, 260):
        assert N.lib().qd_table_slot(t) == oracle.table_slot(t), t


def test_error_messages_match_oracle(N):
    buf = C.create_string_buffer(128)
    for kind in (1, 2, 3, 4):
        for b in range(256 if kind in (1, 4) else 128):  # chars of kinds 2/3 are always ASCII
            e = N.QdErr(kind, b, 0, 0)
            N.lib().qd_err_message(C.byref(e), buf, 128)
            assert buf.value.decode() == oracle.err_message(kind, b), (kind, b)


def test_geometry(N):
    L = N.lib()
    from quickdna_b200 import batch

    for n in list(range(0, 200)) + [999, 1000, 1001, 29901, 250_000_000]:
        for k in range(3):
            want = (n - k) // 3 if n > k else 0
            assert L.qd_frame_len(n, k) == want == batch.frame_len(n, k)
        R = (n + 47) // 48
        assert L.qd_runs(n) == R == batch.runs(n)
        for frames in (1, 3, 6):
            assert L.qd_slots_bytes(n, frames) == 16 * R * frames
            for f in range(1 if frames == 1 else frames):
                off = L.qd_slot_frame_offset(n, frames, f)
                assert off == batch.slot_frame_offset(n, frames, f)
                ln = L.qd_frame_len(n, f % 3)
                assert off % 16 == 0 if f < 3 else (off + ln) % 16 == 0  # aligned start / aligned end
                assert off + ln <= 16 * R * frames
    offs = np.array([0, 10, 10, 1010, 1058], dtype=np.uint64)
    assert L.qd_batch_out_bytes(offs.ctypes.data_as(C.c_void_p), 4, 6) == 96 * (1 + 0 + 21 + 1)
    assert L.qd_slots_bytes(1000, 6) == 2016  # 1996 residues + 20 bytes of slot padding per 1 kb read


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(N):
    import quickdna_b200 as qd

    h = C.c_void_p()
    assert N.lib().qd_ctx_create(C.byref(h), 0) == N.QD_ERR_CUDA
    with pytest.raises(N.CudaError):
        qd.DnaSequence("AAAGGGAAA").translate()
    with pytest.raises(N.CudaError):
        qd.DnaSequence("AAAGGGAAA").reverse_complement()
    # table errors need no device, like TranslationTable::try_from
    from quickdna_b200 import quickdna as native

    with pytest.raises(ValueError) as ei:
        native._check_table(17)
    assert str(ei.value) == V.BAD_TABLE_MESSAGE


def test_python_value_semantics_need_no_gpu():
    import quickdna_b200 as qd

    D, P = qd.DnaSequence, qd.ProteinSequence
    assert D("aaa") == D(b"aaa") and D("aaa") != P("aaa") and hash(P("aaa")) == hash(P("aaa"))
    assert D("atcg")[1:3] == D("tc") and D("atcg")[0] == ord("a") and bytes(D("at")) == b"at" and str(P("KG")) == "KG"
    with pytest.raises(UnicodeEncodeError):
        D("AAčCCG")
    with pytest.raises(TypeError):
        D("at")["x"]
    assert (D("at") + "cg").seq == b"atcg" and (P("k") * 3).seq == b"kkk"


def test_synthetic_generators_agree():
    import torch

    from quickdna_b200.synth import synth_iupac_np, synth_np, synth_torch

    a = synth_np(3, 12345, 100_000)
    assert (a == synth_torch(3, 12345, 100_000, "cpu", chunk=4096).numpy()).all()
    assert (synth_np(3, 12345 + 500, 1000) == a[500:1500]).all()  # counter-based: any slice regenerates
    assert set(np.unique(a)) == set(b"ACGT")
    x = synth_iupac_np(5, 0, 200_000)
    amb = np.isin(x, np.frombuffer(b"WMRYSKBVDHNwmryskbvdhn", dtype=np.uint8)).mean()
    assert 0.04 < amb < 0.06 and 0.45 < (x >= 97).mean() < 0.55


def test_canonical_key_host_helpers():
    """qd_canonical_key_bytes / qd_canonical_fingerprint are host-only: sizes, and the documented fingerprint function."""
    import ctypes as C

    import numpy as np

    from quickdna_b200 import _native as N

    L = N.lib()
    assert [L.qd_canonical_key_bytes(N.CANON_KEY_PACKED, w) for w in (1, 16, 32, 33, 42, 64)] == [8, 8, 8, 16, 16, 16]
    assert L.qd_canonical_key_bytes(N.CANON_KEY_FP64, 42) == 8
    assert L.qd_canonical_key_bytes(N.CANON_KEY_PACKED, 0) == 0 and L.qd_canonical_key_bytes(N.CANON_KEY_PACKED, 65) == 0
    assert L.qd_canonical_key_bytes(7, 42) == 0

    def mix(z):
        z ^= z >> 30
        z = (z * 0xbf58476d1ce4e5b9) & (2**64 - 1)
        z ^= z >> 27
        z = (z * 0x94d049bb133111eb) & (2**64 - 1)
        return z ^ (z >> 31)

    rng = np.random.default_rng(5)
    seen = set()
    for w in (1, 5, 16, 17, 32, 33, 42, 64):
        pw = (w + 31) // 32
        for _ in range(50):
            y = rng.integers(0, 2**32, 2 * pw, dtype=np.uint64)  # lo words, then hi words
            h = (0x9e3779b97f4a7c15 * w) & (2**64 - 1)
            for k in range(pw):
                h = mix(h ^ (int(y[k]) | int(y[pw + k]) << 32))
            arr = y.astype(np.uint32)
            got = L.qd_canonical_fingerprint(arr.ctypes.data_as(C.c_void_p), w)
            assert got == h, (w, y)
            seen.add(got)
    assert len(seen) == 8 * 50
