"""Committed golden fixtures (tests/golden/hotpath_golden.json, made by tests/golden/make_golden.py in the build
container).  CPU: the oracle still reproduces them.  GPU (-m gpu): the CUDA path, through the C ABI, reproduces them."""

import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "hotpath_golden.json")) as f:
    GOLDEN = json.load(f)
CASES = GOLDEN["cases"]
IDS = [f"{c['kind']}-{c['n']}" for c in CASES]


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def _input(case) -> bytes:
    from quickdna_b200.synth import synth_iupac_np, synth_np

    dna = (synth_np(case["seed"], 0, case["n"]) if case["kind"] == "acgt" else synth_iupac_np(case["seed"], 0, case["n"])).tobytes()
    assert sha(dna) == case["input_sha256"]
    return dna


def _same(case, key, got: bytes):
    want = case[key]
    assert (got.decode() if case["n"] <= 64 else sha(got)) == want, key


def _check(case, translate, all_frames, revcomp, canonical_windows, canonical, strict_translate):
    dna = _input(case)
    for table in (1, 2, 11, 22, 33):
        _same(case, f"translate_t{table}", translate(table, dna))
    frames = all_frames(dna)
    assert [len(f) for f in frames] == case["all_frames_lens"]
    if case["n"] <= 64:
        assert [f.decode() for f in frames] == case["all_frames_t1"]
    else:
        assert sha(b"\n".join(frames)) == case["all_frames_t1"]
    _same(case, "revcomp", revcomp(dna))
    if "canonical_windows_42" in case:
        assert sha(canonical_windows(dna)) == case["canonical_windows_42"]
        assert sha(canonical(dna)) == case["canonical"]
    if "strict_error" in case:
        want = case["strict_error"]
        try:
            strict_translate(dna)
            got = None
        except ValueError as e:
            got = {"kind": e.kind, "byte": e.byte, "pos": e.pos, "message": str(e)}
        assert got == want


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_reproduces_golden(case):
    import oracle

    assert sha(oracle.tables()) == GOLDEN["oracle_tables_sha256"]
    _check(case,
           translate=lambda t, d: oracle.translate_bytes(t, d),
           all_frames=lambda d: oracle.py_all_frames(1, d) if len(d) >= 3 else [],
           revcomp=lambda d: oracle.revcomp_bytes(d),
           canonical_windows=lambda d: oracle.canonical_windows(oracle.masks_of(d), 42).tobytes(),
           canonical=lambda d: oracle.canonical(oracle.masks_of(d)).tobytes(),
           strict_translate=lambda d: oracle.translate_bytes(1, d, strict=True))


def test_covid_pins(reference_dir):
    """BASELINE config 1 input: hashes independently derived in SURVEY.md 8c."""
    import oracle

    cov = GOLDEN["covid"]
    with open(os.path.join(reference_dir, "benchmarks", "covid.txt"), "rb") as f:
        genome = f.read().replace(b"\n", b"").replace(b"\r", b"")
    assert len(genome) == cov["n"] == 29901 and sha(genome) == cov["input_sha256"]
    assert cov["translate_t1"] == "e4c87be3e601ee04156595fc5a4cada4c9ed35af8047f4805c1e4242d889baf3"
    assert cov["revcomp"] == "cb628e8bbfb51a5cfb564157d98e9127ef1160ebfd362c3ee0a7ad2f1a441255"
    assert cov["all_frames_t1"] == "c050477ed99d645826c4e073904b843ee5fbf723a38dc7bc680c0518f4d75545"
    assert sha(oracle.translate_bytes(1, genome)) == cov["translate_t1"] == sha(oracle.translate_bytes(11, genome))
    assert sha(oracle.translate_bytes(2, genome)) == cov["translate_t2"]
    assert sha(oracle.translate_bytes(22, genome)) == cov["translate_t22"]
    assert sha(oracle.revcomp_bytes(genome)) == cov["revcomp"]
    frames = oracle.py_all_frames(1, genome)
    assert [len(x) for x in frames] == cov["all_frames_lens"] and sha(b"\n".join(frames)) == cov["all_frames_t1"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_cuda_path_reproduces_golden(case):
    import quickdna_b200 as qd
    from quickdna_b200 import batch as B
    from quickdna_b200 import quickdna as native

    to_mask = np.zeros(256, dtype=np.uint8)
    for ch, m in zip(b"ATCG", (1, 2, 4, 8)):
        to_mask[ch] = m

    def canonical_windows(d):  # golden holds masks (the oracle's typed form)
        return B.canonical_windows(to_mask[np.frombuffer(d, dtype=np.uint8)], 42, enc=1).tobytes()

    _check(case,
           translate=lambda t, d: native._translate(t, d),
           all_frames=lambda d: [f.seq for f in qd.DnaSequence(d).translate_all_frames(1)],
           revcomp=lambda d: native._reverse_complement(d),
           canonical_windows=canonical_windows,
           canonical=lambda d: B.canonical(to_mask[np.frombuffer(d, dtype=np.uint8)], enc=1).tobytes(),
           strict_translate=lambda d: native._translate_strict(1, d))
