"""Long-sequence pipeline and the several-GPUs-from-one-call entry points.  Needs a B200: -m gpu.

The partition logic is exercised on ONE GPU too (several contexts of the same device take the places of several
devices: the arithmetic is the same); with two or more GPUs visible the same tests run across devices."""

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


@pytest.fixture(scope="module")
def B():
    from quickdna_b200 import batch

    return batch


@pytest.fixture(scope="module")
def N():
    from quickdna_b200 import _native

    return _native


@pytest.fixture(scope="module", params=["same-device", "across-devices"])
def ctx_sets(request, torch, N):
    """[[contexts]] of 2, 3 and (if present) all devices."""
    n_dev = torch.cuda.device_count()
    if request.param == "same-device":
        made = [N.Context(0) for _ in range(3)]
        yield [made[:2], made[:3]]
        for c in made:
            c.close()
    else:
        if n_dev < 2:
            pytest.skip("one GPU visible")
        sets = [[N.default_context(d) for d in range(k)] for k in sorted({2, min(4, n_dev), n_dev})]
        yield sets


def _frames_oracle(dna: bytes, frames: int, table=1, strict=False):
    if frames == 1:
        out = oracle.translate_bytes(table, dna, strict)
        return [out] if out else []
    if frames == 3:
        return oracle.py_self_frames(table, dna, strict)
    return oracle.py_all_frames(table, dna, strict)


@pytest.mark.parametrize("n", [8 << 20, (8 << 20) + 1, (9 << 20) + 47, 20_000_003])
def test_long_single_sequence_calls_are_pipelined_and_exact(B, N, n):
    """qd_translate_frames / qd_revcomp above 8 MiB go through chunks with a 2-nt halo: every frame byte equals the oracle."""
    dna = oracle.synth(2, 5, n).tobytes()
    ctx = N.default_context(0)
    before = ctx.launch_count
    for frames in (6, 3, 1):
        assert B.translate_frames(dna, frames=frames) == _frames_oracle(dna, frames), (n, frames)
    assert ctx.launch_count - before >= 3 * 2  # several chunks per call
    assert B.revcomp(dna).tobytes() == oracle.revcomp_bytes(dna)


def test_long_sequence_chunk_size_sweep(B, N):
    ctx = N.Context(0)
    try:
        n = 3_000_017
        dna = oracle.synth(5, 0, n, b"ACGTacgtNRYKMSWBDHV").tobytes()
        want6, want_rc = oracle.py_all_frames(11, dna), oracle.revcomp_bytes(dna)
        for chunk in (1 << 16, 65536 + 48, 100_000, 1 << 20):
            N.lib().qd_ctx_set_chunk_bytes(ctx.handle, chunk)
            # force the pipelined path for a 3 MB sequence: chunk = min(chunk_bytes, max(4 MiB, n / 16)) -> chunk_bytes
            assert B.translate_frames(dna, table=11, frames=6, ctx=ctx) == want6
            assert B.revcomp(dna, ctx=ctx).tobytes() == want_rc
    finally:
        ctx.close()


def test_long_sequence_first_error_is_global(B, N):
    n = 10_000_000
    dna = bytearray(oracle.synth(2, 0, n).tobytes())
    for pos, b in ((9_999_999, ord("!")), (7_000_001, ord("R")), (4_194_304 + 47, 0xC4), (4_194_304 * 2 - 1, ord("x"))):
        dna[pos] = b
        for fn, strict in ((lambda s: B.translate_frames(bytes(dna), frames=6, strict=s), True), (lambda s: B.revcomp(bytes(dna), strict=s), True)):
            with pytest.raises(N.TranslationError) as ei:
                fn(strict)
            try:
                oracle.revcomp_bytes(bytes(dna), strict=True)
                raise AssertionError("oracle accepted it")
            except oracle.OracleError as o:
                assert (ei.value.kind, ei.value.byte, ei.value.pos, str(ei.value)) == (o.kind, o.byte, o.pos, str(o)), pos


def test_multi_batch_equals_single_context(B, N, ctx_sets):
    rng = np.random.default_rng(11)
    lens = np.concatenate([rng.integers(0, 60, 300), rng.integers(500, 1500, 4000), [0, 0, 250_000, 1, 2, 3]])
    rng.shuffle(lens)
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"ACGTacgtNRY", dtype=np.uint8)[rng.integers(0, 11, int(offs[-1]))]
    uni = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (20_011, 1000))]
    for ctxs in ctx_sets:
        for frames in (6, 3, 1):
            one = B.translate_frames_batch(flat, offsets=offs, frames=frames, table=11)
            many = B.translate_frames_batch(flat, offsets=offs, frames=frames, table=11, devices=ctxs)
            assert (one.out_offsets == many.out_offsets).all() and one.data.tobytes() == many.data.tobytes(), (len(ctxs), frames)
        one = B.translate_frames_batch(uni, frames=6)
        many = B.translate_frames_batch(uni, frames=6, devices=ctxs)
        assert one.data.tobytes() == many.data.tobytes()
        for i in (0, 5003, 10_005, 20_010):  # reads on both sides of every cut
            assert many.sequence_frames(i) == oracle.py_all_frames(1, uni[i].tobytes())
        # more devices than sequences; empty batch
        tiny = B.translate_frames_batch([b"ACGTAC"], frames=6, devices=ctxs)
        assert tiny.sequence_frames(0) == oracle.py_all_frames(1, b"ACGTAC")
        assert B.translate_frames_batch([], frames=6, devices=ctxs).data.size == 0


def test_multi_batch_reports_the_first_error_in_input_order(B, N, ctx_sets):
    rng = np.random.default_rng(12)
    reads = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (9000, 300))].copy()
    reads[8000, 17] = ord("N")   # lands on the last device
    reads[2500, 250] = ord("!")  # ... and this one on an earlier device: it is the one a serial pass hits first
    for ctxs in ctx_sets:
        with pytest.raises(N.TranslationError) as ei:
            B.translate_frames_batch(reads, frames=6, strict=True, devices=ctxs)
        e = ei.value
        assert (e.seq_index, e.pos, e.kind, str(e)) == (2500, 250, N.QD_ERR_BAD_NUCLEOTIDE, "bad nucleotide: '!'")
        reads2 = reads.copy()
        reads2[2500, 250] = ord("A")
        with pytest.raises(N.TranslationError) as ei:
            B.translate_frames_batch(reads2, frames=6, strict=True, devices=ctxs)
        assert (ei.value.seq_index, ei.value.pos, ei.value.kind) == (8000, 17, N.QD_ERR_UNEXPECTED_AMBIGUOUS)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 47, 48, 49, 95, 97, 1000, 144_001, 12_000_005])
def test_multi_chromosome_equals_oracle(B, N, ctx_sets, n):
    dna = oracle.synth(5, 3, n, b"ACGTacgtNRYKM").tobytes()
    for ctxs in ctx_sets:
        for frames in (6, 3, 1):
            assert B.translate_frames(dna, table=2, frames=frames, devices=ctxs) == _frames_oracle(dna, frames, table=2), (n, len(ctxs), frames)
        assert B.revcomp(dna, devices=ctxs).tobytes() == oracle.revcomp_bytes(dna)


def test_multi_chromosome_first_error(B, N, ctx_sets):
    n = 1_000_003
    base = bytearray(oracle.synth(2, 0, n).tobytes())
    for pos in (0, 333_360, 333_361, 999_999, n - 1):
        dna = bytearray(base)
        dna[pos] = ord("?")
        if pos > 600_000:
            dna[pos - 5] = ord("R")  # fine unless strict
        for ctxs in ctx_sets:
            for strict in (False, True):
                for call in (lambda: B.translate_frames(bytes(dna), frames=6, strict=strict, devices=ctxs), lambda: B.revcomp(bytes(dna), strict=strict, devices=ctxs)):
                    with pytest.raises(N.TranslationError) as ei:
                        call()
                    try:
                        oracle.revcomp_bytes(bytes(dna), strict=strict)
                        raise AssertionError("oracle accepted it")
                    except oracle.OracleError as o:
                        assert (ei.value.kind, ei.value.byte, ei.value.pos) == (o.kind, o.byte, o.pos), (pos, strict)


def test_calls_restore_the_callers_device(torch, B, N):
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible")
    torch.cuda.set_device(0)
    ctx1 = N.default_context(1)
    x = torch.from_numpy(oracle.synth(3, 0, 48_000)).to("cuda:1")
    B.revcomp_dev(x, ctx=ctx1)
    assert torch.cuda.current_device() == 0
    assert torch.empty(1, device="cuda").device.index == 0
    B.translate_frames_batch([b"ACGTAC"], ctx=ctx1)
    assert torch.cuda.current_device() == 0


def test_import_name_drop_in():
    """`from quickdna import DnaSequence` (quickdna/__init__.py:6 of the reference) binds this implementation."""
    import quickdna
    from quickdna import DnaSequence, ProteinSequence
    from quickdna.quickdna import _check_table, _reverse_complement, _reverse_complement_strict, _translate, _translate_strict

    import quickdna_b200

    assert DnaSequence is quickdna_b200.DnaSequence and quickdna.ProteinSequence is quickdna_b200.ProteinSequence
    assert DnaSequence("AAAGGGAAA").translate(table=1) == ProteinSequence("KGK")
    assert _translate(1, b"AAAGGGAAA") == b"KGK" and _translate_strict(1, b"AAAGGGAAA") == b"KGK"
    assert _reverse_complement(b"AAAAABCCC") == b"GGGVTTTTT"
    with pytest.raises(ValueError):
        _reverse_complement_strict(b"AAAAABCCC")
    with pytest.raises(ValueError, match="not a ncbi translation table: 17"):
        _check_table(17)


def test_reference_wrapper_tests_run_unchanged_when_present():
    """The reference's own tests/test_wrapper.py against the shim package (only where the reference checkout exists)."""
    import os
    import subprocess
    import sys

    ref = "/root/reference/tests/test_wrapper.py"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present (GPU box)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-c", "/dev/null", ref],
                       env=dict(os.environ, PYTHONPATH=root), capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-3000:]
