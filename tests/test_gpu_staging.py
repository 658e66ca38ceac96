"""Pageable caller buffers: large host calls are staged through pinned buffers of the context by a few host threads
(qd_ctx_set_stage_threads).  Every staged result must equal the result of the plain (driver-staged) copies and the
oracle's, byte for byte.  Needs a B200: -m gpu."""

import ctypes as C
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


@pytest.fixture(scope="module")
def N():
    from quickdna_b200 import _native

    return _native


@pytest.fixture(scope="module")
def B():
    from quickdna_b200 import batch

    return batch


@pytest.fixture()
def ctx(N):
    c = N.Context(0)
    yield c
    c.close()


def _staged_and_plain(N, ctx, call):
    """call() -> bytes-like result; run with staging on (3 threads), then off."""
    L = N.lib()
    assert L.qd_ctx_set_stage_threads(ctx.handle, 3) == 0
    a = call()
    assert L.qd_ctx_set_stage_threads(ctx.handle, 0) == 0
    b = call()
    assert L.qd_ctx_set_stage_threads(ctx.handle, -1) == 0
    return a, b


def test_uniform_batch_staged_equals_plain_and_oracle(N, B, ctx):
    n_reads, L_ = 20_011, 1000  # 20 MB in + 40 MB out: above the staging threshold, several 16-MiB chunks
    reads = oracle.synth(3, 0, n_reads * L_).reshape(n_reads, L_)  # pageable numpy memory
    a, b = _staged_and_plain(N, ctx, lambda: B.translate_frames_batch(reads, table=1, frames=6, ctx=ctx).data.copy())
    assert np.array_equal(a, b)
    assert oracle.checksum64(a, 0) == oracle.check_frames_uniform(3, 0, n_reads, L_, 0, 6, _threads())


def test_csr_batch_staged_equals_plain(N, B, ctx):
    rng = np.random.default_rng(4)
    lens = rng.integers(1, 3000, 16_000)
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    flat = oracle.synth(5, 0, int(offs[-1]))
    a, b = _staged_and_plain(N, ctx, lambda: B.translate_frames_batch((flat, offs), table=11, frames=6, ctx=ctx).data.copy())
    assert np.array_equal(a, b)
    res = B.translate_frames_batch((flat, offs), table=11, frames=6, ctx=ctx)
    for i in (0, 1, 7777, len(lens) - 1):
        assert res.sequence_frames(i) == oracle.py_all_frames(11, flat[int(offs[i]):int(offs[i + 1])].tobytes())


@pytest.mark.parametrize("n", [40_000_000, 33_554_433])
def test_long_revcomp_and_frames_staged(N, B, ctx, n):
    dna = oracle.synth(3, 0, n)
    a, b = _staged_and_plain(N, ctx, lambda: B.revcomp(dna, ctx=ctx).copy())
    assert np.array_equal(a, b)
    assert oracle.checksum64(a, 0) == oracle.check_revcomp(3, 0, n, False, _threads())
    fa, fb = _staged_and_plain(N, ctx, lambda: B.translate_frames(dna, table=1, frames=6, ctx=ctx))
    assert len(fa) == len(fb) == 6
    for x, y in zip(fa, fb):
        assert x == y
    want = oracle.py_all_frames(1, dna[:30_000].tobytes())
    assert bytes(fa[0][:9000]) == want[0][:9000]


def test_error_position_through_the_staged_path(N, B, ctx):
    n_reads, L_ = 40_000, 1000
    reads = oracle.synth(3, 0, n_reads * L_).reshape(n_reads, L_).copy()
    reads[31_337, 501] = ord("!")
    assert N.lib().qd_ctx_set_stage_threads(ctx.handle, 2) == 0
    with pytest.raises(ValueError) as ei:
        B.translate_frames_batch(reads, table=1, frames=6, ctx=ctx)
    assert (ei.value.seq_index, ei.value.pos, ei.value.kind) == (31_337, 501, 2)


def test_fasta_text_staged(N, B, ctx):
    rec = 20_000
    lines = []
    body = oracle.synth(7, 0, rec * 1960).reshape(rec, 28, 70)
    for i in range(rec):
        lines.append(b">r%d\n" % i)
        lines.append(b"\n".join(body[i, k].tobytes() for k in range(28)) + b"\n")
    text = b"".join(lines)  # ~ 40 MB
    assert len(text) > (32 << 20)
    L = N.lib()
    assert L.qd_ctx_set_stage_threads(ctx.handle, 3) == 0
    ra = B.fasta_scan(text, ctx=ctx)
    assert L.qd_ctx_set_stage_threads(ctx.handle, 0) == 0
    rb = B.fasta_scan(text, ctx=ctx)
    assert len(ra) == len(rb) == rec
    assert np.array_equal(ra.masks, rb.masks)
    assert ra.masks.size == rec * 1960
    assert oracle.ascii_of(ra.masks[:1960]) == body[0].tobytes()


def test_window_and_parse_calls_staged_equal_plain(N, B, ctx):
    """The other host entry points move their large buffers through the same staging: windows (42 bytes out per byte in),
    canonical windows and keys, the parse, the expansions."""
    n = 2_000_003
    dna = oracle.synth(3, 0, n)
    a, b = _staged_and_plain(N, ctx, lambda: B.windows(dna, 42, ctx=ctx).copy())  # 84 MB out
    assert np.array_equal(a, b) and a.shape == (n - 41, 42)
    assert a[12345].tobytes() == dna[12345:12345 + 42].tobytes()
    a, b = _staged_and_plain(N, ctx, lambda: B.canonical_windows(dna, 42, ctx=ctx).copy())
    assert np.array_equal(a, b)
    want = oracle.ascii_of(oracle.canonical_windows(oracle.masks_of(dna[:5000]), 42).reshape(-1))
    assert a[: 5000 - 41].tobytes() == want
    a, b = _staged_and_plain(N, ctx, lambda: B.canonical_window_keys(dna, 42, kind=N.CANON_KEY_PACKED, ctx=ctx).copy())  # 32 MB of keys
    assert np.array_equal(a, b)
    big = oracle.synth(2, 0, 40_000_000)
    big[1000::61] = 0x20
    a, b = _staged_and_plain(N, ctx, lambda: B.parse(big, ctx=ctx).copy())
    assert np.array_equal(a, b)
    assert np.array_equal(a[:5000], oracle.parse(big[:6000].tobytes())[:5000])
    rng = np.random.default_rng(3)
    win = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (400_000, 42))].copy()
    win[np.arange(400_000), rng.integers(0, 42, 400_000)] = ord("N")
    win[np.arange(400_000), rng.integers(0, 42, 400_000)] = ord("R")
    (ca, ia, ra), (cb, ib, rb) = _staged_and_plain(N, ctx, lambda: B.expansion_windows(win, max_expansions=16, ctx=ctx))  # ~130 MB of rows
    assert np.array_equal(ca, cb) and np.array_equal(ia, ib) and np.array_equal(ra, rb)
    want_sum, want_rows = oracle.check_expansions(win, 16, _threads())
    assert len(ra) == want_rows and oracle.checksum64(np.ascontiguousarray(ra).reshape(-1), 0) == want_sum


def test_pinned_arrays_take_the_direct_path(N, B, ctx):
    """batch.pinned_empty: numpy arrays in page-locked memory; results equal the staged ones."""
    n_reads, L_ = 20_011, 1000
    reads = oracle.synth(3, 0, n_reads * L_).reshape(n_reads, L_)
    pinned = B.pinned_empty((n_reads, L_))
    pinned[:] = reads
    attr_type = None
    try:
        import torch

        attr_type = torch.cuda.cudart().cudaPointerGetAttributes(pinned.ctypes.data)[1].type  # 1 = host (pinned)
    except Exception:
        pass
    a = B.translate_frames_batch(pinned, table=1, frames=6, ctx=ctx).data
    b = B.translate_frames_batch(reads, table=1, frames=6, ctx=ctx).data
    assert np.array_equal(a, b)
    assert attr_type in (None, 1) or int(attr_type) == 1
    del pinned  # releases the block


@pytest.mark.parametrize("n", [131_071, 131_072, 131_073, 700_001, 2_500_000])
def test_single_sequence_size_classes(N, B, ctx, n):
    """One sequence through every host path: tiny (mapped memory, <= 128 KiB), small (result + key in one copy, <= 2 MiB of
    result), plain (separate copies), each against the oracle -- frames, reverse complement, and the first error."""
    dna = oracle.synth(3, n, n)
    got = B.translate_frames(dna, table=11, frames=6, ctx=ctx)
    want = oracle.py_all_frames(11, dna.tobytes())
    assert [bytes(f) for f in got] == want
    assert B.revcomp(dna, ctx=ctx).tobytes() == oracle.revcomp_bytes(dna.tobytes())
    one = B.translate_frames(dna, table=2, frames=N.FRAMES_ONE, ctx=ctx)
    assert bytes(one[0]) == oracle.translate_bytes(2, dna.tobytes())
    bad = dna.copy()
    bad[n - 7] = ord("x")
    bad[n // 2] = ord("B")
    with pytest.raises(ValueError) as ei:
        B.translate_frames(bad, table=1, frames=6, strict=True, ctx=ctx)
    assert (ei.value.pos, ei.value.kind, chr(ei.value.byte)) == (n // 2, 3, "B")
    with pytest.raises(ValueError) as ei:
        B.revcomp(bad, ctx=ctx)
    assert (ei.value.pos, ei.value.kind) == (n - 7, 2)


def test_tiny_and_small_call_paths_agree(N, B, monkeypatch):
    """Inputs of at most 128 KiB run on mapped host memory (one launch); QD_TINY_CALLS=0 (read when a context is made)
    sends them through the copying path instead.  Same bytes, same errors."""
    monkeypatch.setenv("QD_TINY_CALLS", "0")
    plain = N.Context(0)
    monkeypatch.delenv("QD_TINY_CALLS")
    tiny = N.Context(0)
    try:
        rng = np.random.default_rng(8)
        for n in (1, 2, 3, 47, 48, 49, 1000, 29_901, 131_072):
            dna = np.frombuffer(b"ACGTacgtNRYK", dtype=np.uint8)[rng.integers(0, 12, n)].copy()
            for frames in (N.FRAMES_ONE, N.FRAMES_SELF, N.FRAMES_ALL):
                assert B.translate_frames(dna, table=4, frames=frames, ctx=tiny) == B.translate_frames(dna, table=4, frames=frames, ctx=plain)
            assert np.array_equal(B.revcomp(dna, ctx=tiny), B.revcomp(dna, ctx=plain))
            spaced = dna.copy()
            spaced[::17] = 0x20
            assert np.array_equal(B.parse(spaced, ctx=tiny), B.parse(spaced, ctx=plain))
            assert np.array_equal(B.parse(dna, ctx=tiny), oracle.parse(dna.tobytes()))
        reads = oracle.synth(3, 0, 100 * 1000).reshape(100, 1000)
        assert np.array_equal(B.translate_frames_batch(reads, ctx=tiny).data, B.translate_frames_batch(reads, ctx=plain).data)
        bad = reads.copy()
        bad[57, 123] = ord("@")
        errs = []
        for c in (tiny, plain):
            with pytest.raises(ValueError) as ei:
                B.translate_frames_batch(bad, ctx=c)
            errs.append((ei.value.seq_index, ei.value.pos, ei.value.kind, ei.value.byte))
        assert errs[0] == errs[1] == (57, 123, 2, ord("@"))
        # variable-length batches: nucleotides and offsets in mapped memory, the scan and map kernels run on them as they are
        for lens in ([0, 1, 2, 3, 47, 48, 49, 300, 0, 1000, 5], rng.integers(0, 700, 150).tolist(), [100_000], [0, 0, 0], [2, 2]):
            offs = np.zeros(len(lens) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum(lens)
            flat = np.frombuffer(b"ACGTacgtNRYK", dtype=np.uint8)[rng.integers(0, 12, max(int(offs[-1]), 1))].copy()[: int(offs[-1])]
            for frames in (N.FRAMES_ONE, N.FRAMES_SELF, N.FRAMES_ALL):
                a = B.translate_frames_batch((flat, offs), table=2, frames=frames, ctx=tiny)
                b = B.translate_frames_batch((flat, offs), table=2, frames=frames, ctx=plain)
                assert np.array_equal(a.data, b.data) and np.array_equal(a.out_offsets, b.out_offsets)
                for i in range(0, len(lens), 7):
                    assert a.sequence_frames(i) == b.sequence_frames(i)
            if len(lens) > 9 and lens[9] == 1000:
                r = flat[int(offs[9]):int(offs[10])].tobytes()
                assert B.translate_frames_batch((flat, offs), table=2, frames=6, ctx=tiny).sequence_frames(9) == oracle.py_all_frames(2, r)
                dirty = flat.copy()
                dirty[int(offs[9]) + 77] = ord("#")
                errs = []
                for c in (tiny, plain):
                    with pytest.raises(ValueError) as ei:
                        B.translate_frames_batch((dirty, offs), table=2, frames=6, ctx=c)
                    errs.append((ei.value.seq_index, ei.value.pos, ei.value.kind, ei.value.byte))
                assert errs[0] == errs[1] == (9, 77, 2, ord("#"))
    finally:
        tiny.close()
        plain.close()


def test_dependent_launches_and_plain_launches_agree(N, B, monkeypatch):
    """Chains of short kernels (parse, variable-length batches, the FASTA scanner, the whole-sequence canonical form, the nine
    launches of an all_windows call) run as programmatic dependent launches; QD_PDL=0 (read when a context is made) launches
    them plainly.  Same bytes either way, and equal to the oracle's."""
    import torch

    monkeypatch.setenv("QD_PDL", "0")
    plain = N.Context(0)
    monkeypatch.delenv("QD_PDL")
    pdl = N.Context(0)
    try:
        rng = np.random.default_rng(21)
        dna = oracle.synth(3, 0, 3_000_000)
        spaced = dna.copy()
        spaced[500::97] = 0x20
        for src in (dna, spaced):
            a, b = B.parse(src, ctx=pdl), B.parse(src, ctx=plain)
            assert np.array_equal(a, b)
            assert np.array_equal(a[:4000], oracle.parse(src[:5000].tobytes())[:4000])
        lens = np.concatenate([rng.integers(0, 2500, 3000), [48 * 600, 48 * 65536 + 7, 0, 5]])
        offs = np.zeros(len(lens) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(lens)
        flat = oracle.synth(5, 0, int(offs[-1]))
        ra = B.translate_frames_batch((flat, offs), table=11, frames=6, ctx=pdl)
        rb = B.translate_frames_batch((flat, offs), table=11, frames=6, ctx=plain)
        assert np.array_equal(ra.data, rb.data)
        for i in (0, 17, 2999, 3000, 3001, 3003):
            assert ra.sequence_frames(i) == (oracle.py_all_frames(11, flat[int(offs[i]):int(offs[i + 1])].tobytes()) if lens[i] >= 3 else [])
        text = b"".join(b">r%d\n" % i + oracle.synth(7, i * 700, 700).tobytes() + b"\n" for i in range(300))
        sa, fa = B.fasta_translate_frames(text, ctx=pdl)
        sb, fb = B.fasta_translate_frames(text, ctx=plain)
        assert len(sa) == len(sb) == 300 and np.array_equal(fa.data, fb.data)
        assert fa.sequence_frames(7) == oracle.py_all_frames(1, oracle.synth(7, 7 * 700, 700).tobytes())
        assert np.array_equal(B.canonical(dna[:200_001], ctx=pdl), B.canonical(dna[:200_001], ctx=plain))
        x = torch.from_numpy(dna[: 40 * 5000].reshape(40, 5000).copy()).cuda()
        ca, da, aa = B.windows_all_batch_dev(x, 40, 5000, ctx=pdl)
        cb, db, ab = B.windows_all_batch_dev(x, 40, 5000, ctx=plain)
        torch.cuda.synchronize()
        assert ca == cb and torch.equal(da, db) and torch.equal(aa, ab)
        wd, wa, _o, _c = oracle.windows_all(oracle.masks_of(dna[:5000]), 42, 20)
        assert da[: ca[0] * 42].cpu().numpy().tobytes() == wd[: ca[0]].tobytes()
    finally:
        pdl.close()
        plain.close()
