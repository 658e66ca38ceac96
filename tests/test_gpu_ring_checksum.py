"""Checksums, the chunked ("ring") window enumeration and the BASELINE configs at their stated sizes.  Needs a B200: -m gpu.

Every byte the GPU path produces at the benchmarked sizes is compared with the CPU oracle through the
position-weighted linear checksum of qd_checksum64_dev / oracle.checksum64 (the same definition restated
independently in oracle/qd_oracle.c): whole 10 M x 1 kb six-frame output, whole 250 M-nt reverse complement,
whole chunks of the window enumerations.
"""

import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


def _threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_checksum64_matches_oracle_every_alignment(torch, B):
    rng = np.random.default_rng(1)
    big = rng.integers(0, 256, 1 << 20, dtype=np.uint8)
    d = _dev(torch, big)
    for n in (0, 1, 7, 8, 15, 16, 17, 31, 33, 255, 4096, 100003, (1 << 20) - 40):
        for off in (0, 1, 3, 8, 13, 16, 29):
            for base in (0, 1, 5, 8, 123456789, (1 << 40) + 3):
                got = B.u64(B.checksum64_dev(d[off:off + n], base))
                assert got == oracle.checksum64(big[off:off + n], base), (n, off, base)


def test_checksum64_is_linear_over_pieces(torch, B):
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, 3_000_001, dtype=np.uint8)
    d = _dev(torch, a)
    whole = B.u64(B.checksum64_dev(d, 77))
    acc = None
    cuts = [0, 1, 1000, 1003, 65536, 1_000_007, 3_000_001]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        acc = B.checksum64_dev(d[lo:hi], 77 + lo, acc)  # accumulates
    assert B.u64(acc) == whole == oracle.checksum64(a, 77)
    # a single flipped bit anywhere changes it
    a2 = a.copy()
    a2[2_000_000] ^= 0x10
    assert B.u64(B.checksum64_dev(_dev(torch, a2), 77)) != whole


def test_canonical_windows_chunked_equals_unchunked_and_oracle(torch, B, N):
    n_seq, seq_len, w = 37, 1501, 42
    dna = oracle.synth(6, 0, n_seq * seq_len)
    d = _dev(torch, dna)
    nw = seq_len - w + 1
    whole = B.canonical_windows_batch_dev(d, n_seq, seq_len, w)
    whole_sum = B.u64(B.checksum64_dev(whole.reshape(-1)))
    assert whole_sum == oracle.check_canonical_windows(dna, n_seq, seq_len, w, _threads())
    chunk = 8
    per_chunk = N.lib().qd_canonical_chunk_bytes(chunk, seq_len, w)
    slot = (per_chunk + 255) & ~255
    n_chunks = N.lib().qd_chunk_count(n_seq, chunk)
    # (a) a ring that keeps every chunk: slot contents are the unchunked rows
    ring = torch.zeros(slot * n_chunks, dtype=torch.uint8, device="cuda")
    seen = []
    sums = B.canonical_windows_chunked_dev(d, n_seq, seq_len, w, ring, chunk, consume=lambda *a: seen.append(a[:3] + (a[4],)))
    torch.cuda.synchronize()
    assert [s[:3] for s in seen] == [(k, k * chunk, min(chunk, n_seq - k * chunk)) for k in range(n_chunks)]
    flat = whole.reshape(-1)
    for k in range(n_chunks):
        cnt = min(chunk, n_seq - k * chunk)
        assert seen[k][3] == cnt * nw * w
        assert torch.equal(ring[k * slot:k * slot + cnt * nw * w], flat[k * chunk * nw * w:(k * chunk + cnt) * nw * w]), k
    assert sum(int(x) & M64 for x in sums.cpu().tolist()) & M64 == whole_sum
    # per chunk: the oracle's checksum of that chunk's sequences, rebased to the chunk's place in the stream
    for k in (0, n_chunks - 1):
        cnt = min(chunk, n_seq - k * chunk)
        piece = dna[k * chunk * seq_len:(k * chunk + cnt) * seq_len]
        rows = oracle.canonical_windows(oracle.masks_of(piece[:seq_len]), w)  # sanity of the helper itself on one sequence
        assert oracle.checksum64(oracle.ascii_of(rows.reshape(-1)), 0) == oracle.check_canonical_windows(piece[:seq_len], 1, seq_len, w, 1)
        got = int(sums[k].item()) & M64
        want = oracle.checksum64(bytes(flat[k * chunk * nw * w:(k * chunk + cnt) * nw * w].cpu().numpy()), k * chunk * nw * w)
        assert got == want
    # (b) a ring of TWO slots (slots are reused): the checksums do not change, whatever the chunking
    ring2 = torch.zeros(slot * 2, dtype=torch.uint8, device="cuda")
    sums2 = B.canonical_windows_chunked_dev(d, n_seq, seq_len, w, ring2, chunk)
    assert torch.equal(sums, sums2)
    ring3 = torch.zeros(((N.lib().qd_canonical_chunk_bytes(5, seq_len, w) + 255) & ~255), dtype=torch.uint8, device="cuda")
    sums3 = B.canonical_windows_chunked_dev(d, n_seq, seq_len, w, ring3, 5)
    assert sum(int(x) & M64 for x in sums3.cpu().tolist()) & M64 == whole_sum
    # a ring smaller than one chunk is refused
    with pytest.raises(ValueError):
        B.canonical_windows_chunked_dev(d, n_seq, seq_len, w, ring3[:1024], 5)


def test_windows_all_chunked_matches_oracle(torch, B, N):
    n_seq, seq_len = 11, 2999
    dna = oracle.synth(6, 12345, n_seq * seq_len)
    d = _dev(torch, dna)
    chunk = 4
    L = N.lib()
    import ctypes as C

    aa_off = C.c_size_t()
    slot = (L.qd_windows_all_chunk_bytes(chunk, seq_len, 42, 20, C.byref(aa_off)) + 255) & ~255
    n_chunks = L.qd_chunk_count(n_seq, chunk)
    ring = torch.zeros(slot * n_chunks, dtype=torch.uint8, device="cuda")
    sums = B.windows_all_chunked_dev(d, n_seq, seq_len, ring, chunk)
    for k in range(n_chunks):
        cnt = min(chunk, n_seq - k * chunk)
        piece = dna[k * chunk * seq_len:(k * chunk + cnt) * seq_len]
        assert int(sums[k].item()) & M64 == oracle.check_windows_all(piece, cnt, seq_len, 42, 20, _threads()), k
        # and the slot holds what the unchunked call writes for those sequences
        cs, dna_out, aa_out = B.windows_all_batch_dev(d[k * chunk * seq_len:(k * chunk + cnt) * seq_len], cnt, seq_len)
        off = C.c_size_t()
        L.qd_windows_all_chunk_bytes(cnt, seq_len, 42, 20, C.byref(off))
        nd, na = 2 * cnt * cs[0] * 42, cnt * sum(cs[2:]) * 20
        assert torch.equal(ring[k * slot:k * slot + nd], dna_out[:nd])
        assert torch.equal(ring[k * slot + off.value:k * slot + off.value + na], aa_out[:na])
    # two slots, reused
    ring2 = torch.zeros(slot * 2, dtype=torch.uint8, device="cuda")
    assert torch.equal(B.windows_all_chunked_dev(d, n_seq, seq_len, ring2, chunk), sums)


def _ambiguous_windows(rng, n_win, w=42, n_amb=2):
    win = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n_win, w))].copy()
    codes = np.frombuffer(b"WMYHRKDSVBN", dtype=np.uint8)
    for _ in range(n_amb):
        win[np.arange(n_win), rng.integers(0, w, n_win)] = codes[rng.integers(0, 11, n_win)]
    return win


def test_expansions_chunked_matches_oracle(torch, B, N):
    rng = np.random.default_rng(5)
    n_win, w, chunk = 10_007, 42, 3000
    win = _ambiguous_windows(rng, n_win)
    win[17, :6] = np.frombuffer(b"NNNNNN", dtype=np.uint8)  # 4096 expansions: skipped (cap 16)
    d = _dev(torch, win)
    L = N.lib()
    slot = (L.qd_expansion_chunk_bytes(chunk, w, 16) + 255) & ~255
    n_chunks = L.qd_chunk_count(n_win, chunk)
    ring = torch.zeros(slot * n_chunks, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(n_win, dtype=torch.int64, device="cuda")
    rows, sums = B.expansion_windows_chunked_dev(d, ring, chunk, 16, counts=counts)
    c_all, oi_all, out_all = B.expansion_windows_dev(d, 16)
    assert torch.equal(counts, c_all)
    for k in range(n_chunks):
        lo, hi = k * chunk, min(n_win, (k + 1) * chunk)
        want_sum, want_rows = oracle.check_expansions(win[lo:hi], 16, _threads())
        assert int(rows[k].item()) == want_rows == int(oi_all[hi].item()) - int(oi_all[lo].item())
        assert int(sums[k].item()) & M64 == want_sum, k
        r0 = int(oi_all[lo].item())
        assert torch.equal(ring[k * slot:k * slot + want_rows * w], out_all[r0 * w:(r0 + want_rows) * w])
    ring1 = torch.zeros(slot, dtype=torch.uint8, device="cuda")
    rows1, sums1 = B.expansion_windows_chunked_dev(d, ring1, chunk, 16)
    assert torch.equal(rows1, rows) and torch.equal(sums1, sums)


@pytest.mark.parametrize("n_win,w,shift", [(300_011, 42, 0), (70_001, 42, 3), (50_000, 37, 0), (1, 42, 0), (33, 64, 1)])
def test_expansions_one_pass_many_tiles(torch, B, n_win, w, shift):
    """The one-pass kernel chains the first row of every warp tile through status words (decoupled look-back): more tiles
    than resident warps, windows that are skipped, odd input alignment, both encodings -- row index = running sum of
    the emitted counts, rows = the oracle's (whole-output checksum)."""
    rng = np.random.default_rng(n_win + w)
    win = _ambiguous_windows(rng, n_win, w=w, n_amb=3)
    win[rng.integers(0, n_win, max(1, n_win // 50)), :4] = np.frombuffer(b"NNNN", dtype=np.uint8)  # >= 256 expansions: skipped at cap 100
    cap = 100
    for enc, arr in ((0, win), (1, oracle.masks_of(win.reshape(-1)).reshape(n_win, w))):
        flat = torch.zeros(n_win * w + shift, dtype=torch.uint8, device="cuda")
        flat[shift:] = _dev(torch, arr).view(-1)
        d = flat[shift:].view(n_win, w)
        B.dev_error_reset()
        c_d, oi_d, out_d = B.expansion_windows_dev(d, cap, enc=enc)
        assert B.dev_error_fetch(d.reshape(-1)) is None
        counts = c_d.cpu().numpy().astype(np.uint64)
        emit = np.where(counts <= cap, counts, 0).astype(np.uint64)
        want_index = np.concatenate([[0], np.cumsum(emit)]).astype(np.uint64)
        assert (oi_d.cpu().numpy().astype(np.uint64) == want_index).all()
        total = int(want_index[-1])
        want_sum, want_rows = oracle.check_expansions(win, cap, _threads())
        assert total == want_rows
        rows = out_d[: total * w].cpu().numpy()
        if enc == 1:
            rows = oracle.ascii_of(rows)
        assert oracle.checksum64(rows, 0) == want_sum


def test_two_streams_on_one_context_are_ordered(torch, B):
    """scratch and the error key are per context; calls on different streams must not race on them."""
    rng = np.random.default_rng(9)

    def make(n_seq, lo, hi):
        lens = rng.integers(lo, hi, n_seq)
        offs = np.zeros(n_seq + 1, dtype=np.int64)
        offs[1:] = np.cumsum(lens)
        flat = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(offs[-1]))]
        return flat, offs

    batches = [make(20000, 1, 400), make(5000, 500, 1500), make(30000, 1, 100)]
    streams = [torch.cuda.Stream() for _ in batches]
    dev = [(_dev(torch, f), _dev(torch, o)) for f, o in batches]
    torch.cuda.synchronize()
    outs = []
    B.dev_error_reset()
    for rep in range(3):
        outs = []
        for (f, o), st in zip(dev, streams):
            with torch.cuda.stream(st):
                outs.append(B.translate_frames_csr_dev(f, o, table=1, frames=6))
    torch.cuda.synchronize()
    for (flat, offs), out in zip(batches, outs):
        want = B.translate_frames_batch(flat, offsets=offs.astype(np.uint64), frames=6)
        assert bytes(out.cpu().numpy()) == bytes(want.data)
        for i in (0, len(offs) - 2):
            assert want.sequence_frames(i) == oracle.py_all_frames(1, flat[offs[i]:offs[i + 1]].tobytes())


def test_translate_batch_tuple_of_reads_is_a_sequence_of_reads(B):
    reads = (b"ACGTTTAAAG", b"TTTAAACCCGGG")
    res = B.translate_frames_batch(reads, frames=6)
    for i, r in enumerate(reads):
        assert res.sequence_frames(i) == oracle.py_all_frames(1, r)
    with pytest.raises(ValueError):
        B.translate_frames_batch(np.frombuffer(b"ACGTAC", dtype=np.uint8), offsets=np.array([0, 4, 3, 6], dtype=np.uint64))


def test_bad_table_message_on_device_calls(torch, B):
    d = _dev(torch, np.frombuffer(b"ACGTACGTACGT", dtype=np.uint8))
    with pytest.raises(ValueError, match="not a ncbi translation table: 17"):
        B.translate_frames_uniform_dev(d, 1, 12, table=17)


# ------------------------------------------------------------------ BASELINE configs at their stated sizes


def test_config3_whole_output_checksum_10m_reads(torch, B):
    """translate_all_frames on 10 M x 1 kb reads: the checksum of ALL 20.16 GB of output equals the oracle's."""
    from quickdna_b200.synth import synth_torch

    n_reads, L = 10_000_000, 1000
    first = 3_000_000  # not the start of the stream
    x = synth_torch(3, first * L, n_reads * L, torch.device("cuda"))
    B.dev_error_reset()
    out = B.translate_frames_uniform_dev(x, n_reads, L, table=1, frames=6)
    assert B.dev_error_fetch(x, uniform_len=L) is None
    got = B.u64(B.checksum64_dev(out))
    want = oracle.check_frames_uniform(3, first, n_reads, L, 0, 6, _threads())
    assert got == want
    # the CSR form of the same batch produces the same bytes
    offs = torch.arange(0, (n_reads + 1) * L, L, dtype=torch.int64, device="cuda")
    out2 = torch.empty_like(out)
    B.translate_frames_csr_dev(x, offs, out=out2, table=1, frames=6)
    assert B.u64(B.checksum64_dev(out2)) == want
    del out, out2, x


def test_config2_whole_output_checksum_250m_revcomp(torch, B):
    from quickdna_b200.synth import synth_torch

    for n in (250_000_000, 250_000_001, 249_999_999):
        x = synth_torch(2, 0, n, torch.device("cuda"))
        B.dev_error_reset()
        out = B.revcomp_dev(x)
        assert B.dev_error_fetch(x) is None
        assert B.u64(B.checksum64_dev(out)) == oracle.check_revcomp(2, 0, n, False, _threads()), n
        del x, out


def test_config5_full_chunks_at_bench_shape(torch, B, N):
    """One full ring chunk of each window enumeration at the all_windows.rs sequence length (99 999 nt), every byte
    checked through the checksum; and chunking invariance of the canonical rows over a larger batch."""
    from quickdna_b200.synth import synth_torch

    seq_len, w = 99_999, 42
    n_seq, chunk = 600, 128
    x = synth_torch(6, 0, n_seq * seq_len, torch.device("cuda"))
    L = N.lib()
    ring = torch.empty(2 * ((L.qd_canonical_chunk_bytes(chunk, seq_len, w) + 255) & ~255), dtype=torch.uint8, device="cuda")
    B.dev_error_reset()
    sums = B.canonical_windows_chunked_dev(x, n_seq, seq_len, w, ring, chunk)
    assert B.dev_error_fetch(x, uniform_len=seq_len, strict=True) is None
    host = x[:chunk * seq_len].cpu().numpy()
    want0 = oracle.check_canonical_windows(host, chunk, seq_len, w, _threads())
    assert int(sums[0].item()) & M64 == want0
    sums_b = B.canonical_windows_chunked_dev(x, n_seq, seq_len, w, ring, 77)
    assert sum(int(v) & M64 for v in sums.cpu().tolist()) & M64 == sum(int(v) & M64 for v in sums_b.cpu().tolist()) & M64
    # all_windows: 64 sequences per chunk (0.8 GB per slot)
    import ctypes as C

    chunk_a = 64
    slot = (L.qd_windows_all_chunk_bytes(chunk_a, seq_len, 42, 20, C.byref(C.c_size_t())) + 255) & ~255
    ring_a = torch.empty(2 * slot, dtype=torch.uint8, device="cuda")
    sums_a = B.windows_all_chunked_dev(x, 3 * chunk_a, seq_len, ring_a, chunk_a)
    for k in (0, 2):
        piece = x[k * chunk_a * seq_len:(k + 1) * chunk_a * seq_len].cpu().numpy()
        assert int(sums_a[k].item()) & M64 == oracle.check_windows_all(piece, chunk_a, seq_len, 42, 20, _threads()), k
