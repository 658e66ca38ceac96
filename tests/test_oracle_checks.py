"""The oracle's whole-config check functions (oracle/qd_oracle.c) against its own pinned per-sequence functions, on
sizes that finish in seconds: they are what the -m gpu tests and bench.py compare full-size GPU outputs with."""

import numpy as np

import oracle
from quickdna_b200.synth import mix64_np, synth_np

M64 = (1 << 64) - 1


def _checksum_np(buf: np.ndarray, base: int) -> int:
    """qd_checksum64 restated with numpy: sum over stream words of word * (mix64(J * GOLD + 1) | 1)."""
    pad_front = base & 7
    a = np.concatenate([np.zeros(pad_front, np.uint8), np.asarray(buf, np.uint8)])
    a = np.concatenate([a, np.zeros((-a.size) % 8, np.uint8)])
    words = a.view("<u8")
    j = np.arange(words.size, dtype=np.uint64) + np.uint64(base >> 3)
    with np.errstate(over="ignore"):
        # mix64_np adds GOLD first: mix64(J * GOLD + 1) = finaliser(J * GOLD + 1) -> undo the add
        k = mix64_np(j * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1) - np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
        return int((words * k).sum(dtype=np.uint64)) & M64


def test_synth_matches_numpy_generator():
    for seed, start, n, alpha in ((3, 0, 10000, b"ACGT"), (2, 10**12 + 7, 4097, b"ACGT"), (5, 99, 1000, b"WMRYSKBVDHN")):
        assert (oracle.synth(seed, start, n, alpha) == synth_np(seed, start, n, alpha)).all()


def test_checksum64_definition_and_linearity():
    rng = np.random.default_rng(0)
    buf = rng.integers(0, 256, 10_001, dtype=np.uint8)
    for base in (0, 1, 7, 8, 123457, (1 << 40) + 5):
        assert oracle.checksum64(buf, base) == _checksum_np(buf, base)
    whole = oracle.checksum64(buf, 13)
    cuts = [0, 1, 9, 4096, 4099, 10_001]
    parts = sum(oracle.checksum64(buf[a:b], 13 + a) for a, b in zip(cuts[:-1], cuts[1:])) & M64
    assert parts == whole
    assert oracle.checksum64(b"", 5) == 0


def _slot_image(read: bytes, frames: int) -> bytes:
    n = len(read)
    R = (n + 47) // 48
    img = bytearray(16 * R * frames)
    if frames == 1:
        fs = [oracle.translate_bytes(1, read)]
    elif frames == 3:
        fs = [oracle.translate_bytes(1, read[k:]) if n > k else b"" for k in range(3)]
    else:
        rc = oracle.revcomp_bytes(read)
        fs = [oracle.translate_bytes(1, read[k:]) if n > k else b"" for k in range(3)] + \
             [oracle.translate_bytes(1, rc[k:]) if n > k else b"" for k in range(3)]
    for f, aa in enumerate(fs):
        if f < 3:
            img[16 * R * f:16 * R * f + len(aa)] = aa
        else:
            img[16 * R * (f + 1) - len(aa):16 * R * (f + 1)] = aa
    return bytes(img)


def test_check_frames_uniform_is_the_checksum_of_the_slot_layout():
    for L, frames in ((1000, 6), (100, 6), (47, 3), (99, 1), (5, 6)):
        first, n = 1234, 37
        dna = synth_np(3, first * L, n * L)
        img = b"".join(_slot_image(dna[i * L:(i + 1) * L].tobytes(), frames) for i in range(n))
        for threads in (1, 3):
            assert oracle.check_frames_uniform(3, first, n, L, 0, frames, threads) == oracle.checksum64(img, 0), (L, frames)


def test_check_revcomp_is_the_checksum_of_the_reverse_complement():
    for n in (1, 63, 64, 65, 100_003, 2_200_001):
        dna = synth_np(2, 77, n)
        want = oracle.checksum64(oracle.revcomp_bytes(dna.tobytes()), 0)
        for threads in (1, 4):
            assert oracle.check_revcomp(2, 77, n, False, threads) == want, n


def test_check_windows_and_expansions():
    n_seq, L, w = 5, 333, 42
    dna = synth_np(6, 0, n_seq * L)
    rows = b"".join(oracle.ascii_of(oracle.canonical_windows(oracle.masks_of(dna[i * L:(i + 1) * L]), w).reshape(-1)) for i in range(n_seq))
    assert oracle.check_canonical_windows(dna, n_seq, L, w, 2) == oracle.checksum64(rows, 0)
    # all_windows: dna stream [fwd blocks of all sequences][rc blocks], then the six aa frame blocks
    per = [oracle.windows_all(oracle.masks_of(dna[i * L:(i + 1) * L]), 42, 20) for i in range(n_seq)]
    c = per[0][3]
    fwd = b"".join(p[0][:c[0]].tobytes() for p in per)
    rc = b"".join(p[0][c[0]:].tobytes() for p in per)
    aa = b""
    at = 0
    for f in range(6):
        aa += b"".join(p[1][at:at + c[2 + f]].tobytes() for p in per)
        at += c[2 + f]
    assert oracle.check_windows_all(dna, n_seq, L, 42, 20, 2) == oracle.checksum64(fwd + rc + aa, 0)
    # expansions
    rng = np.random.default_rng(3)
    win = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (200, 20))].copy()
    win[np.arange(200), rng.integers(0, 20, 200)] = np.frombuffer(b"WMYHRKDSVBN", np.uint8)[rng.integers(0, 11, 200)]
    win[np.arange(200), rng.integers(0, 20, 200)] = np.frombuffer(b"WMYHRKDSVBN", np.uint8)[rng.integers(0, 11, 200)]
    win[7, :4] = np.frombuffer(b"NNNN", np.uint8)  # 256 expansions or more: skipped at cap 16
    rows = b""
    n_rows = 0
    for i in range(200):
        m = oracle.masks_of(win[i])
        hint = oracle.expansions_size_hint(m)
        if hint is not None and hint <= 16:
            e = oracle.expansions(m)
            rows += oracle.ascii_of(e.reshape(-1))
            n_rows += e.shape[0]
    for threads in (1, 3):
        assert oracle.check_expansions(win, 16, threads) == (oracle.checksum64(rows, 0), n_rows)
