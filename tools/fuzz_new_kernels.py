"""Randomised checks of the kernels reworked in round 2 against numpy / the oracle: window tiles cut on output lines (any w,
length, batch size, output alignment), the one-pass expansion kernel (any w <= 64, cap < 128, ambiguity density, input
alignment), the gated parse (any length, whitespace density, alignment).  python tools/fuzz_new_kernels.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import oracle  # noqa: E402
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dev = torch.device("cuda:0")
threads = len(os.sched_getaffinity(0))
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
AMB = np.frombuffer(b"WMYHRKDSVBNwrn", dtype=np.uint8)

for it in range(iters):
    # ---- windows: every w-byte window of n_seq sequences, output base at any byte offset
    w = int(rng.integers(1, 65))
    L = int(rng.integers(w, w + rng.choice([5, 200, 3000, 20000])))
    n_seq = int(rng.integers(1, 40))
    off = int(rng.choice([0, 0, 16, 4, 2, 1, 7, 64, 100]))
    seqs = rng.integers(65, 91, (n_seq, L), dtype=np.uint8)
    nw = L - w + 1
    buf = torch.zeros(n_seq * nw * w + off + 64, dtype=torch.uint8, device=dev)
    got = B.windows_batch_dev(torch.from_numpy(seqs).to(dev), n_seq, L, w, mode=0, out=buf[off:])
    want = np.lib.stride_tricks.sliding_window_view(seqs, w, axis=1).reshape(n_seq * nw, w)
    assert np.array_equal(got.cpu().numpy(), want), ("windows", w, L, n_seq, off)
    assert int(buf[:off].sum()) == 0 and int(buf[off + n_seq * nw * w:].sum()) == 0, ("windows wrote outside", w, L, n_seq, off)

    # ---- expansions: one pass, any w / cap / density / alignment
    w = int(rng.choice([42, 42, int(rng.integers(1, 65))]))
    n_win = int(rng.integers(1, rng.choice([40, 3000, 40000])))
    cap = int(rng.choice([1, 4, 16, 16, 100, 127]))
    dens = float(rng.choice([0.0, 0.02, 0.05, 0.1]))
    win = ACGT[rng.integers(0, 4, (n_win, w))].copy()
    m = rng.random((n_win, w)) < dens
    win[m] = AMB[rng.integers(0, len(AMB), int(m.sum()))]
    shift = int(rng.choice([0, 0, 1, 2, 3, 8, 15]))
    flat = torch.zeros(n_win * w + shift, dtype=torch.uint8, device=dev)
    flat[shift:] = torch.from_numpy(win.reshape(-1)).to(dev)
    d = flat[shift:].view(n_win, w)
    B.dev_error_reset()
    c_d, oi_d, out_d = B.expansion_windows_dev(d, cap)
    assert B.dev_error_fetch(d.reshape(-1)) is None
    counts = c_d.cpu().numpy().astype(np.uint64)
    hints = np.array([min(oracle.expansions_size_hint(oracle.masks_of(win[i])) or 2 ** 64 - 1, 2 ** 64 - 1) for i in range(min(n_win, 300))], dtype=np.uint64)
    assert np.array_equal(counts[: len(hints)], hints), ("expansion counts", w, n_win, cap, dens, shift)
    emit = np.where(counts <= cap, counts, 0).astype(np.uint64)
    want_index = np.concatenate([[0], np.cumsum(emit)]).astype(np.uint64)
    assert np.array_equal(oi_d.cpu().numpy().astype(np.uint64), want_index), ("expansion index", w, n_win, cap, dens, shift)
    total = int(want_index[-1])
    want_sum, want_rows = oracle.check_expansions(win, cap, threads)
    assert total == want_rows, ("expansion rows", total, want_rows)
    assert oracle.checksum64(out_d[: total * w].cpu().numpy(), 0) == want_sum, ("expansion bytes", w, n_win, cap, dens, shift)

    # ---- parse: any length, whitespace density, input / output alignment
    n = int(rng.integers(1, rng.choice([100, 5000, 300000])))
    dens = float(rng.choice([0.0, 0.0, 1e-4, 0.02, 0.3]))
    txt = np.frombuffer(b"ACGTacgtNRYKMswbdhv", dtype=np.uint8)[rng.integers(0, 19, n)].copy()
    ws = rng.random(n) < dens
    txt[ws] = np.frombuffer(b" \t", dtype=np.uint8)[rng.integers(0, 2, int(ws.sum()))]
    a_in, a_out = int(rng.integers(0, 32)), int(rng.integers(0, 32))
    if rng.random() < 0.5:
        a_out = a_in  # the one-pass case
    src = torch.zeros(n + a_in + 16, dtype=torch.uint8, device=dev)
    src[a_in:a_in + n] = torch.from_numpy(txt).to(dev)
    dst = torch.full((n + a_out + 48,), 0xEE, dtype=torch.uint8, device=dev)
    masks, cnt = B.parse_dev(src[a_in:a_in + n], out=dst[a_out:a_out + n + 16])
    want = oracle.parse(txt.tobytes())
    k = int(cnt.item())
    assert k == len(want) and np.array_equal(dst[a_out:a_out + k].cpu().numpy(), want), ("parse", n, dens, a_in, a_out)
    assert int((dst[:a_out] != 0xEE).sum()) == 0 and int((dst[a_out + n + 16:] != 0xEE).sum()) == 0, ("parse wrote outside", n, a_in, a_out)
    if it % 20 == 19:
        print(f"{it + 1} iterations ok", flush=True)
print("fuzz ok")
