"""Randomised checks of variable-length batches (the scan whose last pass writes the tile and 32-run maps, long sequences
mapped by a warp or a CTA, the frame kernel behind them as a programmatic dependent): every sequence against the oracle.
python tools/fuzz_csr.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = N.Context(0)
ALPHA = np.frombuffer(b"ACGTacgtNRYKMSWBDHV", dtype=np.uint8)
TABLES = (1, 2, 4, 11, 22)


def lengths():
    kind = rng.integers(0, 5)
    n = int(rng.integers(1, 400))
    if kind == 0:
        return rng.integers(0, 8, n)
    if kind == 1:
        return rng.integers(0, 3000, n)
    if kind == 2:  # a few long ones among reads: thread / warp / CTA tiers of the map
        a = rng.integers(0, 2000, n)
        for _ in range(int(rng.integers(1, 4))):
            a[rng.integers(0, n)] = int(rng.choice([48 * 512, 48 * 512 + 1, 48 * 513, 60_000, 48 * 65536, 48 * 65536 + 1, 3_400_000])) + int(rng.integers(0, 3))
        return a
    if kind == 3:
        return rng.integers(20_000, 200_000, min(n, 40))
    return np.full(n, int(rng.integers(0, 5000)))


for it in range(iters):
    lens = np.asarray(lengths(), dtype=np.int64)
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    flat = ALPHA[rng.integers(0, 4 if rng.integers(0, 2) else len(ALPHA), int(offs[-1]))]
    table = int(rng.choice(TABLES))
    frames = int(rng.choice((1, 3, 6)))
    if rng.integers(0, 3) == 0:
        N.lib().qd_ctx_set_big_tile_min_runs(ctx.handle, 0)
    else:
        N.lib().qd_ctx_set_big_tile_min_runs(ctx.handle, 2 * 1024 * N.lib().qd_ctx_sm_count(ctx.handle))
    res = B.translate_frames_batch((flat, offs), table=table, frames=frames, ctx=ctx)
    pick = range(len(lens)) if len(lens) <= 64 else sorted(set(rng.integers(0, len(lens), 48).tolist()) | set(np.argsort(lens)[-4:].tolist()))
    for i in pick:
        r = flat[int(offs[i]):int(offs[i + 1])].tobytes()
        if frames == 6:
            want = oracle.py_all_frames(table, r) if len(r) >= 3 else []
        elif frames == 3:
            want = oracle.py_self_frames(table, r)
        else:
            want = [oracle.translate_bytes(table, r)] if len(r) >= 3 else []
        got = res.sequence_frames(i)
        assert got == want, (it, i, len(r), table, frames)
    if (it + 1) % 10 == 0:
        print(f"{it + 1} iterations ok", flush=True)
print("done")
