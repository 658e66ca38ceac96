"""Randomised checks of qd_windows_all_batch_dev (nine launches chained as programmatic dependents) against the oracle's
windows_all, sequence by sequence, repeated back to back on one stream so that a call meets the tail of the call before it.
python tools/fuzz_windows_all.py [iterations] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ALPHA = np.frombuffer(b"ACGTacgt", dtype=np.uint8)


def check(a, n_seq, L, w_dna, w_aa, res):
    counts, dna_w, aa_w = res
    dna_w, aa_w = dna_w.cpu().numpy(), aa_w.cpu().numpy()
    nd = counts[0]
    want_all = [oracle.windows_all(oracle.masks_of(a[s]), w_dna, w_aa) for s in range(n_seq)]
    for s in range(n_seq):
        wd, _wa, _origin, cs = want_all[s]
        assert cs[:2] == counts[:2] and [c for c in cs[2:] if c] == [c for c in counts[2:] if c]
        assert dna_w[(s * nd) * w_dna:((s + 1) * nd) * w_dna].tobytes() == wd[:nd].tobytes(), ("fwd", n_seq, L, s)
        assert dna_w[(n_seq * nd + s * nd) * w_dna:(n_seq * nd + (s + 1) * nd) * w_dna].tobytes() == wd[nd:].tobytes(), ("rc", n_seq, L, s)
    aa_base = 0
    for f in range(6):
        cf = counts[2 + f]
        if not cf:
            continue
        for s in range(n_seq):
            wa, cs = want_all[s][1], want_all[s][3]
            lo = sum(cs[2:][:f])
            got_f = aa_w[(aa_base + s * cf) * w_aa:(aa_base + (s + 1) * cf) * w_aa]
            assert got_f.tobytes() == wa[lo:lo + cf].tobytes(), ("aa", f, n_seq, L, s)
        aa_base += n_seq * cf


for it in range(iters):
    cases = []
    for _ in range(int(rng.integers(2, 5))):  # several calls in flight on the stream before anything is read back
        n_seq = int(rng.integers(1, 60))
        L = int(rng.choice([rng.integers(3, 200), rng.integers(200, 4000), rng.integers(4000, 40000)]))
        w_dna, w_aa = (42, 20) if rng.integers(0, 2) else (int(rng.integers(1, 65)), int(rng.integers(1, 65)))
        if L > 4000:
            n_seq = min(n_seq, 6)
        a = ALPHA[rng.integers(0, len(ALPHA), (n_seq, L))]
        x = torch.from_numpy(a).cuda()
        cases.append((a, n_seq, L, w_dna, w_aa, x))
    B.dev_error_reset()
    results = [B.windows_all_batch_dev(x, n_seq, L, w_dna, w_aa) for (a, n_seq, L, w_dna, w_aa, x) in cases]
    for (a, n_seq, L, w_dna, w_aa, x), res in zip(cases, results):
        check(a, n_seq, L, w_dna, w_aa, res)
    if (it + 1) % 10 == 0:
        print(f"{it + 1} iterations ok", flush=True)
print("done")
