"""Device-resident timing of the expansion path: python tools/exp_bench.py [windows]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402
from quickdna_b200.synth import synth_torch  # noqa: E402

n_win, w = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000, 42
peak = 6548.5
dev = torch.device("cuda")
ctx = N.default_context(0)
win = synth_torch(3, 0, n_win * w, dev).view(n_win, w)
g2 = torch.Generator(device=dev).manual_seed(11)
codes = torch.tensor(list(b"WMYHRKDSVBN"), dtype=torch.uint8, device=dev)
idx = torch.arange(n_win, device=dev)
p1 = torch.randint(0, w, (n_win,), generator=g2, device=dev)
win[idx, p1] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
p1 = (p1 + 1 + torch.randint(0, w - 1, (n_win,), generator=g2, device=dev)) % w
win[idx, p1] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
del idx, p1
out = torch.empty(n_win * 8 * w, dtype=torch.uint8, device=dev)
c, oi, _ = B.expansion_windows_dev(win, 16, out=out, ctx=ctx)
rows = int(oi[n_win].item())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps


ms = timed(lambda: B.expansion_windows_dev(win, 16, out=out, ctx=ctx))
algo = n_win * (w + 16) + rows * w
print(f"single call   {n_win} windows -> {rows} rows  {ms:.3f} ms  {algo / ms / 1e6:.1f} GB/s  frac {algo / ms / 1e6 / peak:.4f}")
L = N.lib()
for chunk in (2_000_000, 4_000_000, 8_000_000):
    if chunk > n_win:
        continue
    slot = (L.qd_expansion_chunk_bytes(chunk, w, 16) + 255) & ~255
    ring = torch.empty(2 * slot, dtype=torch.uint8, device=dev)
    ms = timed(lambda: B.expansion_windows_chunked_dev(win, ring, chunk, 16, checksums=False, ctx=ctx), reps=3)
    algo = n_win * w + rows * w
    print(f"ring chunks of {chunk}  {ms:.3f} ms  {algo / ms / 1e6:.1f} GB/s  frac {algo / ms / 1e6 / peak:.4f}")
    del ring
