"""Device-resident timing of qd_parse_dev: python tools/parse_sizes_bench.py [letters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402
from quickdna_b200.synth import synth_torch  # noqa: E402

peak = 6548.5
dev = torch.device("cuda")
ctx = N.default_context(0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps


for n in [int(a) for a in sys.argv[1:]] or [250_000_000, 1_000_000, 30_000]:
    x = synth_torch(3, 0, n, dev)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = timed(lambda: B.parse_dev(x, out=out, ctx=ctx))
    print(f"clean {n}: {ms * 1e3:.1f} us  {2 * n / ms / 1e6:.0f} GB/s  frac {2 * n / ms / 1e6 / peak:.4f}")
    y = x.clone()
    y[1000::61] = 0x20
    ms = timed(lambda: B.parse_dev(y, out=out, ctx=ctx))
    print(f"blank every 61st {n}: {ms * 1e3:.1f} us")
