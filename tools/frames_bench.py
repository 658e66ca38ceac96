"""Device-resident timing of the frame kernels on uniform batches: python tools/frames_bench.py [reads] [len ...]
Prints ms and the fraction of the measured HBM copy peak for 1 / 3 / 6 frames (and strict, IUPAC input)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402
from quickdna_b200.synth import synth_iupac_np, synth_torch  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
lens = [int(a) for a in sys.argv[2:]] or [1000]
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6548.5
dev = torch.device("cuda")
ctx = N.default_context(0)
for L in lens:
    n = min(n_reads, 4_000_000_000 // L)
    x = synth_torch(3, 0, n * L, dev)
    out = torch.empty(n * 16 * B.runs(L) * 6, dtype=torch.uint8, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    variants = [("1 frame", dict(frames=1)), ("3 frames", dict(frames=3)), ("6 frames", dict(frames=6)), ("6 strict", dict(frames=6, strict=True))]
    order = os.environ.get("QD_BENCH_ORDER")  # e.g. "2,1,0,3": which variants, in which order
    if order:
        variants = [variants[int(i)] for i in order.split(",")]
    for name, kw in variants:
        f = kw["frames"]
        algo = n * (L + (2 if f == 6 else 1) * sum((L - k) // 3 for k in range(1 if f == 1 else 3)))
        for _ in range(3):
            B.translate_frames_uniform_dev(x, n, L, out=out, ctx=ctx, **kw)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(10):
            B.translate_frames_uniform_dev(x, n, L, out=out, ctx=ctx, **kw)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        print(f"L={L:6d} reads={n:9d} {name:9s} {ms:8.4f} ms  {algo / ms / 1e6:7.1f} GB/s  frac {algo / ms / 1e6 / peak:.4f}", flush=True)
    if L == 1000:
        base = torch.from_numpy(synth_iupac_np(5, 0, 50_000 * L)).to(dev)
        xa = base.repeat(20)
        algo = 1_000_000 * (L + 2 * sum((L - k) // 3 for k in range(3)))
        for _ in range(3):
            B.translate_frames_uniform_dev(xa, 1_000_000, L, out=out, table=11, frames=6, ctx=ctx)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(10):
            B.translate_frames_uniform_dev(xa, 1_000_000, L, out=out, table=11, frames=6, ctx=ctx)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        print(f"L={L:6d} reads=  1000000 6 IUPAC   {ms:8.4f} ms  {algo / ms / 1e6:7.1f} GB/s  frac {algo / ms / 1e6 / peak:.4f}", flush=True)
    del x, out
