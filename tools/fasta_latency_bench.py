"""Host-call latency of the FASTA scanner and FASTA -> six frames on small and medium files: python tools/fasta_latency_bench.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (synthetic letters only)
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402

ctx = N.default_context(0)


def fasta(records, lines_per_record):
    body = oracle.synth(7, 0, records * lines_per_record * 70).reshape(records, lines_per_record, 70)
    parts = []
    for i in range(records):
        parts.append(b">r%d\n" % i)
        parts.append(b"\n".join(body[i, k].tobytes() for k in range(lines_per_record)) + b"\n")
    return b"".join(parts)


def timed(fn, reps):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


for records, lines, reps in ((1, 428, 200), (100, 14, 200), (10_000, 14, 50), (100_000, 28, 5)):
    text = fasta(records, lines)
    a = timed(lambda: B.fasta_scan(text, ctx=ctx), reps)
    b = timed(lambda: B.fasta_translate_frames(text, table=1, frames=6, ctx=ctx), reps)
    print(f"{records} records, {len(text)} bytes of text: scan {a * 1e6:.0f} us   scan + six frames {b * 1e6:.0f} us")
