"""Times batch.fasta_scan on ~200 MB of FASTA text (20 000 records of 143 lines x 70 nt), host in / host out."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_np

n_rec, lines_per, width = 20_000, 143, 70
body = synth_np(7, 0, n_rec * lines_per * width).reshape(n_rec, lines_per, width)
lines = np.concatenate([body, np.full((n_rec, lines_per, 1), 10, np.uint8)], axis=2).reshape(n_rec, -1)
hdr = np.frombuffer(b">synthetic record, 143 lines of 70 nt\n", dtype=np.uint8)
text = np.concatenate([np.broadcast_to(hdr, (n_rec, hdr.size)), lines], axis=1).reshape(-1).tobytes()
for i in range(4):
    t0 = time.perf_counter()
    scan = B.fasta_scan(text)
    dt = time.perf_counter() - t0
    print(f"call {i}: {dt * 1e3:7.1f} ms  {len(text) / dt / 1e9:5.2f} GB/s of text  ({len(scan)} records)")
