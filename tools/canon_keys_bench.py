"""Times qd_canonical_window_keys_dev (both key kinds) and the materialising kernel on the bench's canonical workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import _native as N
from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n_seq, L, w = int(os.environ.get("N_SEQ", 4000)), 99_999, int(os.environ.get("W", 42))
x = synth_torch(6, 0, n_seq * L, dev)
nwin = n_seq * (L - w + 1)
out = torch.empty(nwin * 4, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rows = torch.empty(nwin * w + 16, dtype=torch.uint8, device=dev) if os.environ.get("ROWS", "1") == "1" else None
if rows is not None:
    for fwd in (False, True):
        for _ in range(3):
            B.canonical_windows_batch_dev(x, n_seq, L, w, out=rows, forward_only=fwd)
        e0.record()
        for _ in range(10):
            B.canonical_windows_batch_dev(x, n_seq, L, w, out=rows, forward_only=fwd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"rows   forward_only={fwd!s:5s} {ms:7.3f} ms  {nwin / ms / 1e6:7.1f} G windows/s  {nwin * (1 + w) / ms / 1e6:7.1f} GB/s")
    del rows
for kind, name in ((N.CANON_KEY_PACKED, "packed"), (N.CANON_KEY_FP64, "fp64")):
    for fwd in (False, True):
        for _ in range(3):
            B.canonical_window_keys_dev(x, n_seq, L, w, kind=kind, out=out, forward_only=fwd)
        e0.record()
        for _ in range(10):
            B.canonical_window_keys_dev(x, n_seq, L, w, kind=kind, out=out, forward_only=fwd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:6s} forward_only={fwd!s:5s} {ms:7.3f} ms  {nwin / ms / 1e6:7.1f} G windows/s")
