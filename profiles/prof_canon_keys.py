"""Small fixed workload for ncu captures of canonical_keys_kernel: 400 x 99 999-nt strict sequences, 42-nt windows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import _native as N
from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

kind = N.CANON_KEY_FP64 if (len(sys.argv) > 1 and sys.argv[1] == "fp64") else N.CANON_KEY_PACKED
dev = torch.device("cuda:0")
n_seq, L, w = 400, 99_999, 42
x = synth_torch(6, 0, n_seq * L, dev)
out = torch.empty(n_seq * (L - w + 1) * 4, dtype=torch.int32, device=dev)
for _ in range(3):
    B.canonical_window_keys_dev(x, n_seq, L, w, kind=kind, out=out)
torch.cuda.synchronize()
print("done", kind)
