"""Small fixed workload for ncu captures of canonical_tile_kernel<3, 42, rows>: 400 x 99 999-nt strict sequences, 42-nt windows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n_seq, L, w = 400, 99_999, 42
x = synth_torch(6, 0, n_seq * L, dev)
out = torch.empty(n_seq * (L - w + 1) * w + 16, dtype=torch.uint8, device=dev)
for _ in range(3):
    B.canonical_windows_batch_dev(x, n_seq, L, w, out=out)
torch.cuda.synchronize()
print("done")
