"""Times qd_windows_all_batch_dev (the all_windows shape) on 1000 x 99 999-nt sequences."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n, L = 1000, 99_999
x = synth_torch(3, 0, n * L, dev)
cs, d, a = B.windows_all_batch_dev(x, n, L)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    B.windows_all_batch_dev(x, n, L, dna_out=d, aa_out=a)
e0.record()
for _ in range(5):
    B.windows_all_batch_dev(x, n, L, dna_out=d, aa_out=a)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"windows_all {ms:.3f} ms  {399796 * n / ms / 1e6:.1f} G windows/s  {n * (L + 12394072) / ms / 1e6:.0f} GB/s")
