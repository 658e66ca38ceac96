"""all_windows shape for 1000 x 99 999-nt sequences, for ncu launch lists / captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n, L = 1000, 99_999
x = synth_torch(3, 0, n * L, dev)
for _ in range(2):
    cs, d, a = B.windows_all_batch_dev(x, n, L)
torch.cuda.synchronize()
torch.cuda.profiler.start()
B.windows_all_batch_dev(x, n, L, dna_out=d, aa_out=a)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
