"""Variable-length six-frame batch (1 M reads of 500..1500 nt), for ncu launch lists / captures of the scan + map kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import _native as N
from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
ctx = N.default_context(0)
g = torch.Generator().manual_seed(4)
lens = torch.randint(500, 1501, (1_000_000,), generator=g, dtype=torch.int64)
offs = torch.zeros(len(lens) + 1, dtype=torch.int64)
offs[1:] = torch.cumsum(lens, 0)
x = synth_torch(3, 0, int(offs[-1]), dev)
out = torch.empty(int(((lens + 47) // 48).sum()) * 96, dtype=torch.uint8, device=dev)
offs_d = offs.to(dev)
for _ in range(2):
    B.translate_frames_csr_dev(x, offs_d, out=out, table=1, frames=6, ctx=ctx)
torch.cuda.synchronize()
torch.cuda.profiler.start()
B.translate_frames_csr_dev(x, offs_d, out=out, table=1, frames=6, ctx=ctx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
