"""Variable-length six-frame batch (lengths 500..1500) for ncu launch lists: scan + tile map + frame kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n = 4_000_000
g = torch.Generator(device="cpu").manual_seed(4)
lens = torch.randint(500, 1501, (n,), generator=g, dtype=torch.int64)
offs = torch.zeros(n + 1, dtype=torch.int64)
offs[1:] = torch.cumsum(lens, 0)
tot = int(offs[-1])
x = synth_torch(3, 0, tot, dev)
offs_d = offs.to(dev)
runs = int(((lens + 47) // 48).sum())
out = torch.empty(runs * 96, dtype=torch.uint8, device=dev)
for _ in range(3):
    B.translate_frames_csr_dev(x, offs_d, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
B.translate_frames_csr_dev(x, offs_d, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", tot)
