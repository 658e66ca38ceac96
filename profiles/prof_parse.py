"""Small fixed workload for ncu captures of the parse kernels: 250 M clean letters, device-resident."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n = 250_000_000
x = synth_torch(2, 0, n, dev)
out = torch.empty(n + 16, dtype=torch.uint8, device=dev)
for _ in range(4):
    B.parse_dev(x, out=out)
torch.cuda.synchronize()
print("done")
