"""Times qd_parse_dev on 250 M letters: without whitespace (the usual case) and with a blank every 61st byte."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n = 250_000_000
x = synth_torch(2, 0, n, dev)
y = x.clone()
y[60::61] = 0x20
out = torch.empty(n + 16, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, src in (("no whitespace", x), ("blank every 61st", y), ("no whitespace, out shifted 1", x)):
    o = out[1:] if "shifted" in name else out
    for _ in range(3):
        B.parse_dev(src, out=o)
    e0.record()
    for _ in range(20):
        B.parse_dev(src, out=o)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:30s} {ms:7.4f} ms  {n / ms / 1e9:6.3f} Tnt/s")
