"""Times translate_all_frames on reads of 500..1500 nt (CSR) and on uniform 1000-nt reads of the same total size."""
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
out = torch.empty(max(runs, n * 21) * 96, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("csr 500..1500", lambda: B.translate_frames_csr_dev(x, offs_d, out=out)),
                 ("uniform 1000", lambda: B.translate_frames_uniform_dev(x, n, 1000, out=out))):
    for _ in range(3):
        fn()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nt = tot if name.startswith("csr") else n * 1000
    print(f"{name:14s} {ms:7.3f} ms  {nt / ms / 1e9:6.3f} Tnt/s  {nt * 2.996 / ms / 1e6:7.1f} GB/s")
