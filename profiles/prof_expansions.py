"""Expansion windows (20 M windows of 42 nt, 2 ambiguity codes each, cap 16) for ncu launch lists / captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n_win, w = 20_000_000, 42
win = synth_torch(3, 0, n_win * w, dev).view(n_win, w)
g2 = torch.Generator(device=dev).manual_seed(11)
codes = torch.tensor(list(b"WMYHRKDSVBN"), dtype=torch.uint8, device=dev)
p1 = torch.randint(0, w, (n_win,), generator=g2, device=dev)
p2 = (p1 + 1 + torch.randint(0, w - 1, (n_win,), generator=g2, device=dev)) % w
idx = torch.arange(n_win, device=dev)
win[idx, p1] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
win[idx, p2] = codes[torch.randint(0, 11, (n_win,), generator=g2, device=dev)]
out = torch.empty(n_win * 8 * w, dtype=torch.uint8, device=dev)
for _ in range(2):
    B.expansion_windows_dev(win, 16, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
B.expansion_windows_dev(win, 16, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
