"""Small fixed workload for ncu captures: six-frame translate (1 M x 1 kb reads) and revcomp (250 Mb)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

which = sys.argv[1] if len(sys.argv) > 1 else "frames"
dev = torch.device("cuda:0")
if which == "frames":
    n_reads, L = 2_000_000, 1000
    x = synth_torch(3, 0, n_reads * L, dev)
    out = torch.empty(n_reads * 16 * B.runs(L) * 6, dtype=torch.uint8, device=dev)
    for _ in range(4):
        B.translate_frames_uniform_dev(x, n_reads, L, out=out)
elif which == "frames1":
    n_reads, L = 4_000_000, 1000
    x = synth_torch(3, 0, n_reads * L, dev)
    out = torch.empty(n_reads * 16 * B.runs(L), dtype=torch.uint8, device=dev)
    for _ in range(4):
        B.translate_frames_uniform_dev(x, n_reads, L, out=out, frames=1)
elif which == "frames3":
    n_reads, L = 4_000_000, 1000
    x = synth_torch(3, 0, n_reads * L, dev)
    out = torch.empty(n_reads * 16 * B.runs(L) * 3, dtype=torch.uint8, device=dev)
    for _ in range(4):
        B.translate_frames_uniform_dev(x, n_reads, L, out=out, frames=3)
elif which == "frames_iupac":
    from quickdna_b200.synth import synth_iupac_np
    import numpy as np
    n_reads, L = 1_000_000, 1000
    x = torch.from_numpy(np.tile(synth_iupac_np(5, 0, 50_000 * L), 20)).to(dev)
    out = torch.empty(n_reads * 16 * B.runs(L) * 6, dtype=torch.uint8, device=dev)
    for _ in range(4):
        B.translate_frames_uniform_dev(x, n_reads, L, out=out)
else:
    n = 250_000_000
    x = synth_torch(2, 0, n, dev)
    out = torch.empty_like(x)
    for _ in range(4):
        B.revcomp_dev(x, out=out)
torch.cuda.synchronize()
print("done", which)
