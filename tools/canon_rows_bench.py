"""canonical_tile_kernel<3, 42, rows> on n_seq x 99 999-nt strict sequences: python tools/canon_rows_bench.py [n_seq]  (QD_CTILE_WPT_ROWS sweeps the tile size)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n_seq, L, w = (int(sys.argv[1]) if len(sys.argv) > 1 else 768), 99_999, 42
x = synth_torch(6, 0, n_seq * L, dev)
out = torch.empty(n_seq * (L - w + 1) * w + 16, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    B.canonical_windows_batch_dev(x, n_seq, L, w, out=out)
e0.record()
for _ in range(5):
    B.canonical_windows_batch_dev(x, n_seq, L, w, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
nwin = n_seq * (L - w + 1)
print(f"WPT={os.environ.get('QD_CTILE_WPT_ROWS', 'default')}: {ms:.3f} ms  {nwin / ms / 1e6:.1f} G windows/s  {nwin * 43 / ms / 1e6:.0f} GB/s  frac {nwin * 43 / ms / 1e6 / 6548.5:.3f}")
