"""windows_batch_kernel by mode and window length on the batch shape of the all_windows leg: python tools/windows_modes_bench.py [n_seq]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
peak = 6548.5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for L, w, modes in ((99_999, 42, (0, 1, 2)), (33_333, 20, (0,)), (99_999, 20, (0,))):
    x = synth_torch(3, 0, n_seq * L, dev)
    nw = L - w + 1
    out = torch.empty(n_seq * nw * w + 16, dtype=torch.uint8, device=dev)
    for mode in modes:
        for _ in range(2):
            B.windows_batch_dev(x, n_seq, L, w, mode=mode, out=out)
        e0.record()
        for _ in range(5):
            B.windows_batch_dev(x, n_seq, L, w, mode=mode, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gb = n_seq * (L + nw * w) / 1e9
        print(f"n_seq={n_seq} L={L} w={w} mode={mode}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  frac {gb / ms * 1e3 / peak:.3f}", flush=True)
    del x, out
