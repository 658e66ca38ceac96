"""windows_kernel throughput: every 42-byte window of one long sequence, device-resident."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import _native as N
from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
ctx = N.default_context(0)
L = N.lib()
for n, w in ((100_000_000, 42), (100_000_000, 20)):
    x = synth_torch(3, 0, n, dev)
    out = torch.empty((n - w + 1) * w + 16, dtype=torch.uint8, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        rc = L.qd_windows_dev(ctx.handle, C.c_void_p(x.data_ptr()), n, w, C.c_void_p(out.data_ptr()), B._stream_ptr(torch))
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        rc = L.qd_windows_dev(ctx.handle, C.c_void_p(x.data_ptr()), n, w, C.c_void_p(out.data_ptr()), B._stream_ptr(torch))
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 3
    print(f"windows w={w}: {n - w + 1} windows  {ms:.3f} ms  {(n - w + 1) * (w + 1) / ms / 1e6:.0f} GB/s")
    del x, out
