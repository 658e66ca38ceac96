"""e2e leg of bench.py at several pipeline chunk sizes (host pinned in, host pinned out)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from quickdna_b200 import _native as N
from quickdna_b200 import batch as B
from quickdna_b200.synth import synth_torch

dev = torch.device("cuda:0")
ctx = N.default_context(0)
L = N.lib()
er, RL = 2_000_000, 1000
per = 16 * B.runs(RL) * 6
x = synth_torch(3, 0, er * RL, dev)
h_in = torch.empty(er * RL, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(er * per, dtype=torch.uint8, pin_memory=True)
h_in.copy_(x)
torch.cuda.synchronize()
err = N.QdErr()
for mib in (8, 16, 32, 64, 128, 256):
    L.qd_ctx_set_chunk_bytes(ctx.handle, mib << 20)
    for rep in range(2):
        t0 = time.perf_counter()
        rc = L.qd_translate_frames_batch_uniform(ctx.handle, 1, 0, 0, 6, C.c_void_p(h_in.data_ptr()), er, RL, C.c_void_p(h_out.data_ptr()), C.byref(err))
        dt = time.perf_counter() - t0
    assert rc == 0
    print(f"chunk {mib:4d} MiB  {dt * 1e3:7.2f} ms  {er * RL / dt / 1e9:6.2f} Gnt/s  D2H {er * per / dt / 1e9:5.1f} GB/s")
