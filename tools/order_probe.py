"""Why does a frame kernel run slower after ANOTHER frame kernel ran in the process?  Per-launch CUDA-event times of the
three-frame kernel before / after other work (one-frame kernel, reverse complement, a torch copy, sleep, L2 flush)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402
from quickdna_b200.synth import synth_torch  # noqa: E402

dev = torch.device("cuda")
ctx = N.default_context(0)
n, L = 4_000_000, 1000
x = synth_torch(3, 0, n * L, dev)
out = torch.empty(n * 16 * 21 * 6, dtype=torch.uint8, device=dev)
frames = int(os.environ.get("PROBE_FRAMES", "3"))


import pynvml

pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)


def clocks():
    return (f"sm {pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM)} mem {pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_MEM)} MHz "
            f"{pynvml.nvmlDeviceGetPowerUsage(H) / 1000:.0f} W reasons {pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(H):#x} pstate {pynvml.nvmlDeviceGetPerformanceState(H)}")


def probe(label, reps=6):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for i in range(reps):
        B.translate_frames_uniform_dev(x, n, L, out=out, frames=frames, ctx=ctx)
        evs[i + 1].record()
    during = clocks()
    torch.cuda.synchronize()
    print(f"{label:32s}", " ".join(f"{evs[i].elapsed_time(evs[i + 1]):6.3f}" for i in range(reps)), "|", during, flush=True)


probe("fresh")
probe("again")
torch.cuda.synchronize(); time.sleep(0.5)
probe("after sleep")
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
big.fill_(1); torch.cuda.synchronize()
probe("after torch fill 1 GB")
B.revcomp_dev(x[:250_000_000], out=out[:250_000_000], ctx=ctx); torch.cuda.synchronize()
probe("after revcomp kernel")
other = int(os.environ.get("PROBE_OTHER", "1"))
B.translate_frames_uniform_dev(x, n, L, out=out, frames=other, ctx=ctx); torch.cuda.synchronize()
probe(f"after ONE {other}-frame launch")
time.sleep(0.5)
probe("after sleep")
big.fill_(2); torch.cuda.synchronize()
probe("after torch fill")
ctx2 = N.Context(0)
for _ in range(2):
    B.translate_frames_uniform_dev(x, n, L, out=out, frames=frames, ctx=ctx2)
torch.cuda.synchronize()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
evs[0].record()
B.translate_frames_uniform_dev(x, n, L, out=out, frames=frames, ctx=ctx2); evs[1].record()
B.translate_frames_uniform_dev(x, n, L, out=out, frames=frames, ctx=ctx2); evs[2].record()
torch.cuda.synchronize()
print("new context", evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]))
