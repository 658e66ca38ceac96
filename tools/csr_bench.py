"""Device-resident timing of the variable-length (CSR) six-frame path, scan and map kernels included:
python tools/csr_bench.py [reads]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402
from quickdna_b200.synth import synth_torch  # noqa: E402

peak = 6548.5
dev = torch.device("cuda")
ctx = N.default_context(0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=7):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps


def case(name, lens):
    offs = torch.zeros(len(lens) + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens, 0)
    tot = int(offs[-1].item())
    x = synth_torch(3, 0, tot, dev)
    runs = int(((lens + 47) // 48).sum().item())
    out = torch.empty(runs * 96, dtype=torch.uint8, device=dev)
    offs_d = offs.to(dev)
    algo = int((lens + 2 * (lens // 3 + (lens - 1).clamp(min=0) // 3 + (lens - 2).clamp(min=0) // 3)).sum().item())
    l0 = ctx.launch_count
    ms = timed(lambda: B.translate_frames_csr_dev(x, offs_d, out=out, table=1, frames=6, ctx=ctx))
    print(f"{name}: {len(lens)} sequences, {tot} nt  {ms * 1e3:.1f} us  {algo / ms / 1e6:.0f} GB/s  frac {algo / ms / 1e6 / peak:.4f}  "
          f"launches/call {(ctx.launch_count - l0) // 10}")


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
g = torch.Generator().manual_seed(4)
case("reads of 500..1500 nt", torch.randint(500, 1501, (n,), generator=g, dtype=torch.int64))
case("100 reads of 500..1500 nt", torch.randint(500, 1501, (100,), generator=g, dtype=torch.int64))
case("one chromosome among reads", torch.cat([torch.randint(500, 1501, (200_000,), generator=g, dtype=torch.int64), torch.tensor([250_000_000]),
                                             torch.randint(500, 1501, (200_000,), generator=g, dtype=torch.int64)]))
case("contigs of 50..400 kb", torch.randint(50_000, 400_001, (8_000,), generator=g, dtype=torch.int64))
case("long reads of 5..50 kb", torch.randint(5_000, 50_001, (100_000,), generator=g, dtype=torch.int64))
