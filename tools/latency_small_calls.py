import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quickdna_b200 as qd
from quickdna_b200.synth import synth_np
g = qd.DnaSequence(synth_np(1, 0, 29_901, b"acgt").tobytes())
for name, fn in (("translate", lambda: g.translate(1)), ("revcomp", lambda: g.reverse_complement()), ("all_frames", lambda: g.translate_all_frames(1))):
    for _ in range(20): fn()
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(200): fn()
        print(name, rep, (time.perf_counter() - t0) / 200 * 1e3, "ms")
