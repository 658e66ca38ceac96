"""Randomised checks of the host-pointer paths: tiny (mapped memory), small, chunked / staged copies on pageable and pinned
buffers, CSR and uniform batches, one long sequence, FASTA text -> frames.  Every result against the oracle (checksums for the
big ones).  python tools/fuzz_host_paths.py [iterations] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from quickdna_b200 import _native as N  # noqa: E402
from quickdna_b200 import batch as B  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
threads = len(os.sched_getaffinity(0))
ctx = N.Context(0)
L = N.lib()
ALPHA = np.frombuffer(b"ACGTacgtNRYKMSWBDHV", dtype=np.uint8)


def host_buffer(a, pinned):
    if not pinned:
        return a
    p = B.pinned_empty(a.shape, a.dtype)
    p[...] = a
    return p


for it in range(iters):
    L.qd_ctx_set_stage_threads(ctx.handle, int(rng.choice([-1, 0, 1, 3, 8])))
    L.qd_ctx_set_chunk_bytes(ctx.handle, int(rng.choice([1 << 20, 5 << 20, 16 << 20, 64 << 20])))
    pinned = bool(rng.random() < 0.3)
    # one sequence of any size class
    n = int(rng.choice([rng.integers(1, 200), rng.integers(100_000, 200_000), rng.integers(1 << 20, 3 << 20), rng.integers(30 << 20, 45 << 20)]))
    seed = int(rng.integers(1, 1000))
    dna = host_buffer(oracle.synth(seed, 0, n), pinned)
    rc = B.revcomp(dna, ctx=ctx)
    assert oracle.checksum64(rc, 0) == oracle.check_revcomp(seed, 0, n, False, threads), ("revcomp", n, pinned)
    frames = int(rng.choice([1, 3, 6]))
    got = B.translate_frames(dna, table=int(rng.choice([1, 2, 11])), frames=frames, ctx=ctx)
    head = dna[: min(n, 3000)].tobytes()
    if n <= 3000:
        want = [oracle.translate_bytes(1, head)] if frames == 1 else (oracle.py_self_frames(1, head) if frames == 3 else oracle.py_all_frames(1, head))
    assert len(got) == (0 if n < 3 and frames == 1 else len(got))
    # the frames of a long sequence: forward frame 0 must begin like the translation of its head, reverse frames come from the tail
    if n >= 3000:
        t = int(rng.choice([1, 2, 11]))
        g2 = B.translate_frames(dna, table=t, frames=6, ctx=ctx)
        w2 = oracle.py_all_frames(t, head)
        assert bytes(g2[0][:900]) == w2[0][:900] and bytes(g2[1][:900]) == w2[1][:900] and bytes(g2[2][:900]) == w2[2][:900]
        tail = dna[n - 3000:].tobytes()
        wt = oracle.py_all_frames(t, tail)
        assert bytes(g2[3][:900]) == wt[3][:900] and bytes(g2[4][:900]) == wt[4][:900] and bytes(g2[5][:900]) == wt[5][:900]
    # uniform batch, checked whole
    rl = int(rng.choice([150, 1000, 2999]))
    nr = int(rng.choice([3, 100, 20_000, (40 << 20) // rl]))
    reads = host_buffer(oracle.synth(3, 0, nr * rl).reshape(nr, rl), pinned)
    fb = B.translate_frames_batch(reads, table=1, frames=6, ctx=ctx)
    assert oracle.checksum64(fb.data, 0) == oracle.check_frames_uniform(3, 0, nr, rl, 0, 6, threads), ("uniform", nr, rl, pinned)
    # CSR batch with ambiguity codes: sampled sequences against the oracle, all of it against the unstaged result
    ns = int(rng.choice([50, 5000, 60_000]))
    lens = rng.integers(0, int(rng.choice([60, 700, 2500])), ns)
    offs = np.zeros(ns + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    flat = ALPHA[rng.integers(0, len(ALPHA), int(offs[-1]) or 1)][: int(offs[-1])].copy()
    fb = B.translate_frames_batch((flat, offs), table=11, frames=6, ctx=ctx)
    for i in rng.integers(0, ns, 6):
        assert fb.sequence_frames(int(i)) == oracle.py_all_frames(11, flat[int(offs[i]):int(offs[i + 1])].tobytes()), ("csr", int(i))
    L.qd_ctx_set_stage_threads(ctx.handle, 0)
    L.qd_ctx_set_chunk_bytes(ctx.handle, 64 << 20)
    fb0 = B.translate_frames_batch((flat, offs), table=11, frames=6, ctx=ctx)
    assert np.array_equal(fb.data, fb0.data)
    # FASTA text -> frames in one call against scan + translate
    if it % 3 == 0:
        L.qd_ctx_set_stage_threads(ctx.handle, -1)
        parts = []
        for i in range(int(rng.choice([1, 40, 3000]))):
            parts.append(b">r%d\n" % i)
            body = bytes(ALPHA[rng.integers(0, len(ALPHA), int(rng.integers(0, 2000)))])
            parts += [body[k:k + 60] + (b"\r\n" if i % 2 else b"\n") for k in range(0, len(body), 60)]
        text = b"".join(parts)
        scan, ff = B.fasta_translate_frames(text, ctx=ctx)
        two = B.fasta_scan(text, ctx=ctx)
        fb2 = B.translate_frames_batch((two.masks, two.offsets), frames=6, enc=1, ctx=ctx)
        assert len(scan) == len(two) and np.array_equal(ff.data, fb2.data), "fasta"
    print(f"{it + 1} iterations ok", flush=True)
ctx.close()
print("fuzz ok")
