"""Multi-rank host logic on CPU: world_size-2 gloo.  Each rank takes its shard of a batch / of one long
sequence, processes it with the ORACLE standing in for the device path, and rank 0 checks that the
concatenation equals the unsharded result -- i.e. that the partition arithmetic needs no exchange."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from quickdna_b200 import shard
from quickdna_b200.synth import synth_np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- batch of variable-length reads, byte-balanced at sequence boundaries
        rng = np.random.default_rng(0)
        lens = rng.integers(0, 3000, 400).astype(np.uint64)
        offsets = np.zeros(len(lens) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offsets[1:])
        flat = synth_np(9, 0, int(offsets[-1]))
        parts = shard.partition_batch(offsets, world)
        a, b = parts[rank]
        mine = [b"|".join(oracle.all_frames_masks(0, oracle.masks_of(flat[int(offsets[i]):int(offsets[i + 1])])))
                for i in range(a, b)]
        gathered = [None] * world
        dist.all_gather_object(gathered, (a, b, mine))
        # ---- one long sequence cut at multiples of 48 with a 2-nt halo
        n = 100_003
        chrom = synth_np(10, 0, n)
        s, e, re_ = shard.partition_chromosome(n, world)[rank]
        ranges = shard.chromosome_frame_ranges(n, s, e)
        piece = chrom[s:re_]
        m = oracle.masks_of(piece)
        tab = oracle.tables()[:4096]
        local = {}
        for f in range(6):
            lo, hi = ranges[f]
            outv = bytearray()
            for idx in range(lo, hi):
                if f < 3:
                    p = 3 * idx + f
                    c = m[p - s:p - s + 3]
                    outv.append(tab[(int(c[0]) << 8) | (int(c[1]) << 4) | int(c[2])])
                else:
                    p = n - 3 - (3 * idx + (f - 3))
                    c = [oracle.revcomp_masks(m[p - s:p - s + 3])[j] for j in range(3)]
                    outv.append(tab[(int(c[0]) << 8) | (int(c[1]) << 4) | int(c[2])])
            local[f] = (lo, hi, bytes(outv))
        rc_piece = oracle.revcomp_bytes(chrom[s:e].tobytes())
        g2 = [None] * world
        dist.all_gather_object(g2, (s, e, local, rc_piece))
        if rank == 0:
            ok = True
            cover = sorted((x[0], x[1]) for x in gathered)
            ok &= cover[0][0] == 0 and cover[-1][1] == len(lens) and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
            sizes = [int(offsets[y] - offsets[x]) for x, y in cover]
            ok &= max(sizes) - min(sizes) <= 3000  # balanced by bytes to within one sequence
            joined = [r for part in sorted(gathered) for r in part[2]]
            want = [b"|".join(oracle.all_frames_masks(0, oracle.masks_of(flat[int(offsets[i]):int(offsets[i + 1])])))
                    for i in range(len(lens))]
            ok &= joined == want
            full = oracle.all_frames_masks(0, oracle.masks_of(chrom))
            for f in range(6):
                buf = bytearray(len(full[f]))
                seen = 0
                for (_s, _e, loc, _rc) in g2:
                    lo, hi, data = loc[f]
                    buf[lo:hi] = data
                    seen += hi - lo
                ok &= seen == len(full[f]) and bytes(buf) == full[f]
            rc_full = oracle.revcomp_bytes(chrom.tobytes())
            rc_join = bytearray(n)
            for (_s, _e, _loc, rcp) in g2:
                rc_join[n - _e:n - _s] = rcp  # shard [s, e) lands at [n-e, n-s)
            ok &= bytes(rc_join) == rc_full
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_needs_no_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_partition_helpers():
    offs = np.array([0, 100, 100, 400, 1000, 1001], dtype=np.uint64)
    for w in (1, 2, 3, 4, 8):
        parts = shard.partition_batch(offs, w)
        assert parts[0][0] == 0 and parts[-1][1] == 5 and all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert shard.partition_uniform(10_000_000, w)[-1][1] == 10_000_000
        chrom = shard.partition_chromosome(250_000_001, w)
        assert chrom[0][0] == 0 and chrom[-1][1] == 250_000_001
        for (s, e, r), nxt in zip(chrom, chrom[1:]):
            assert s % 48 == 0 and e % 48 == 0 and e == nxt[0] and r == e + 2


def test_chromosome_shard_placement():
    """A shard translated as a stand-alone sequence (its nucleotides + the 2-nt halo) lands in the whole sequence's
    frames by host-side placement alone; checked here with the oracle standing in for the device path."""
    rng = np.random.default_rng(61)
    iupac = np.frombuffer(b"ACGTacgtWMRYSKBVDHNwmryskbvdhn", dtype=np.uint8)

    def six(compact, length):  # the reference omits empty frames: put the returned ones back at their frame numbers
        present = [k for k in range(3) if length > k and (length - k) // 3 > 0]
        assert len(compact) == 2 * len(present)
        full = [b""] * 6
        for j, k in enumerate(present):
            full[k], full[3 + k] = compact[j], compact[len(present) + j]
        return full

    for n, world in ((100_003, 2), (300_001, 4), (48 * 7, 8), (50, 2), (99_999, 3), (5, 4), (3, 2), (97, 3), (49, 2), (4, 1), (96, 2)):
        seq = bytes(iupac[rng.integers(0, len(iupac), n)])
        whole = six(oracle.py_all_frames(11, seq) if n >= 3 else [], n)
        frames = [bytearray(len(f)) for f in whole]
        seen = [0] * 6
        for s, e, read_end in shard.partition_chromosome(n, world):
            if e <= s:
                continue
            piece = seq[s:read_end]
            local = six(oracle.py_all_frames(11, piece) if len(piece) >= 3 else [], len(piece))
            for lf, (gf, first, ln) in enumerate(shard.chromosome_shard_placement(n, s, e)):
                assert ln == len(local[lf])
                frames[gf][first:first + ln] = local[lf]
                seen[gf] += ln
        assert [bytes(f) for f in frames] == whole and seen == [len(f) for f in whole], (n, world)
