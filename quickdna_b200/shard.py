"""Multi-GPU partitioning: independent units, no data-path collective.

* a batch of sequences is cut at SEQUENCE boundaries, balanced by BYTES (not by count);
* one long sequence is cut at offsets that are multiples of 48 (a whole number of runs and of
  codons), each shard reading a 2-nt right halo; forward frame k of shard j is the contiguous
  range [start_j/3, ...) of the global frame, reverse frames use the global length.

Pure host arithmetic (numpy); used by bench.py and by callers that spread work over the GPUs of
one box, one process / one stream per GPU.
"""

from __future__ import annotations

from typing import List, Tuple

import numpy as np

RUN_NT = 48


def partition_batch(offsets: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """[(first_seq, end_seq)] per rank: cut points are the sequence boundaries nearest k*total/world."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n_seq = len(offsets) - 1
    total = int(offsets[-1] - offsets[0])
    cuts = [0]
    for k in range(1, world_size):
        target = int(offsets[0]) + (total * k) // world_size
        j = int(np.searchsorted(offsets, np.uint64(target), side="left"))
        if j > 0 and j <= n_seq and target - int(offsets[j - 1]) < int(offsets[min(j, n_seq)]) - target:
            j -= 1
        cuts.append(min(max(j, cuts[-1]), n_seq))
    cuts.append(n_seq)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def partition_uniform(n_seq: int, world_size: int) -> List[Tuple[int, int]]:
    return [(n_seq * r // world_size, n_seq * (r + 1) // world_size) for r in range(world_size)]


def partition_chromosome(n: int, world_size: int) -> List[Tuple[int, int, int]]:
    """[(start, end, read_end)] per rank for one n-nt sequence: start/end multiples of 48 (end = n for
    the last rank); the shard reads [start, read_end) with read_end = min(n, end + 2)."""
    runs = (n + RUN_NT - 1) // RUN_NT
    out = []
    for r in range(world_size):
        s = (runs * r // world_size) * RUN_NT
        e = min(n, (runs * (r + 1) // world_size) * RUN_NT)
        out.append((min(s, n), e, min(n, e + 2) if e > s else e))
    return out


def chromosome_frame_ranges(n: int, start: int, end: int):
    """Residue index ranges a shard [start, end) of an n-nt sequence produces.

    Returns {frame: (first, stop)} for forward frames 0..2 and reverse frames 3..5 (global indices).
    A codon start p in [start, min(end, n-2)) is forward frame p%3 residue p//3 and reverse frame
    (n-3-p)%3 residue (n-3-p)//3.
    """
    hi = min(end, n - 2)
    res = {}
    for k in range(3):
        ps = [p for p in range(start, min(start + 3, hi)) if p % 3 == k]
        if not ps or hi <= start:
            res[k] = (0, 0)
            continue
        first = ps[0]
        last = first + 3 * ((hi - 1 - first) // 3)
        res[k] = (first // 3, last // 3 + 1)
    for k in range(3):
        qs_hi = n - 3 - start  # largest q
        qs_lo = n - 3 - (hi - 1)  # smallest q
        if hi <= start:
            res[3 + k] = (0, 0)
            continue
        lo = qs_lo + ((k - qs_lo) % 3)
        if lo > qs_hi:
            res[3 + k] = (0, 0)
            continue
        top = qs_hi - ((qs_hi - k) % 3)
        res[3 + k] = (lo // 3, top // 3 + 1)
    return res


def chromosome_shard_placement(n: int, start: int, end: int):
    """Where the frames of ONE shard land in the frames of the whole n-nt sequence.

    The shard [start, end) (start a multiple of 48; end a multiple of 48 or n) is translated as an ordinary
    stand-alone sequence made of its nucleotides plus the 2-nt halo: piece = seq[start:min(n, end + 2)].  With
    n_local = len(piece), every codon start of the whole sequence in [start, min(end, n - 2)) is a codon start of the
    piece and vice versa, so nothing is exchanged between shards; only the frame NUMBERS and residue offsets differ:

      forward local frame k  -> global frame k, first residue at index start // 3
      reverse local frame k  -> global frame (k + d) % 3, first residue at local index 0 = global index (k + d) // 3,
                                with d = n - (start + n_local)  (nucleotides of the sequence to the right of the piece)

    Returns [(global_frame 0..5, global_first_index, length)] for local frames 0..5 (length 0 = frame absent).
    """
    read_end = min(n, end + 2)
    n_local = read_end - start
    d = n - read_end
    out = []
    for k in range(3):
        ln = (n_local - k) // 3 if n_local > k else 0
        out.append((k, start // 3, ln))
    for k in range(3):
        ln = (n_local - k) // 3 if n_local > k else 0
        out.append((3 + (k + d) % 3, (k + d) // 3, ln))
    return out
