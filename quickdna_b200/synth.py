"""Counter-based synthetic sequence generators: nt[i] = ALPHABET[mix64(seed + i) % len].

The same function exists for numpy (host, tests, CPU baseline) and torch (device-side generation for
the benchmark) so every shard of every rank can regenerate its slice without shipping data.
"""

from __future__ import annotations

import numpy as np

_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_GOLD = 0x9E3779B97F4A7C15


def mix64_np(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(_GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        return z ^ (z >> np.uint64(31))


def synth_np(seed: int, start: int, count: int, alphabet: bytes = b"ACGT") -> np.ndarray:
    idx = np.arange(start, start + count, dtype=np.uint64) + np.uint64((seed * 0x1000003D) & 0xFFFFFFFFFFFFFFFF)
    h = mix64_np(idx)
    lut = np.frombuffer(alphabet, dtype=np.uint8)
    return lut[((h >> np.uint64(33)) % np.uint64(len(alphabet))).astype(np.int64)]


def synth_iupac_np(seed: int, start: int, count: int, ambiguous_frac: float = 0.05, lower_frac: float = 0.5) -> np.ndarray:
    """Config 4: ACGT with `ambiguous_frac` of positions replaced by a code from WMRYSKBVDHN, `lower_frac` lower case."""
    base = synth_np(seed, start, count)
    idx = np.arange(start, start + count, dtype=np.uint64)
    h2 = mix64_np(idx + np.uint64(0xA5A5A5A5 + seed))
    amb = ((h2 & np.uint64(0xFFFF)).astype(np.float64) / 65536.0) < ambiguous_frac
    codes = np.frombuffer(b"WMRYSKBVDHN", dtype=np.uint8)
    base = np.where(amb, codes[((h2 >> np.uint64(20)) % np.uint64(11)).astype(np.int64)], base)
    lower = (((h2 >> np.uint64(40)) & np.uint64(0xFFFF)).astype(np.float64) / 65536.0) < lower_frac
    return np.where(lower, base | np.uint8(0x20), base).astype(np.uint8)


def synth_torch(seed: int, start: int, count: int, device, alphabet: bytes = b"ACGT", chunk: int = 1 << 27):
    """Device-side generator, bit-identical to synth_np (int64 arithmetic wraps like uint64)."""
    import torch

    def s64(v):  # reinterpret a uint64 constant as int64
        v &= 0xFFFFFFFFFFFFFFFF
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(z, k):  # logical shift right on int64
        return (z >> k) & ((1 << (64 - k)) - 1)

    out = torch.empty(count, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(alphabet), dtype=torch.uint8, device=device)
    off = (seed * 0x1000003D) & 0xFFFFFFFFFFFFFFFF
    for c0 in range(0, count, chunk):
        c1 = min(count, c0 + chunk)
        z = torch.arange(start + c0, start + c1, dtype=torch.int64, device=device) + s64(off) + s64(_GOLD)
        z = (z ^ lsr(z, 30)) * s64(_M1)
        z = (z ^ lsr(z, 27)) * s64(_M2)
        z = z ^ lsr(z, 31)
        out[c0:c1] = lut[lsr(z, 33) % len(alphabet)]
        del z
    return out
