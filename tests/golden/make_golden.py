#!/usr/bin/env python
"""Regenerates tests/golden/hotpath_golden.json.

Run in the BUILD container (where /root/reference exists): `python tests/golden/make_golden.py`.
Outputs come from the pinned CPU oracle (oracle/, itself checked against the reference's tables.dat and the
reference's own test vectors in tests/test_oracle_golden.py).  Two kinds of case:

  * synthetic inputs regenerated from (seed, n) by quickdna_b200.synth on any machine, expected outputs stored
    as sha256 (or literally when short) -- these travel to the GPU box, which has no /root/reference;
  * benchmarks/covid.txt of the reference (BASELINE config 1): input and outputs stored as sha256 only, checked
    in this container against the pins SURVEY.md section 8c derived independently.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from quickdna_b200.synth import synth_iupac_np, synth_np  # noqa: E402

REFERENCE = "/root/reference"


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


def make_input(kind: str, seed: int, n: int) -> bytes:
    return (synth_np(seed, 0, n) if kind == "acgt" else synth_iupac_np(seed, 0, n)).tobytes()


def expected(kind: str, seed: int, n: int) -> dict:
    dna = make_input(kind, seed, n)
    out = {"kind": kind, "seed": seed, "n": n, "input_sha256": sha(dna)}
    for table in (1, 2, 11, 22, 33):
        t = oracle.translate_bytes(table, dna)
        out[f"translate_t{table}"] = t.decode() if n <= 64 else sha(t)
    frames = oracle.py_all_frames(1, dna) if n >= 3 else []
    out["all_frames_t1"] = [f.decode() for f in frames] if n <= 64 else sha(b"\n".join(frames))
    out["all_frames_lens"] = [len(f) for f in frames]
    rc = oracle.revcomp_bytes(dna)
    out["revcomp"] = rc.decode() if n <= 64 else sha(rc)
    if kind == "acgt" and n >= 42:
        out["canonical_windows_42"] = sha(oracle.canonical_windows(oracle.masks_of(dna), 42).tobytes())
        out["canonical"] = sha(oracle.canonical(oracle.masks_of(dna)).tobytes())
    if kind == "iupac":
        try:
            oracle.translate_bytes(1, dna, strict=True)
            out["strict_error"] = None
        except oracle.OracleError as e:
            out["strict_error"] = {"kind": e.kind, "byte": e.byte, "pos": e.pos, "message": str(e)}
    return out


def main():
    cases = [expected(kind, seed, n) for kind, seed, n in [
        ("acgt", 1, 0), ("acgt", 1, 2), ("acgt", 2, 3), ("acgt", 3, 4), ("acgt", 4, 5), ("acgt", 5, 47), ("acgt", 6, 48),
        ("acgt", 7, 49), ("acgt", 8, 64), ("acgt", 9, 1000), ("acgt", 10, 12289), ("acgt", 11, 99999), ("acgt", 12, 1000003),
        ("iupac", 1, 7), ("iupac", 2, 50), ("iupac", 3, 1000), ("iupac", 5, 30011), ("iupac", 6, 1000003)]]
    doc = {"generator": "tests/golden/make_golden.py", "oracle_tables_sha256": sha(oracle.tables()), "cases": cases}
    covid = os.path.join(REFERENCE, "benchmarks", "covid.txt")
    if os.path.exists(covid):
        with open(covid, "rb") as f:
            genome = f.read().replace(b"\n", b"").replace(b"\r", b"")
        frames = oracle.py_all_frames(1, genome)
        doc["covid"] = {
            "source": "benchmarks/covid.txt of the reference, newlines stripped (benchmarks/bench.py:12-15)",
            "n": len(genome), "input_sha256": sha(genome),
            "translate_t1": sha(oracle.translate_bytes(1, genome)), "translate_t2": sha(oracle.translate_bytes(2, genome)),
            "translate_t22": sha(oracle.translate_bytes(22, genome)), "revcomp": sha(oracle.revcomp_bytes(genome)),
            "all_frames_lens": [len(f) for f in frames], "all_frames_t1": sha(b"\n".join(frames)),
            "translate_t1_head": oracle.translate_bytes(1, genome)[:40].decode(),
        }
    else:  # keep the section generated where the reference was present
        try:
            with open(os.path.join(os.path.dirname(__file__), "hotpath_golden.json")) as f:
                doc["covid"] = json.load(f).get("covid")
        except Exception:
            pass
    with open(os.path.join(os.path.dirname(__file__), "hotpath_golden.json"), "w") as f:
        json.dump(doc, f, indent=1)
        f.write("\n")
    print("wrote", len(cases), "cases", "+ covid" if doc.get("covid") else "")


if __name__ == "__main__":
    main()
