"""The oracle's FASTA scanner (oracle/qd_oracle.c: qdo_fasta_scan) against the reference's own parser tests
(src/fasta.rs:596-1396), restated with DNA content in tests/vectors.py."""

import pytest

import oracle
from tests import vectors as V


@pytest.mark.parametrize("text,settings,want", V.FASTA)
def test_fasta_vectors(text, settings, want):
    for concat, allow in settings:
        for strict in (False, True):
            if strict and any(ch in "BDHKMNRSVWYbdhkmnrsvwy" for _, content, _ in want for ch in content):
                continue
            got = oracle.fasta_scan(text, strict=strict, concatenate_headers=concat, allow_preceding_comment=allow)
            assert [(h, oracle.ascii_of(m).decode(), lr) for h, m, lr in got] == want, (concat, allow, strict)


@pytest.mark.parametrize("text,strict,line,kind,byte,message", V.FASTA_ERRORS)
def test_fasta_errors(text, strict, line, kind, byte, message):
    with pytest.raises(oracle.FastaError) as ei:
        oracle.fasta_scan(text, strict=strict)
    assert (ei.value.line, ei.value.kind, ei.value.byte) == (line, kind, byte)
    assert ei.value.located() == message
    # the same text parses when the offending content precedes every header and preceding comments are allowed
    if not text.startswith(">"):
        oracle.fasta_scan(text, strict=strict, allow_preceding_comment=True)
