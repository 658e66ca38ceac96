"""Mirror of the reference's native extension module `quickdna.quickdna`
(/root/reference/src/python_api.rs:21-74): the same five functions, the same argument
meaning, the same ValueError texts -- computed by the sm_100a kernels behind the C ABI.
"""

from __future__ import annotations

import ctypes as C

from . import _native as N

__all__ = ["_check_table", "_translate", "_translate_strict", "_reverse_complement", "_reverse_complement_strict"]


def _table_error(table: int) -> N.TranslationError:
    return N.make_error(N.QdErr(N.QD_ERR_BAD_TABLE, table & 0xFF, 0, 0))


def _as_u8_table(table) -> int:
    table = int(table)
    if not 0 <= table <= 255:
        raise OverflowError("out of range integral type conversion attempted")  # PyO3's u8 extraction
    return table


def _check_table(table: int) -> None:
    """src/python_api.rs:21-25"""
    table = _as_u8_table(table)
    if N.lib().qd_table_slot(table) < 0:
        raise _table_error(table)


def _translate_impl(table: int, dna: bytes, strict: int) -> bytes:
    if not isinstance(dna, (bytes, bytearray)):
        raise TypeError("argument 'dna': expected bytes")
    table = _as_u8_table(table)
    ctx = N.default_context()
    n = len(dna)
    out = C.create_string_buffer(max(n // 3, 1))
    err = N.QdErr()
    rc = N.lib().qd_translate(ctx.handle, table, N.ENC_ASCII, strict, bytes(dna), n, out, C.byref(err))
    ctx.check(rc, err)
    return out.raw[: n // 3]


def _translate(table: int, dna: bytes) -> bytes:
    """src/python_api.rs:33-38 -- IUPAC ambiguity codes allowed."""
    return _translate_impl(table, dna, 0)


def _translate_strict(table: int, dna: bytes) -> bytes:
    """src/python_api.rs:46-51 -- only A, T, C, G (either case)."""
    return _translate_impl(table, dna, 1)


def _revcomp_impl(dna: bytes, strict: int) -> bytes:
    if not isinstance(dna, (bytes, bytearray)):
        raise TypeError("argument 'dna': expected bytes")
    ctx = N.default_context()
    n = len(dna)
    out = C.create_string_buffer(max(n, 1))
    err = N.QdErr()
    rc = N.lib().qd_revcomp(ctx.handle, N.ENC_ASCII, strict, bytes(dna), n, out, C.byref(err))
    ctx.check(rc, err)
    return out.raw[:n]


def _reverse_complement(dna: bytes) -> bytes:
    """src/python_api.rs:58-62"""
    return _revcomp_impl(dna, 0)


def _reverse_complement_strict(dna: bytes) -> bytes:
    """src/python_api.rs:70-74"""
    return _revcomp_impl(dna, 1)
