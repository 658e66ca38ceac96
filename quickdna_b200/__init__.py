"""quickdna_b200 -- B200-native drop-in for the sequence hot path of SecureDNA/quickdna.

The Python surface mirrors the reference wrapper (/root/reference/quickdna/__init__.py): the
same three bytes-backed classes and method names, so `from quickdna_b200 import DnaSequence`
replaces `from quickdna import DnaSequence`.  All arithmetic runs in hand-written sm_100a
kernels behind the C ABI of include/quickdna_b200.h; there is no CPU fallback.

Differences that are deliberate:
  * translate_self_frames / translate_all_frames cross the native boundary ONCE (the reference
    slices in Python and makes 3 / 7 native calls); results and errors are identical;
  * ValueErrors are `TranslationError` (a ValueError subclass) that also carry `.pos`.
"""

from __future__ import annotations

import ctypes as _C
import typing as ty

from . import _native as _N
from .quickdna import _reverse_complement, _reverse_complement_strict, _translate, _translate_strict

TranslationError = _N.TranslationError
SeqT = ty.TypeVar("SeqT", bound="BaseSequence")


def ensure_bytes(value: ty.Union[str, bytes]) -> bytes:
    """str -> ASCII bytes (UnicodeEncodeError otherwise); bytes pass through (quickdna/__init__.py:11-16)."""
    return value.encode("ascii", "strict") if isinstance(value, str) else value


class BaseSequence:
    """Bytes-backed sequence; performs no validation at construction (quickdna/__init__.py:19-101)."""

    __slots__ = ("_seq",)

    def __init__(self, seq: ty.Union[str, bytes]) -> None:
        self._seq = ensure_bytes(seq)

    @property
    def seq(self) -> bytes:
        return self._seq

    def __getitem__(self, key):
        if isinstance(key, slice):
            return type(self)(self._seq[key])
        if isinstance(key, ty.SupportsIndex):
            return self._seq[key]
        raise TypeError(f"can't index with {type(key)}")

    def __len__(self) -> int:
        return len(self._seq)

    def __iter__(self) -> ty.Iterator[str]:
        return map(chr, self._seq)

    def __add__(self: SeqT, other: ty.Union[str, bytes, SeqT]) -> SeqT:
        if isinstance(other, type(self)):
            tail = other.seq
        elif isinstance(other, (str, bytes)):
            tail = ensure_bytes(other)
        else:  # e.g. DnaSequence + ProteinSequence
            raise TypeError(f"can't concat {type(self)} to {type(other)}")
        return type(self)(self._seq + tail)

    def __mul__(self: SeqT, by: ty.SupportsIndex) -> SeqT:
        return type(self)(self._seq * by)

    def __repr__(self) -> str:
        # only the bytes that are shown get decoded (quickdna/__init__.py:70-76)
        shown = self._seq[:30].decode("ascii") + "..." if len(self._seq) > 30 else self._seq.decode("ascii")
        return f"{type(self).__name__}(seq={shown!r})"

    def __str__(self) -> str:
        return self._seq.decode("ascii")

    def __bytes__(self) -> bytes:
        return self._seq

    def __eq__(self, other) -> bool:
        if isinstance(other, type(self)):
            return self._seq == other._seq
        return NotImplemented

    def __hash__(self) -> int:
        return hash((type(self), self._seq))


class ProteinSequence(BaseSequence):
    """IUPAC amino-acid letters as ASCII bytes (quickdna/__init__.py:96-101)."""

    __slots__ = ()


class DnaSequence(BaseSequence):
    """A, T, C, G or IUPAC ambiguity codes as ASCII bytes (quickdna/__init__.py:104-201)."""

    __slots__ = ()

    def translate(self, table: int = 1, strict: bool = False) -> ProteinSequence:
        """quickdna/__init__.py:109-126 -> _translate / _translate_strict."""
        native = _translate_strict if strict else _translate
        return ProteinSequence(native(table, self._seq))

    def _frames(self, table: int, strict: bool, which: int) -> ty.List[ProteinSequence]:
        from .quickdna import _as_u8_table

        table = _as_u8_table(table)
        ctx = _N.default_context()
        n = len(self._seq)
        cap = sum(_N.lib().qd_frame_len(n, k) for k in range(3)) * (2 if which == _N.FRAMES_ALL else 1)
        out = _C.create_string_buffer(max(cap, 1))
        lens = (_C.c_size_t * 6)()
        n_frames = _C.c_int()
        err = _N.QdErr()
        rc = _N.lib().qd_translate_frames(ctx.handle, table, _N.ENC_ASCII, int(bool(strict)), which, self._seq, n, out,
                                          lens, _C.byref(n_frames), _C.byref(err))
        ctx.check(rc, err)
        frames, off, raw = [], 0, out.raw
        for ln in lens:
            if ln:
                frames.append(ProteinSequence(raw[off:off + ln]))
                off += ln
        assert len(frames) == n_frames.value
        return frames

    def translate_self_frames(self, table: int = 1, strict: bool = False) -> ty.List[ProteinSequence]:
        """Up to 3 reading frames of this strand (quickdna/__init__.py:128-161): 3 if len >= 5,
        2 if len == 4, 1 if len == 3, else none.  One native call instead of three."""
        return self._frames(table, strict, _N.FRAMES_SELF)

    def translate_all_frames(self, table: int = 1, strict: bool = False) -> ty.List[ProteinSequence]:
        """Up to 6 reading frames, forward then reverse complement (quickdna/__init__.py:163-185).
        One native call instead of seven; the reverse complement is never materialised."""
        if len(self._seq) < 3:
            # The reference still runs reverse_complement() here -- WITHOUT forwarding `strict`
            # (quickdna/__init__.py:182) -- so invalid bytes raise even though no frame exists.
            # It never looks at `table` on this path.
            _reverse_complement(self._seq)
            return []
        return self._frames(table, strict, _N.FRAMES_ALL)

    def reverse_complement(self, strict: bool = False) -> "DnaSequence":
        """quickdna/__init__.py:187-201 -> _reverse_complement / _reverse_complement_strict."""
        native = _reverse_complement_strict if strict else _reverse_complement
        return DnaSequence(native(self._seq))


__all__ = ["BaseSequence", "DnaSequence", "ProteinSequence", "TranslationError"]
