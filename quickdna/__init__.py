"""Import-name drop-in: `from quickdna import DnaSequence` binds the B200-native implementation.

The reference package is `quickdna` with its native module `quickdna.quickdna`
(/root/reference/quickdna/__init__.py:6); this shim gives quickdna_b200 the same import names so that callers --
and the reference's own tests/test_wrapper.py -- run unchanged.  It contains no logic.
"""

from quickdna_b200 import BaseSequence, DnaSequence, ProteinSequence, TranslationError, ensure_bytes  # noqa: F401

__all__ = ["BaseSequence", "DnaSequence", "ProteinSequence", "TranslationError", "ensure_bytes"]
