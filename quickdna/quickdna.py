"""`quickdna.quickdna`: the five functions of the reference's PyO3 module (/root/reference/src/python_api.rs:21-74)."""

from quickdna_b200.quickdna import (  # noqa: F401
    _check_table,
    _reverse_complement,
    _reverse_complement_strict,
    _translate,
    _translate_strict,
)

__all__ = ["_check_table", "_translate", "_translate_strict", "_reverse_complement", "_reverse_complement_strict"]
