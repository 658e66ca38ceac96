"""ctypes binding of libquickdna_b200.so (the C ABI in include/quickdna_b200.h).

There is no CPU fallback: if the CUDA library is missing, or no B200 is visible, every call
fails loudly.  Nothing here imports oracle/.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QUICKDNA_B200_LIB") or os.path.join(_HERE, "libquickdna_b200.so")  # the override is for A/B builds of the same ABI

QD_OK = 0
QD_ERR_NON_ASCII_BYTE = 1
QD_ERR_BAD_NUCLEOTIDE = 2
QD_ERR_UNEXPECTED_AMBIGUOUS = 3
QD_ERR_BAD_TABLE = 4
QD_ERR_BAD_MASK = 5
QD_ERR_INVALID_ARGUMENT = 90
QD_ERR_CUDA = 100

ENC_ASCII = 0
ENC_MASK = 1
FRAMES_ONE = 1
FRAMES_SELF = 3
CANON_KEY_PACKED, CANON_KEY_FP64 = 0, 1
FRAMES_ALL = 6


class QdErr(C.Structure):
    _fields_ = [("kind", C.c_int32), ("byte", C.c_uint8), ("pos", C.c_uint64), ("seq_index", C.c_uint64)]


class NativeLibraryMissing(RuntimeError):
    pass


class CudaError(RuntimeError):
    pass


class TranslationError(ValueError):
    """ValueError carrying the reference's message (src/python_api.rs:15-19) plus position info."""

    def __init__(self, message: str, kind: int, byte: int, pos: int, seq_index: int):
        super().__init__(message)
        self.kind = kind
        self.byte = byte
        self.pos = pos
        self.seq_index = seq_index


_vp = C.c_void_p
_sz = C.c_size_t
_i = C.c_int

# name -> (restype, argtypes): every symbol include/quickdna_b200.h declares
SIGNATURES = {
    "qd_abi_version": (_i, []),
    "qd_ctx_create": (_i, [C.POINTER(_vp), _i]),
    "qd_ctx_destroy": (None, [_vp]),
    "qd_ctx_device": (_i, [_vp]),
    "qd_ctx_stream": (_vp, [_vp]),
    "qd_ctx_sm_count": (_i, [_vp]),
    "qd_last_cuda_error": (C.c_char_p, [_vp]),
    "qd_ctx_set_chunk_bytes": (_i, [_vp, _sz]),
    "qd_ctx_set_stage_threads": (_i, [_vp, _i]),
    "qd_host_alloc": (_vp, [_sz]),
    "qd_host_free": (None, [_vp]),
    "qd_table_slot": (_i, [_i]),
    "qd_table_data": (_vp, [_i]),
    "qd_err_message": (_i, [C.POINTER(QdErr), C.c_char_p, _sz]),
    "qd_frame_len": (_sz, [_sz, _i]),
    "qd_runs": (_sz, [_sz]),
    "qd_slots_bytes": (_sz, [_sz, _i]),
    "qd_slot_frame_offset": (_sz, [_sz, _i, _i]),
    "qd_translate": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp, C.POINTER(QdErr)]),
    "qd_translate_frames": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp, C.POINTER(_sz), C.POINTER(_i), C.POINTER(QdErr)]),
    "qd_revcomp": (_i, [_vp, _i, _i, _vp, _sz, _vp, C.POINTER(QdErr)]),
    "qd_batch_out_bytes": (_sz, [_vp, _sz, _i]),
    "qd_translate_frames_batch": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp, _vp, C.POINTER(QdErr)]),
    "qd_translate_frames_batch_uniform": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _sz, _vp, C.POINTER(QdErr)]),
    "qd_canonical_windows": (_i, [_vp, _i, _i, _vp, _sz, _sz, _vp, C.POINTER(QdErr)]),
    "qd_expansion_windows": (_i, [_vp, _i, _vp, _sz, _sz, C.c_uint64, _vp, _vp, _vp, C.POINTER(QdErr)]),
    "qd_windows": (_i, [_vp, _vp, _sz, _sz, _vp, C.POINTER(_sz)]),
    "qd_windows_all": (_i, [_vp, _i, _vp, _sz, _sz, _sz, _vp, _vp, _vp, C.POINTER(_sz), C.POINTER(QdErr)]),
    "qd_dev_error_reset": (_i, [_vp, _vp]),
    "qd_dev_error_fetch": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _sz, C.POINTER(QdErr)]),
    "qd_revcomp_dev": (_i, [_vp, _i, _i, _vp, _sz, _vp, _vp]),
    "qd_translate_frames_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "qd_translate_frames_batch_uniform_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _sz, _vp, _vp]),
    "qd_translate_frames_batch_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _sz, _sz, _vp, _vp]),
    "qd_canonical_windows_dev": (_i, [_vp, _i, _i, _vp, _sz, _sz, _vp, _vp]),
    "qd_windows_dev": (_i, [_vp, _vp, _sz, _sz, _vp, _vp]),
    "qd_launch_count": (C.c_uint64, [_vp]),
    "qd_ctx_set_big_tile_min_runs": (C.c_int, [_vp, C.c_uint64]),
    "qd_canonical": (_i, [_vp, _i, _i, _vp, _sz, _vp, C.POINTER(QdErr)]),
    "qd_canonical_dev": (_i, [_vp, _i, _i, _vp, _sz, _vp, _vp]),
    "qd_canonical_windows_batch_dev": (_i, [_vp, _i, _i, _vp, _sz, _sz, _sz, _vp, _vp]),
    "qd_canonical_key_bytes": (_sz, [_i, _sz]),
    "qd_canonical_fingerprint": (C.c_uint64, [_vp, _sz]),
    "qd_canonical_window_keys": (_i, [_vp, _i, _i, _i, _vp, _sz, _sz, _vp, C.POINTER(QdErr)]),
    "qd_canonical_window_keys_dev": (_i, [_vp, _i, _i, _i, _vp, _sz, _sz, _sz, _vp, _vp]),
    "qd_expansion_windows_dev": (_i, [_vp, _i, _vp, _sz, _sz, C.c_uint64, _vp, _vp, _vp, C.c_uint64, _vp]),
    "qd_fasta_scan": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp, C.POINTER(_sz), _vp, _sz, C.POINTER(_sz), C.POINTER(QdErr)]),
    "qd_fasta_scan_dev": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp, C.POINTER(_sz), _vp, _sz, C.POINTER(_sz), _vp, C.POINTER(QdErr), _vp]),
    "qd_fasta_translate_frames": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _sz, _vp, _sz, C.POINTER(_sz), _vp, _sz, _vp, C.POINTER(_sz), C.POINTER(QdErr)]),
    "qd_windows_batch_dev": (_i, [_vp, _vp, _sz, _sz, _sz, _i, _vp, _vp]),
    "qd_windows_all_batch_counts": (_sz, [_sz, _sz, _sz, C.POINTER(_sz)]),
    "qd_windows_all_batch_dev": (_i, [_vp, _i, _vp, _sz, _sz, _sz, _sz, _vp, _vp, _vp]),
    "qd_parse": (_i, [_vp, _i, _vp, _sz, _vp, C.POINTER(_sz), C.POINTER(QdErr)]),
    "qd_parse_dev": (_i, [_vp, _i, _vp, _sz, _vp, _vp, _vp]),
    "qd_translate_frames_batch_multi": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _sz, _sz, _vp, _vp, C.POINTER(QdErr)]),
    "qd_translate_frames_multi": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _sz, _vp, C.POINTER(_sz), C.POINTER(_i), C.POINTER(QdErr)]),
    "qd_revcomp_multi": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp, C.POINTER(QdErr)]),
    "qd_checksum64_dev": (_i, [_vp, _vp, _sz, C.c_uint64, _vp, _vp]),
    "qd_checksum64_weight": (C.c_uint64, [C.c_uint64]),
    "qd_chunk_count": (_sz, [_sz, _sz]),
    "qd_canonical_chunk_bytes": (_sz, [_sz, _sz, _sz]),
    "qd_windows_all_chunk_bytes": (_sz, [_sz, _sz, _sz, _sz, C.POINTER(_sz)]),
    "qd_expansion_chunk_bytes": (_sz, [_sz, _sz, C.c_uint64]),
    "qd_canonical_windows_chunked_dev": (_i, [_vp, _i, _i, _vp, _sz, _sz, _sz, _vp, _sz, _sz, _vp, _vp, _vp, _vp]),
    "qd_windows_all_chunked_dev": (_i, [_vp, _i, _vp, _sz, _sz, _sz, _sz, _vp, _sz, _sz, _vp, _vp, _vp, _vp]),
    "qd_expansion_windows_chunked_dev": (_i, [_vp, _i, _vp, _sz, _sz, C.c_uint64, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
}

# qd_chunk_fn: (user, chunk, first_unit, n_units, slot_dev, slot_bytes, n_rows_dev, stream)
CHUNK_FN = C.CFUNCTYPE(None, _vp, _sz, _sz, _sz, _vp, _sz, _vp, _vp)

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load the C-ABI library (no device needed for this step)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeLibraryMissing(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  quickdna_b200 has no CPU fallback."
                )
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)  # AttributeError here = header/library mismatch
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


class Context:
    """One qd_ctx (one device, one stream, device-resident LUTs, grow-only scratch)."""

    def __init__(self, device: int = 0):
        self._h = _vp()
        rc = lib().qd_ctx_create(C.byref(self._h), int(device))
        if rc != QD_OK:
            raise CudaError(
                f"qd_ctx_create(device={device}) failed with status {rc}: no usable CUDA device. "
                "quickdna_b200 runs only on the GPU (no CPU fallback)."
            )
        self.device = device

    @property
    def handle(self) -> _vp:
        return self._h

    def close(self):
        if self._h:
            lib().qd_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, err: Optional[QdErr] = None, table: Optional[int] = None):
        if rc == QD_OK:
            return
        if rc == QD_ERR_BAD_TABLE and err is None and table is not None:
            err = QdErr(rc, int(table) & 0xFF, 0, 0)  # device-pointer calls carry no qd_err: the id is the payload
        if rc == QD_ERR_CUDA:
            raise CudaError(lib().qd_last_cuda_error(self._h).decode())
        if rc == QD_ERR_INVALID_ARGUMENT:
            msg = lib().qd_last_cuda_error(self._h).decode()
            raise ValueError("invalid argument" + (": " + msg if msg else ""))
        raise make_error(err if err is not None else QdErr(rc, 0, 0, 0))

    @property
    def launch_count(self) -> int:
        return int(lib().qd_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(lib().qd_ctx_stream(self._h) or 0)


def make_error(err: QdErr) -> TranslationError:
    buf = C.create_string_buffer(128)
    lib().qd_err_message(C.byref(err), buf, 128)
    return TranslationError(buf.value.decode("ascii"), err.kind, err.byte, err.pos, err.seq_index)


_contexts = {}
_ctx_lock = threading.Lock()


def context_array(ctxs):
    """qd_ctx*[] for the *_multi entry points."""
    arr = (_vp * len(ctxs))(*[c.handle for c in ctxs])
    return arr


def device_contexts(devices=None):
    """One default context per device (all visible devices when None)."""
    if devices is None:
        import torch

        devices = range(torch.cuda.device_count())
    return [default_context(int(d)) for d in devices]


def default_context(device: int = 0) -> Context:
    with _ctx_lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _contexts[device] = Context(device)
        return ctx
