"""ctypes face of ``csrc/dataio.cpp`` (``internevo_b200/_dataio.so``): corpus scan and token-line parser for the tokenised-corpus
format.  The library is optional - without it (or for lines that are not the plain ``{"tokens": [...]}`` form) the callers use the
JSON decoder - and is built together with the CUDA extension (``csrc/build.py::build_dataio``; a missing library is compiled once
on first use when a host compiler is around)."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_dataio.so")
_lib = None
_tried = False


def lib():
    """The loaded library or ``None``."""
    global _lib, _tried
    if _lib is not None or _tried:
        return _lib
    _tried = True
    if os.environ.get("B200_NATIVE_DATAIO", "1") == "0":
        return None
    if not os.path.exists(_SO):
        try:
            from internevo_b200.csrc.build import build_dataio

            build_dataio()
        except Exception:       # no compiler / read-only tree: the Python path is always there
            return None
    try:
        handle = ctypes.CDLL(_SO)
    except OSError:
        return None
    i64p = ctypes.POINTER(ctypes.c_int64)
    handle.b200_parse_tokens.argtypes = [ctypes.c_char_p, ctypes.c_int64, i64p, ctypes.c_int64]
    handle.b200_parse_tokens.restype = ctypes.c_int64
    handle.b200_scan_jsonl.argtypes = [ctypes.c_char_p, i64p, ctypes.c_int64]
    handle.b200_scan_jsonl.restype = ctypes.c_int64
    _lib = handle
    return _lib


def parse_tokens(line: bytes) -> Optional[np.ndarray]:
    """int64 tokens of one JSON line, or ``None`` when the line is not the plain form (the caller decodes it with ``json``)."""
    h = lib()
    if h is None:
        return None
    cap = len(line) // 2 + 1                         # every token takes at least one digit and one separator
    out = np.empty(cap, dtype=np.int64)
    n = h.b200_parse_tokens(line, len(line), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), cap)
    return out[:n] if n >= 0 else None


def scan_jsonl(path: str) -> Optional[np.ndarray]:
    """``[lines, 2]`` int64 table of (byte offset, token count), or ``None`` (no library / a line in another form)."""
    h = lib()
    if h is None:
        return None
    n = h.b200_scan_jsonl(path.encode(), None, 0)
    if n < 0:
        return None
    out = np.empty((n, 2), dtype=np.int64)
    got = h.b200_scan_jsonl(path.encode(), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), n)
    return out if got == n else None
