"""Loader for the in-tree sm_100a extension (``internevo_b200/_C.so`` -> ``torch.ops.b200``).

On a machine with a GPU the extension is mandatory: ops raise instead of silently falling back to eager PyTorch.
On CPU-only machines (unit tests of the host-side logic) ``available()`` is False and the ops use their PyTorch
reference implementations.
"""
from __future__ import annotations

import os
import threading

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO_PATH = os.path.join(_PKG, "_C.so")
_lock = threading.Lock()
_loaded = False
_load_error: Exception | None = None


def load(build_if_missing: bool = True) -> bool:
    """Load (building first if necessary) the native library. Returns True when ``torch.ops.b200`` is usable."""
    global _loaded, _load_error
    if _loaded:
        return True
    with _lock:
        if _loaded:
            return True
        try:
            if not os.path.exists(SO_PATH) and build_if_missing:
                from internevo_b200.csrc.build import build

                build()
            torch.ops.load_library(SO_PATH)
            _loaded = True
        except Exception as e:  # pragma: no cover - depends on toolchain
            _load_error = e
            _loaded = False
    return _loaded


def available() -> bool:
    """True when CUDA kernels can actually run: a GPU is present and the extension is loaded."""
    if not torch.cuda.is_available():
        return False
    if not load():
        raise RuntimeError(
            f"internevo_b200 native extension failed to load on a CUDA machine ({_load_error}); "
            "run `python -m internevo_b200.csrc.build`"
        )
    return True


def use_native(*tensors: torch.Tensor) -> bool:
    """Whether the hand-written kernels should handle these tensors (all CUDA) -- raises if the lib is missing."""
    if not tensors or not all(t.is_cuda for t in tensors if t is not None):
        return False
    return available()


def ops():
    load()
    return torch.ops.b200
