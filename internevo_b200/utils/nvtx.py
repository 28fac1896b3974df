"""NVTX ranges for Nsight Systems / ``torch.profiler`` timelines (the reference's accelerator exposes
``range_push/pop`` (``internlm/accelerator/cuda_accelerator.py:270-282``) but nothing calls them).

Off by default (one boolean test per call site); ``B200_NVTX=1`` or ``--profiling`` turns the ranges on.  Ranges mark the
decoder layers, the scheduler's forward / backward of every micro-batch, the optimizer phases and the fused
compute + collective kernels, so a timeline shows which collective overlaps which layer.
"""
from __future__ import annotations

import os
from contextlib import contextmanager

import torch

_enabled = os.environ.get("B200_NVTX", "0") == "1"


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = bool(flag)


def enabled() -> bool:
    return _enabled and torch.cuda.is_available()


def range_push(name: str) -> None:
    if _enabled and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)


def range_pop() -> None:
    if _enabled and torch.cuda.is_available():
        torch.cuda.nvtx.range_pop()


@contextmanager
def nvtx_range(name: str):
    on = _enabled and torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()
