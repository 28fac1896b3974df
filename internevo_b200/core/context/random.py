"""Per-parallel-mode RNG streams (reference ``internlm/core/context/random.py:16-133``): each ``ParallelMode`` owns a
device RNG state; ``set_mode`` swaps the live generator state so that e.g. dropout is identical inside a DP group but
different across TP ranks.  Works on CUDA and on CPU (gloo plumbing tests)."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict

import torch

from .process_groups import ParallelMode


def _get_state():
    return torch.cuda.get_rng_state() if torch.cuda.is_available() else torch.get_rng_state()


def _set_state(state):
    if torch.cuda.is_available():
        torch.cuda.set_rng_state(state)
    else:
        torch.set_rng_state(state)


def _manual_seed(seed):
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    else:
        torch.manual_seed(seed)


class SeedManager:
    def __init__(self):
        self._current_mode = None
        self._seeds: Dict[ParallelMode, int] = {}
        self._seed_states: Dict[ParallelMode, torch.Tensor] = {}

    @property
    def current_mode(self):
        return self._current_mode

    @property
    def seeds(self):
        return self._seeds

    @property
    def seed_states(self):
        return self._seed_states

    def set_state(self, mode: ParallelMode, state: torch.Tensor):
        assert mode in self._seed_states, f"{mode} not found in seed manager"
        self._seed_states[mode] = state

    def set_mode(self, mode: ParallelMode):
        if self._current_mode is not None:
            self._seed_states[self._current_mode] = _get_state()
        self._current_mode = mode
        _set_state(self._seed_states[mode])

    def add_seed(self, mode: ParallelMode, seed: int, overwrite: bool = False):
        assert isinstance(mode, ParallelMode)
        if not overwrite:
            assert mode not in self._seed_states, f"seed for {mode} already exists"
        current = _get_state()
        _manual_seed(seed)
        self._seed_states[mode] = _get_state()
        self._seeds[mode] = seed
        _set_state(current)

    def reset(self):
        self._current_mode = None
        self._seeds = {}
        self._seed_states = {}


_SEED_MANAGER = SeedManager()


def get_seeds():
    return _SEED_MANAGER.seeds


def get_states(copy=False):
    states = _SEED_MANAGER.seed_states
    return {k: v.clone() for k, v in states.items()} if copy else states


def get_current_mode():
    return _SEED_MANAGER.current_mode


def add_seed(mode, seed, overwrite=False):
    _SEED_MANAGER.add_seed(mode, seed, overwrite)


def set_mode(mode):
    _SEED_MANAGER.set_mode(mode)


def set_seed_states(mode, state):
    _SEED_MANAGER.set_state(mode, state)


def sync_states():
    set_seed_states(get_current_mode(), _get_state())


def reset_seeds():
    _SEED_MANAGER.reset()


@contextmanager
def seed(mode: ParallelMode):
    """Temporarily switch the RNG stream: ``with seed(ParallelMode.DATA): ...``"""
    current = _SEED_MANAGER.current_mode
    try:
        yield _SEED_MANAGER.set_mode(mode)
    finally:
        _SEED_MANAGER.set_mode(current)
