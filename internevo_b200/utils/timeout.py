"""Wall-clock guards around long host-side phases (reference behaviour: ``internlm/utils/timeout.py:13-113``).

``@llm_timeout(func_name="save_checkpoint")`` arms a per-phase limit while the wrapped function runs and turns a stall into a
``TimeoutError`` that names host and rank.  Guards are inert unless ``INTERNLM_ENABLE_TIMEOUT`` is set (the switch the
reference uses), so normal runs never install a signal handler.  The limit table keeps the reference's phase names -
schedulers, the checkpoint manager and the data pipeline refer to them - with limits as ONE immutable mapping queried through
``limit_for`` instead of a dict that is rewritten at import time.

Implementation: ``signal.setitimer(ITIMER_REAL)`` with save / restore of the previous handler and timer, so guards nest (an
inner phase re-arms the outer one's remaining time on exit); guards requested off the main thread degrade to no-ops, because
Python delivers signals to the main thread only.
"""
from __future__ import annotations

import datetime
import os
import signal
import socket
import threading
import time
import traceback
from functools import wraps
from types import MappingProxyType

from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)

ENABLE_TIMEOUT = os.getenv("INTERNLM_ENABLE_TIMEOUT", None)

_PHASE_LIMITS_S = MappingProxyType({
    # start-up
    "initialize_distributed_env": 240, "initialize_model": 60, "initialize_optimizer": 60,
    "get_train_data_loader": 600, "get_validation_data_loader": 60,
    # steady state
    "load_new_batch": 20, "nopp_forward_backward_step": 360, "nointerleaved_forward_backward_step": 600,
    "interleaved_forward_backward_step": 600, "optim_step": 60, "record_current_batch_training_metrics": 20,
    # checkpointing
    "save_checkpoint": 1200,
})


def limit_for(phase: str, default: int = 0) -> int:
    """Seconds allowed for ``phase``; 0 (no guard) while the feature is switched off."""
    if ENABLE_TIMEOUT is None:
        return 0
    return int(_PHASE_LIMITS_S.get(phase, default))


# name kept for code written against the reference: a read-only view with the limits that are in force
timeout_threshold_dict = MappingProxyType({k: limit_for(k) for k in _PHASE_LIMITS_S})

if ENABLE_TIMEOUT is not None:
    # a hung collective must raise inside the process instead of blocking forever
    os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "1")
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
    LLM_NCCL_TIMEOUT = datetime.timedelta(seconds=int(os.getenv("NCCL_TIMEOUT", "60")))
else:
    LLM_NCCL_TIMEOUT = datetime.timedelta(seconds=1800)


class Timeout:
    """``with Timeout(seconds, "message"):`` raises ``TimeoutError(message)`` in the main thread when the block overruns."""

    def __init__(self, seconds=1, error_message="Timeout"):
        self.seconds = float(seconds)
        self.error_message = error_message
        self._armed = False

    def handle_timeout(self, signum, frame):
        raise TimeoutError(self.error_message)

    timeout_handler = handle_timeout      # the reference's name for the same SIGALRM handler (``utils/timeout.py:27``)

    def __enter__(self):
        if self.seconds > 0 and threading.current_thread() is threading.main_thread():
            self._prev_handler = signal.signal(signal.SIGALRM, self.handle_timeout)
            self._prev_left, _ = signal.setitimer(signal.ITIMER_REAL, self.seconds)
            self._t0 = time.monotonic()
            self._armed = True
        return self

    def __exit__(self, exc_type, exc, tb):
        if self._armed:
            signal.setitimer(signal.ITIMER_REAL, 0)
            signal.signal(signal.SIGALRM, self._prev_handler)
            if self._prev_left > 0:   # an enclosing guard was running: give it back what is left of its budget
                signal.setitimer(signal.ITIMER_REAL, max(0.001, self._prev_left - (time.monotonic() - self._t0)))
            self._armed = False


def try_get_gpc_rank():
    try:
        from internevo_b200.core.context import global_context as gpc

        rank = gpc.get_global_rank()
    except Exception:  # pragma: no cover - before the context exists
        rank = "unknown"
    return f"host-{socket.gethostname()}-rank-{rank}"


def llm_timeout(seconds=0, func_name=None):
    """Decorator form.  The limit is the table entry of ``func_name`` (default: the function's own name) when the feature is
    on, else ``seconds`` as given (0 = unguarded)."""

    def decorator(func):
        phase = func_name or func.__name__

        @wraps(func)
        def guarded(*args, **kwargs):
            limit = limit_for(phase, seconds) if ENABLE_TIMEOUT is not None else seconds
            if not limit:
                return func(*args, **kwargs)
            try:
                with Timeout(limit, f"{phase} did not finish within {limit} s"):
                    return func(*args, **kwargs)
            except TimeoutError:
                logger.error(f"TimeoutError at {try_get_gpc_rank()}: {phase}\n{traceback.format_exc()}")
                raise

        return guarded

    return decorator
