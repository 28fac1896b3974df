"""SIGALRM-based guards around long host functions (reference ``internlm/utils/timeout.py:13-113``).  Active only when
``INTERNLM_ENABLE_TIMEOUT`` is set; thresholds come from per-function env-free defaults below."""
from __future__ import annotations

import datetime
import os
import signal
import socket
import traceback
from functools import wraps

from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)


class Timeout:
    """``with Timeout(seconds, "msg"):`` raises ``TimeoutError`` from the alarm handler."""

    def __init__(self, seconds=1, error_message="Timeout"):
        self.seconds = int(seconds)
        self.error_message = error_message

    def handle_timeout(self, signum, frame):
        raise TimeoutError(self.error_message)

    def __enter__(self):
        if self.seconds > 0:
            signal.signal(signal.SIGALRM, self.handle_timeout)
            signal.alarm(self.seconds)
        return self

    def __exit__(self, exc_type, exc, tb):
        if self.seconds > 0:
            signal.alarm(0)


ENABLE_TIMEOUT = os.getenv("INTERNLM_ENABLE_TIMEOUT", None)

timeout_threshold_dict = {
    "initialize_distributed_env": 240,
    "nopp_forward_backward_step": 360,
    "initialize_model": 60,
    "initialize_optimizer": 60,
    "optim_step": 60,
    "get_train_data_loader": 600,
    "get_validation_data_loader": 60,
    "load_new_batch": 20,
    "record_current_batch_training_metrics": 20,
    "save_checkpoint": 1200,
    "interleaved_forward_backward_step": 600,
    "nointerleaved_forward_backward_step": 600,
}

if ENABLE_TIMEOUT is not None:
    os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "1")
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
    LLM_NCCL_TIMEOUT = datetime.timedelta(seconds=int(os.getenv("NCCL_TIMEOUT", str(60))))
else:
    timeout_threshold_dict = dict.fromkeys(timeout_threshold_dict.keys(), 0)
    LLM_NCCL_TIMEOUT = datetime.timedelta(seconds=1800)


def try_get_gpc_rank():
    try:
        from internevo_b200.core.context import global_context as gpc

        rank = gpc.get_global_rank()
    except Exception:  # pragma: no cover
        rank = "unknown"
    return f"host-{socket.gethostname()}-rank-{rank}"


def llm_timeout(seconds=0, func_name=None):
    """Decorator: abort ``func`` with ``TimeoutError`` after ``seconds`` (or the table entry for ``func_name``)."""

    def decorator(func):
        nonlocal func_name
        if func_name is None:
            func_name = func.__name__

        @wraps(func)
        def wrapper(*args, **kwargs):
            limit = timeout_threshold_dict.get(func_name, seconds) if ENABLE_TIMEOUT is not None or seconds == 0 else seconds
            if ENABLE_TIMEOUT is None and seconds == 0:
                limit = 0
            try:
                with Timeout(limit, f"{func_name} timed out after {limit}s"):
                    return func(*args, **kwargs)
            except TimeoutError as e:
                logger.error(f"TimeoutError at {try_get_gpc_rank()}: {func_name}\n{traceback.format_exc()}")
                raise e

        return wrapper

    return decorator
