"""Small shared helpers (reference ``internlm/utils/common.py``): CLI parsing, device moves, batch skipping, the
Megatron FLOPs formula used for the TFLOPS metric, scheduler-hook ABC."""
from __future__ import annotations

import bisect
import inspect
import os
import random
from abc import ABC, abstractmethod
from contextlib import contextmanager
from datetime import datetime
from typing import Union

import numpy as np
import torch

from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)


def parse_args():
    from internevo_b200.initialize.launch import get_default_parser

    return get_default_parser().parse_args()


def get_master_node():
    import subprocess

    if os.getenv("SLURM_JOB_ID") is None:
        raise RuntimeError("get_master_node can only used in Slurm launch!")
    result = subprocess.check_output('scontrol show hostnames "$SLURM_JOB_NODELIST" | head -n 1', shell=True)
    return result.decode("utf8").strip()


def get_current_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device(f"cuda:{torch.cuda.current_device()}")
    return torch.device("cpu")


def move_norm_to_cuda(norm: Union[float, torch.Tensor]):
    if torch.is_tensor(norm) and norm.device.type != "cuda" and torch.cuda.is_available():
        norm = norm.to(get_current_device())
    return norm


def _move_tensor(element):
    if not torch.is_tensor(element):
        if isinstance(element, (list, tuple)):
            return type(element)(_move_tensor(e) for e in element)
        return element
    dev = get_current_device()
    if element.device != dev:
        element = element.to(dev, non_blocking=True)
    return element.detach()


def move_to_device(data):
    if torch.is_tensor(data):
        return _move_tensor(data)
    if isinstance(data, (list, tuple)):
        return type(data)(move_to_device(d) for d in data)
    if isinstance(data, dict):
        return {k: move_to_device(v) for k, v in data.items()}
    return data


def get_tensor_norm(norm: Union[float, torch.Tensor], move_to_cuda) -> torch.Tensor:
    if isinstance(norm, float):
        norm = torch.Tensor([norm])
    if move_to_cuda:
        norm = norm.to(get_current_device())
    return norm


def get_batch_size(data):
    if isinstance(data, torch.Tensor):
        return data.size(0)
    if isinstance(data, (list, tuple)):
        return get_batch_size(data[0])
    if isinstance(data, dict):
        return get_batch_size(next(iter(data.values())))
    raise TypeError(type(data))


def check_data_is_packed(data):
    if isinstance(data, torch.Tensor):
        return False
    if isinstance(data, (list, tuple)):
        return check_data_is_packed(data[0])
    if isinstance(data, dict):
        return "indexes" in data
    return False


def filter_kwargs(func, kwargs):
    sig = inspect.signature(func)
    return {k: v for k, v in kwargs.items() if k in sig.parameters}


def launch_time():
    global _CURRENT_TIME
    if _CURRENT_TIME is None:
        _CURRENT_TIME = datetime.now().strftime("%m-%d-%H:%M:%S")
    return _CURRENT_TIME


_CURRENT_TIME = None


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


@contextmanager
def conditional_context(context_manager, enable=True):
    if enable:
        with context_manager:
            yield
    else:
        yield


class BatchSkipper:
    """``data.skip_batches = "2-5,7"`` → skip batches 2,3,4,5,7 (reference ``utils/common.py:165-188``)."""

    def __init__(self, skip_batches):
        if skip_batches == "" or skip_batches is None:
            self.ranges = []
        else:
            ranges = []
            for part in str(skip_batches).split(","):
                if "-" in part:
                    start, end = map(int, part.split("-"))
                else:
                    start = end = int(part)
                if ranges and ranges[-1][1] == start - 1:
                    ranges[-1] = (ranges[-1][0], end)
                else:
                    ranges.append((start, end))
            self.ranges = sorted(ranges)

    def __call__(self, batch_count):
        idx = bisect.bisect_right(self.ranges, (batch_count, float("inf"))) - 1
        return idx >= 0 and self.ranges[idx][0] <= batch_count <= self.ranges[idx][1]


class SingletonMeta(type):
    _instances = {}

    def __call__(cls, *args, **kwargs):
        if cls not in cls._instances:
            cls._instances[cls] = super().__call__(*args, **kwargs)
        return cls._instances[cls]


def get_megatron_flops(elapsed_time_per_iter, checkpoint=False, seq_len=2048, hidden_size=12, num_layers=32,
                       vocab_size=12, global_batch_size=4, global_world_size=1, mlp_ratio=4, use_swiglu=True):
    """TFLOPS per GPU, Megatron formula — kept verbatim-equivalent to the reference (``utils/common.py:208-238``) so
    numbers are comparable: ignores GQA and causal masking; x4 with activation checkpointing, else x3."""
    factor = 4 if checkpoint else 3
    if use_swiglu:
        mlp_ratio = mlp_ratio * 3 / 2
    flops_per_iteration = (
        factor
        * ((8 + mlp_ratio * 4) * global_batch_size * seq_len * hidden_size**2
           + 4 * global_batch_size * seq_len**2 * hidden_size)
    ) * num_layers + 6 * global_batch_size * seq_len * hidden_size * vocab_size
    return flops_per_iteration / (elapsed_time_per_iter * global_world_size * (10**12))


def enable_pytorch_expandable_segments():
    if torch.cuda.is_available():
        setting = "expandable_segments:True"
        if os.getenv("PYTORCH_CUDA_ALLOC_CONF"):
            setting = os.getenv("PYTORCH_CUDA_ALLOC_CONF") + "," + setting
        try:
            torch.cuda.memory._set_allocator_settings(setting)
        except Exception as e:  # pragma: no cover
            logger.warning(f"could not enable expandable segments: {e}")


class DummyProfile:
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, a, b, c):
        pass

    def step(self):
        pass


class SchedulerHook(ABC):
    """Callbacks invoked by the schedulers around forward / loss / backward (reference ``utils/common.py:269-300``)."""

    @abstractmethod
    def before_forward(self, scheduler, inputs) -> None:
        """before the forward pass"""

    @abstractmethod
    def after_forward(self, scheduler, outputs) -> None:
        """after the forward pass"""

    @abstractmethod
    def before_criterion(self, scheduler, outputs, label) -> None:
        """before loss computation"""

    @abstractmethod
    def after_criterion(self, scheduler, loss) -> None:
        """after loss computation"""

    @abstractmethod
    def before_backward(self, scheduler, outputs, outputs_grad) -> None:
        """before backward"""

    @abstractmethod
    def after_backward(self, scheduler, inputs_grad) -> None:
        """after backward"""

    @abstractmethod
    def post_helper_func(self, scheduler, outputs, label) -> None:
        """metrics etc."""


from internevo_b200.core.context.config import read_base  # noqa: E402,F401  (config files import it from here too)
