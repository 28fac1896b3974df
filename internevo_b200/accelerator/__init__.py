"""Device facade with the reference's ``internlm.accelerator`` call surface (``get_accelerator()``, ``AcceleratorType``).

The reference dispatches between CUDA and Ascend NPU back-ends here (``internlm/accelerator/{abstract,cuda,npu}_accelerator.py``).
This framework targets one device family (B200, sm_100a), so there is exactly ONE implementation and no dispatch: the class
below forwards to ``torch.cuda`` and degrades to harmless CPU answers when no GPU is visible (unit tests, config checks).
User scripts written against ``internlm.accelerator.get_accelerator()`` keep working.
"""
from __future__ import annotations

import enum
import functools

import torch


class AcceleratorType(enum.Enum):
    GPU = 1
    NPU = 2    # kept for config compatibility only; never returned
    CPU = 3
    OTHER = 4


class B200Accelerator:
    """``torch.cuda`` behind the reference's accelerator method names."""

    def __init__(self) -> None:
        self._has_gpu = torch.cuda.is_available()
        self._name = "cuda" if self._has_gpu else "cpu"
        self._communication_backend_name = "nccl" if self._has_gpu else "gloo"
        self.amp = None

    # ---- identity
    def get_backend_name(self):
        return self._name

    def get_accelerator_backend(self):
        return AcceleratorType.GPU if self._has_gpu else AcceleratorType.CPU

    def communication_backend_name(self):
        return self._communication_backend_name

    def is_available(self):
        return self._has_gpu

    # ---- devices
    def device_name(self, device_index=None):
        if not self._has_gpu:
            return "cpu"
        return "cuda" if device_index is None else f"cuda:{device_index}"

    def set_device(self, device_index):
        if self._has_gpu:
            torch.cuda.set_device(device_index)

    def get_device_id(self):
        return torch.cuda.current_device() if self._has_gpu else 0

    def current_device_name(self):
        return f"cuda:{torch.cuda.current_device()}" if self._has_gpu else "cpu"

    def device_count(self):
        return torch.cuda.device_count() if self._has_gpu else 0

    def synchronize(self, device_index=None):
        if self._has_gpu:
            torch.cuda.synchronize(device_index)

    def total_memory(self, device_index=None):
        return torch.cuda.get_device_properties(device_index or 0).total_memory if self._has_gpu else 0

    # ---- RNG
    def random(self):
        return torch.random

    def set_rng_state(self, new_state, device_index=None):
        if not self._has_gpu:
            return torch.set_rng_state(new_state)
        return torch.cuda.set_rng_state(new_state) if device_index is None else torch.cuda.set_rng_state(new_state, device_index)

    def get_rng_state(self, device_index=None):
        if not self._has_gpu:
            return torch.get_rng_state()
        return torch.cuda.get_rng_state() if device_index is None else torch.cuda.get_rng_state(device_index)

    def manual_seed(self, seed):
        return torch.cuda.manual_seed(seed) if self._has_gpu else torch.manual_seed(seed)

    def manual_seed_all(self, seed):
        return torch.cuda.manual_seed_all(seed) if self._has_gpu else torch.manual_seed(seed)

    def initial_seed(self):
        return torch.cuda.initial_seed() if self._has_gpu else torch.initial_seed()

    def default_generator(self, device_index):
        return torch.cuda.default_generators[device_index] if self._has_gpu else torch.default_generator

    # ---- streams / events
    @property
    def Stream(self):
        return torch.cuda.Stream

    def stream(self, _stream):
        return torch.cuda.stream(_stream)

    def current_stream(self, device_index=None):
        return torch.cuda.current_stream(device_index)

    def default_stream(self, device_index=None):
        return torch.cuda.default_stream(device_index)

    @property
    def Event(self):
        return torch.cuda.Event

    # ---- memory
    def empty_cache(self):
        if self._has_gpu:
            torch.cuda.empty_cache()

    def _mem(self, fn, device_index=None):
        return getattr(torch.cuda, fn)(device_index) if self._has_gpu else 0

    def memory_allocated(self, device_index=None):
        return self._mem("memory_allocated", device_index)

    def max_memory_allocated(self, device_index=None):
        return self._mem("max_memory_allocated", device_index)

    def reset_max_memory_allocated(self, device_index=None):
        return self.reset_peak_memory_stats(device_index)

    def memory_cached(self, device_index=None):
        return self._mem("memory_reserved", device_index)

    def max_memory_cached(self, device_index=None):
        return self._mem("max_memory_reserved", device_index)

    def reset_max_memory_cached(self, device_index=None):
        return self.reset_peak_memory_stats(device_index)

    def memory_stats(self, device_index=None):
        return torch.cuda.memory_stats(device_index) if self._has_gpu else {}

    def reset_peak_memory_stats(self, device_index=None):
        if self._has_gpu:
            torch.cuda.reset_peak_memory_stats(device_index)

    def memory_reserved(self, device_index=None):
        return self._mem("memory_reserved", device_index)

    def max_memory_reserved(self, device_index=None):
        return self._mem("max_memory_reserved", device_index)

    # ---- dtypes / amp
    def is_bf16_supported(self):
        return True

    def is_fp16_supported(self):
        return True

    def get_amp(self):
        return torch.amp

    def set_allow_tf32(self, enable: bool):
        torch.backends.cudnn.allow_tf32 = enable
        torch.backends.cuda.matmul.allow_tf32 = enable

    def return_custom_bwd(self):
        return functools.partial(torch.amp.custom_bwd, device_type=self._name)

    def return_custom_fwd(self):
        return functools.partial(torch.amp.custom_fwd, device_type=self._name)

    # ---- profiling ranges (see utils/nvtx.py for the toggled ranges the framework itself emits)
    def range_push(self, msg):
        if self._has_gpu:
            return torch.cuda.nvtx.range_push(msg)

    def range_pop(self):
        if self._has_gpu:
            return torch.cuda.nvtx.range_pop()

    def lazy_call(self, callback):
        return torch.cuda._lazy_call(callback) if self._has_gpu else callback()

    # ---- tensor helpers
    def _tensor_type(self, dtype):
        return functools.partial(torch.tensor, dtype=dtype, device=self.current_device_name())

    @property
    def BFloat16Tensor(self):
        return self._tensor_type(torch.bfloat16)

    @property
    def ByteTensor(self):
        return self._tensor_type(torch.uint8)

    @property
    def DoubleTensor(self):
        return self._tensor_type(torch.float64)

    @property
    def FloatTensor(self):
        return self._tensor_type(torch.float32)

    @property
    def HalfTensor(self):
        return self._tensor_type(torch.float16)

    @property
    def IntTensor(self):
        return self._tensor_type(torch.int32)

    @property
    def LongTensor(self):
        return self._tensor_type(torch.int64)

    def pin_memory(self, tensor):
        return tensor.pin_memory() if self._has_gpu else tensor

    def on_accelerator(self, tensor):
        return tensor.is_cuda


internlm_accelerator = None


def get_accelerator() -> B200Accelerator:
    global internlm_accelerator
    if internlm_accelerator is None:
        internlm_accelerator = B200Accelerator()
    return internlm_accelerator


get_accelerator()

# reference class names (``accelerator/abstract_accelerator.py`` / ``cuda_accelerator.py``); there is one backend
Accelerator = B200Accelerator
CUDA_Accelerator = B200Accelerator

__all__ = ["AcceleratorType", "get_accelerator", "internlm_accelerator", "B200Accelerator", "Accelerator", "CUDA_Accelerator"]
