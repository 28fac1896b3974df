"""Mixed-precision model wrapper (reference ``internlm/core/naive_amp.py:38-208``): casts the module to the low
precision dtype, keeps modules tagged fp32 in fp32 (casting their inputs/outputs with hooks), converts inputs.

On B200 the loss kernel consumes bf16 logits directly with fp32 math, so ``output_to_fp32`` defaults to False in
``initialize_model`` (the reference materialises an fp32 copy of the ``[T, V]`` logits every micro-batch)."""
from __future__ import annotations

from typing import Any

import torch
import torch.distributed as dist
from torch import Tensor, nn
from torch.distributed import ReduceOp

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc


def set_fp32_attr_to_module(module: nn.Module):
    setattr(module, "is_fp32_module", True)


def module_has_fp32_attr(module: nn.Module):
    return getattr(module, "is_fp32_module", False)


def set_output_attr_to_module(module: nn.Module):
    setattr(module, "is_output", True)


def module_is_output(module: nn.Module):
    return getattr(module, "is_output", False)


class NaiveAMPModel(nn.Module):
    def __init__(self, model: nn.Module, output_to_fp32: bool = True, parallel_mode: ParallelMode = ParallelMode.DATA,
                 sync_buffer: bool = True, dtype=torch.float16):
        super().__init__()
        self.model = model.to(dtype)
        self._output_to_fp32 = output_to_fp32
        self._sync_buf = sync_buffer
        self.dtype = dtype
        if gpc.is_initialized(parallel_mode):
            self._process_group = gpc.get_group(parallel_mode)
            self._world_size = gpc.get_world_size(parallel_mode)
        else:
            self._process_group, self._world_size, self._sync_buf = None, 1, False
        self._first_eval_run = False
        self._register_fp32_parameters_hook()

    @property
    def sync_buffer(self):
        return self._sync_buf

    @sync_buffer.setter
    def sync_buffer(self, state: bool):
        self._sync_buf = state

    def _convert_to_fp16(self, input_: Any):
        if isinstance(input_, Tensor) and input_.dtype == torch.float32:
            input_ = input_.to(self.dtype)
        return input_

    def _convert_to_fp32(self, input_: Any):
        if isinstance(input_, Tensor) and input_.dtype == self.dtype:
            input_ = input_.float()
        return input_

    def convert_to_fp32(self, out):
        if isinstance(out, Tensor):
            return self._convert_to_fp32(out)
        if isinstance(out, (tuple, list)):
            return type(out)(self.convert_to_fp32(v) for v in out)
        if isinstance(out, dict):
            return {k: self.convert_to_fp32(v) for k, v in out.items()}
        return out

    def _reduce_module_buffer(self):
        bufs = [b for b in self.model.buffers() if b is not None]
        if bufs and self._process_group is not None and self._world_size > 1:
            flat = torch.cat([b.reshape(-1).float() for b in bufs])
            dist.all_reduce(flat, op=ReduceOp.SUM, group=self._process_group)
            flat.div_(self._world_size)
            off = 0
            for b in bufs:
                b.copy_(flat[off: off + b.numel()].view_as(b))
                off += b.numel()

    def forward(self, *args, **kwargs):
        if self.training:
            self._first_eval_run = False
        elif not self._first_eval_run and self._sync_buf:
            self._reduce_module_buffer()
            self._first_eval_run = True
        if args:
            args = [self._convert_to_fp16(a) for a in args]
        if kwargs:
            kwargs = {k: self._convert_to_fp16(v) for k, v in kwargs.items()}
        out = self.model(*args, **kwargs)
        if self._output_to_fp32:
            out = self.convert_to_fp32(out)
        return out

    def _register_fp32_parameters_hook(self) -> None:
        """fp32-tagged sub-modules keep fp32 weights; hooks cast their inputs up and outputs back down."""
        dtype = torch.float32

        def to_dtype(x, dt):
            if isinstance(x, Tensor) and x.is_floating_point():
                return x.to(dt)
            if isinstance(x, (tuple, list)):
                return type(x)(to_dtype(v, dt) for v in x)
            return x

        def _pre(model, inputs):
            assert isinstance(inputs, tuple)
            return to_dtype(inputs, dtype)

        def _post(model, inputs, outputs):
            return to_dtype(outputs, self.dtype)

        output_tf32 = bool(gpc.config.get("output_tf32", False)) if gpc.config is not None else False

        def _is_output_head(mod) -> bool:
            from internevo_b200.parallel.linear import BaseScaleColumnParallelLinear

            return module_is_output(mod) or isinstance(mod, BaseScaleColumnParallelLinear)

        modules = self.model if isinstance(self.model, nn.ModuleList) else [self.model]
        for m in modules:
            for sub in m.modules():
                if module_has_fp32_attr(sub):
                    sub.to(dtype)
                    sub.register_forward_pre_hook(_pre)
                    sub.register_forward_hook(_post)
                elif output_tf32 and _is_output_head(sub):
                    # ``output_tf32``: the LM head keeps fp32 weights and takes fp32 inputs; its GEMM runs on the TF32 tensor
                    # path and the logits stay fp32 for the loss (reference ``naive_amp.py:203-208``)
                    sub.to(dtype)
                    torch.backends.cudnn.allow_tf32 = True
                    torch.backends.cuda.matmul.allow_tf32 = True
                    sub.register_forward_pre_hook(_pre)
