"""Gradient-accumulation scheduler without pipeline parallelism (reference
``internlm/core/scheduler/no_pipeline_scheduler.py:28-239``)."""
from __future__ import annotations

from typing import Any, Callable, Iterable, List, Optional

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import SchedulerHook, check_data_is_packed, conditional_context, get_current_device
from internevo_b200.utils.timeout import llm_timeout

from .base_scheduler import BaseScheduler


class NonPipelineScheduler(BaseScheduler):
    def __init__(self, data_process_func: Callable = None, gradient_accumulation_size: int = 1,
                 scheduler_hooks: Optional[List[SchedulerHook]] = None):
        self._grad_accum_size = gradient_accumulation_size
        self._grad_accum_offset = 0
        self._bsz_stride = 1
        self._hooks = scheduler_hooks or []
        super().__init__(data_process_func)

    def pre_processing(self, engine):
        pass

    def _call_hooks(self, func_name: str, *args, **kwargs) -> None:
        for hook in self._hooks:
            getattr(hook, func_name)(self, *args, **kwargs)

    def _load_accum_batch(self, data: Any, label: Any):
        _data, _label = self._load_micro_batch(data=data, label=label, offset=self._grad_accum_offset,
                                               bsz_stride=self._bsz_stride)
        self._grad_accum_offset += self._bsz_stride
        if self.data_process_func:
            _data["input_ids"] = self.data_process_func(_data["input_ids"], _data["cu_seqlens"])
            _label = self.data_process_func(_label, _data["cu_seqlens"], **self._label_pad)
            _data.pop("cu_seqlens")
            _data.pop("indexes")
            _data.pop("max_seqlen", None)
        return _data, _label

    def _train_one_batch(self, data, label, engine, forward_only=False, return_loss=True, return_output=False,
                         scale_loss: int = 1):
        is_moe = hasattr(gpc.config.model, "num_experts")
        with conditional_context(torch.no_grad(), enable=forward_only):
            self._call_hooks("before_forward", data)
            if is_moe:
                output, moe_losses = self._call_engine(engine, data)
            else:
                output = self._call_engine(engine, data)
            self._call_hooks("after_forward", output)
            self._call_hooks("post_helper_func", output, label)
            loss = moe_loss = None
            if return_loss:
                self._call_hooks("before_criterion", output, label)
                loss = self._call_engine_criterion(engine, output, label)
                self._call_hooks("after_criterion", loss)
                if is_moe and gpc.config.model.num_experts > 1:
                    moe_loss = sum(moe_losses) * gpc.config.loss.moe_loss_coeff
                    if gpc.config.parallel.sequence_parallel and gpc.get_world_size(ParallelMode.TENSOR) > 1:
                        if moe_loss.is_cuda:
                            dist.all_reduce(moe_loss, op=dist.ReduceOp.AVG, group=gpc.get_group(ParallelMode.TENSOR))
                        else:
                            dist.all_reduce(moe_loss, group=gpc.get_group(ParallelMode.TENSOR))
                            moe_loss = moe_loss / gpc.get_world_size(ParallelMode.TENSOR)
                    moe_loss = moe_loss / scale_loss
                    loss = loss / scale_loss + moe_loss
                else:
                    moe_loss = torch.zeros((), device=get_current_device())
                    loss = loss / scale_loss
        if not return_output:
            output = None
        if not forward_only:
            self._call_hooks("before_backward", None, None)
            engine.backward(loss)
            self._call_hooks("after_backward", None)
        if not return_loss:
            loss, moe_loss = None, None
        return output, loss, moe_loss

    @llm_timeout(func_name="nopp_forward_backward_step")
    def forward_backward_step(self, engine, data_iter: Iterable, forward_only: bool = False, return_loss: bool = True,
                              return_output_label: bool = True):
        assert forward_only or return_loss, "'return_loss' has to be True when 'forward_only' is False"
        batch_data, actual_batch_size = engine.load_batch(data_iter)
        micro_num = actual_batch_size if check_data_is_packed(batch_data) else actual_batch_size // gpc.config.data["micro_bsz"]
        self._grad_accum_size = max(1, micro_num)
        self._bsz_stride = actual_batch_size // self._grad_accum_size
        data, label = batch_data
        loss = 0 if return_loss else None
        moe_loss = 0 if return_loss else None
        outputs, labels = [], []
        self._grad_accum_offset = 0
        if engine.optimizer is not None and hasattr(engine.optimizer, "wait_param_sync"):
            engine.optimizer.wait_param_sync()
        for step in range(self._grad_accum_size):
            if engine.optimizer is not None:
                engine.optimizer.skip_grad_reduce = step != self._grad_accum_size - 1
            _data, _label = self._load_accum_batch(data, label)
            _output, _loss, _moe_loss = self._train_one_batch(_data, _label, engine, forward_only, return_loss,
                                                              return_output_label, self._grad_accum_size)
            if return_loss:
                loss = loss + _loss.detach()
                moe_loss = moe_loss + _moe_loss.detach()
            if return_output_label:
                outputs.append(_output)
                labels.append(_label)
        if not return_output_label:
            outputs, labels = None, None
        if hasattr(gpc.config.model, "num_experts"):
            return outputs, labels, loss, moe_loss
        return outputs, labels, loss
