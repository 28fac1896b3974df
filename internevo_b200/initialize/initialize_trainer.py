"""``initialize_trainer``: pick the scheduler for the parallel layout and assemble Engine + Trainer (reference
``internlm/initialize/initialize_trainer.py:31-137``)."""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Tuple

from torch import nn
from torch.nn.modules.loss import _Loss
from torch.utils.data import DataLoader

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.core.engine import Engine
from internevo_b200.core.gradient_handler import PipelineSharedModuleGradientHandler
from internevo_b200.core.scheduler import (
    InterleavedPipelineScheduler,
    NonPipelineScheduler,
    PipelineScheduler,
    get_tensor_shape,
)
from internevo_b200.core.trainer import Trainer
from internevo_b200.data.datasets import unpack_data
from internevo_b200.solver.schedulers import Beta2Scheduler
from internevo_b200.utils.common import SchedulerHook


def initialize_trainer(model: nn.Module, optimizer, criterion: Optional[_Loss] = None,
                       train_dataloader: Optional[Iterable] = None, test_dataloader: Optional[Iterable] = None,
                       lr_scheduler=None, beta2_scheduler: Optional[Beta2Scheduler] = None,
                       scheduler_hooks: Optional[List[SchedulerHook]] = None
                       ) -> Tuple[Trainer, DataLoader, DataLoader, object]:
    if isinstance(model, nn.Module):
        model = model.to(next(model.parameters()).device)
    clip_grad_norm = gpc.config.hybrid_zero_optimizer.get("clip_grad_norm", 0.0)
    assert isinstance(gpc.config.parallel.pipeline, dict) or hasattr(gpc.config.parallel.pipeline, "get")
    pp_size = gpc.config.parallel.pipeline.get("size", 1)
    tensor_shape = get_tensor_shape()
    use_interleaved = hasattr(gpc.config, "model") and gpc.config.model.get("num_chunks", 1) > 1
    scatter_gather = gpc.is_initialized(ParallelMode.TENSOR) and gpc.get_world_size(ParallelMode.TENSOR) > 1
    data_fn: Optional[Callable] = None if gpc.config.data.get("use_packed_dataset", True) else unpack_data
    for h in scheduler_hooks or []:
        if hasattr(h, "bind_criterion"):
            h.bind_criterion(criterion)
    if pp_size > 1:
        gpc.config.NUM_MICRO_BATCHES = gpc.config.data.micro_num
        if use_interleaved:
            overlap = gpc.config.parallel["pipeline"].get("interleaved_overlap", False)
            scheduler = InterleavedPipelineScheduler(
                num_microbatches=gpc.config.NUM_MICRO_BATCHES, num_chunks=gpc.config.model.num_chunks,
                dtype=gpc.config.model["dtype"], tensor_shape=tensor_shape, scatter_gather_tensors=scatter_gather,
                scheduler_hooks=scheduler_hooks, communication_overlap=overlap, data_process_func=data_fn)
        else:
            scheduler = PipelineScheduler(
                data_process_func=data_fn, num_microbatches=gpc.config.NUM_MICRO_BATCHES,
                dtype=gpc.config.model["dtype"], tensor_shape=tensor_shape, scatter_gather_tensors=scatter_gather,
                scheduler_hooks=scheduler_hooks)
    else:
        scheduler = NonPipelineScheduler(data_process_func=data_fn,
                                         gradient_accumulation_size=gpc.config.data.gradient_accumulation,
                                         scheduler_hooks=scheduler_hooks)
    gradient_handlers = [PipelineSharedModuleGradientHandler(model=model, optimizer=optimizer)]
    engine = Engine(model=model, optimizer=optimizer, lr_scheduler=lr_scheduler, beta2_scheduler=beta2_scheduler,
                    criterion=criterion, gradient_handlers=gradient_handlers, clip_grad_norm=clip_grad_norm)
    return Trainer(engine, scheduler), train_dataloader, test_dataloader, lr_scheduler
