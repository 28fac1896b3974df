"""Gradient handlers run by ``Engine.step`` before the optimizer (reference ``internlm/core/gradient_handler.py``)."""
from abc import ABC, abstractmethod
from collections import defaultdict

import torch
import torch.distributed as dist

from internevo_b200.core.context import global_context as gpc  # noqa: F401


class BaseGradientHandler(ABC):
    def __init__(self, model, optimizer):
        self._model = model
        self._optimizer = optimizer

    @abstractmethod
    def handle_gradient(self):
        """reduce / synchronise gradients"""


class PipelineSharedModuleGradientHandler(BaseGradientHandler):
    """All-reduce (SUM) the gradients of parameters tagged with ``pipeline_shared_module_pg`` (tied modules living on
    several pipeline stages). Nothing in the shipped models sets the tag; kept for API parity."""

    def handle_gradient(self):
        buckets = defaultdict(list)
        for p in self._model.parameters():
            group = getattr(p, "pipeline_shared_module_pg", None)
            g = p.grad if p.grad is not None else (p.grad_buf if getattr(p, "grad_ready", False) else None)
            if group is not None and g is not None:
                buckets[group].append(g)
        for group, grads in buckets.items():
            if group in (None,) or dist.get_world_size(group) <= 1:
                continue
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            off = 0
            for g in grads:
                g.copy_(flat[off: off + g.numel()].view_as(g))
                off += g.numel()
