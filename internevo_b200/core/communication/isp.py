"""ISP (Intern Sequence Parallel) weight-parallel communicator.

Every ``ISPLinear`` keeps a ``1/wp`` row-shard of its weight; the full weight is all-gathered over the WEIGHT group just
in time for the forward and again for the backward GEMMs, and the weight gradient is reduce-scattered (AVG) back to the
shard owner (reference ``internlm/core/communication/isp.py:31-567``).

What is kept from the reference: module ordering per block, prefetch of the *next* module's weight while the current
one computes (previous module in backward), a small rotating memory pool for gathered weights so the allocator is not
hit every layer, deferred accumulation of reduce-scattered gradients, the scheduler hook that flips forward/backward
state.  What changes: gathers are single ``all_gather_into_tensor`` calls into pooled buffers (no list API, no copies),
gradient results are accumulated straight into the optimizer's persistent gradient arena.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch import nn

from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import SchedulerHook


class ISPCommModelConfig:
    """dtype / device / activation-checkpoint fraction of the model (reference ``isp.py:31-43``)."""

    def __init__(self, dtype: torch.dtype = torch.half, device: torch.device = None, activation_checkpointing: float = 0.0,
                 module_shapes: Dict[str, torch.Size] = None) -> None:
        self.dtype = dtype
        self.device = device or (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                                 else torch.device("cpu"))
        self.activation_checkpointing = activation_checkpointing
        self.module_shapes = module_shapes


class MemoryPool:
    """Rotating pool of gathered-weight buffers keyed by shape (reference ``isp.py:45-140``): with prefetch there are at
    most two gathered weights of a given shape alive (current + next), so ``depth`` = 3 is always safe."""

    def __init__(self, dtype, device, depth: int = 3):
        self.dtype, self.device, self.depth = dtype, device, depth
        self._bufs: Dict[Tuple[int, ...], List[torch.Tensor]] = {}
        self._next: Dict[Tuple[int, ...], int] = {}

    def get(self, shape) -> torch.Tensor:
        key = tuple(shape)
        if key not in self._bufs:
            self._bufs[key] = [torch.empty(key, dtype=self.dtype, device=self.device) for _ in range(self.depth)]
            self._next[key] = 0
        i = self._next[key]
        self._next[key] = (i + 1) % self.depth
        return self._bufs[key][i]

    def reset_lazy_pools(self):
        """kept for API parity (``train.py`` calls it every step in the reference)"""


class ISPCommunicator:
    def __init__(self, model: Union[nn.Module, nn.ModuleList], model_conf: ISPCommModelConfig, overlap: bool = False,
                 enable_memory_pool: bool = False, process_group: dist.ProcessGroup = None) -> None:
        from internevo_b200.parallel.linear import ISPLinear

        self.process_group = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.overlap = overlap
        self.model_conf = model_conf
        self.is_forward = True
        self.memory_pool = MemoryPool(model_conf.dtype, model_conf.device) if enable_memory_pool else None
        chunks = list(model) if isinstance(model, nn.ModuleList) else [model]
        chunks = [getattr(c, "model", c) for c in chunks]
        self._order: List[nn.Module] = []
        for c in chunks:
            mods = c if isinstance(c, nn.ModuleList) else [c]
            for m in mods:
                self._order += [sub for sub in m.modules() if isinstance(sub, ISPLinear)]
        self._index = {id(m): i for i, m in enumerate(self._order)}
        self._gathered: Dict[int, Tuple[torch.Tensor, Optional[dist.Work]]] = {}
        self._pending_grads: List[Tuple[torch.nn.Parameter, torch.Tensor, Optional[dist.Work]]] = []

    # ---------------------------------------------------------------------------------------------------------------
    def _launch_gather(self, module, weight: torch.Tensor, async_op: bool):
        shape = (weight.shape[0] * self.world, *weight.shape[1:])
        out = self.memory_pool.get(shape) if self.memory_pool is not None else torch.empty(
            shape, dtype=weight.dtype, device=weight.device)
        h = dist.all_gather_into_tensor(out, weight.contiguous(), group=self.process_group, async_op=async_op)
        return out, (h if async_op else None)

    def all_gather_weight(self, module, weight: torch.Tensor, is_bias: bool = False, backward: bool = False):
        """Full weight (or bias) of ``module``; consumes a prefetched gather when one is in flight and starts the
        prefetch of the neighbouring module (next in forward, previous in backward)."""
        if self.world <= 1:
            return weight
        if is_bias:
            out, _ = self._launch_gather(module, weight, async_op=False)
            return out
        key = id(module)
        if key in self._gathered:
            out, h = self._gathered.pop(key)
            if h is not None:
                h.wait()
        else:
            out, _ = self._launch_gather(module, weight, async_op=False)
        if self.overlap:
            i = self._index.get(key)
            if i is not None:
                j = i - 1 if backward else i + 1
                if 0 <= j < len(self._order):
                    nxt = self._order[j]
                    fused_fwd, fused_bwd = getattr(nxt, "_b200_isp_fused", (False, False))
                    if id(nxt) not in self._gathered and not (fused_bwd if backward else fused_fwd):
                        self._gathered[id(nxt)] = self._launch_gather(nxt, nxt.weight, async_op=True)
        return out

    def release_weight(self, module):
        """Gathered buffers come from the rotating pool (or the caching allocator): nothing to free explicitly."""

    def reduce_scatter_grad(self, module, param, grad_full: torch.Tensor, is_bias: bool = False):
        """AVG reduce-scatter of the full-weight gradient to this rank's shard. With overlap the result is accumulated
        into the parameter's gradient buffer later (``flush_grads``) and ``None`` is returned to autograd."""
        if self.world <= 1:
            return grad_full
        out = torch.empty(grad_full.shape[0] // self.world, *grad_full.shape[1:], dtype=grad_full.dtype,
                          device=grad_full.device)
        if grad_full.is_cuda:
            h = dist.reduce_scatter_tensor(out, grad_full.contiguous(), op=dist.ReduceOp.AVG, group=self.process_group,
                                           async_op=self.overlap)
        else:
            h = dist.reduce_scatter_tensor(out, grad_full.contiguous(), group=self.process_group, async_op=self.overlap)
        if not self.overlap:
            if not grad_full.is_cuda:
                out.div_(self.world)
            return out
        self._pending_grads.append((param, out, h, grad_full))
        return None

    def flush_grads(self):
        """Wait for outstanding reduce-scatters and fold them into the gradient arena (or ``.grad``)."""
        for param, out, h, keep in self._pending_grads:
            if h is not None:
                h.wait()
            if not out.is_cuda:
                out.div_(self.world)
            buf = getattr(param, "grad_buf", None)
            if buf is not None:
                if getattr(param, "grad_ready", False):
                    buf.add_(out)
                else:
                    buf.copy_(out)
                    param.grad_ready = True
            elif param.grad is None:
                param.grad = out
            else:
                param.grad.add_(out)
            del keep
        self._pending_grads = []


class ISPCommunicatorSchedulerHook(SchedulerHook):
    """Flips the communicator between forward and backward mode and flushes deferred gradients (reference
    ``isp.py:529-567``)."""

    def __init__(self, overlap_handler: ISPCommunicator, zero_optim) -> None:
        self._isp_communicator = overlap_handler
        self._zero_optim = zero_optim

    def before_forward(self, scheduler, inputs) -> None:
        self._isp_communicator.is_forward = True
        gpc.is_forward = True

    def after_forward(self, scheduler, outputs) -> None:
        pass

    def before_criterion(self, scheduler, outputs, label) -> None:
        pass

    def after_criterion(self, scheduler, loss) -> None:
        pass

    def before_backward(self, scheduler, outputs, outputs_grad) -> None:
        self._isp_communicator.is_forward = False
        gpc.is_forward = False

    def after_backward(self, scheduler, inputs_grad) -> None:
        self._isp_communicator.flush_grads()
        self._isp_communicator._gathered.clear()

    def post_helper_func(self, scheduler, outputs, label) -> None:
        pass
