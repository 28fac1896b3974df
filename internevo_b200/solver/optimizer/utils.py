"""Loss scaling (reference ``internlm/solver/optimizer/utils.py:381-543``)."""
from __future__ import annotations

from typing import Optional

from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)


class DynamicGradScaler:
    """Dynamic loss scale with growth interval / backoff / hysteresis. State is kept on the host (a python float): the
    optimizer reads the overflow flag once per step together with the grad norm, so no extra device sync is needed."""

    def __init__(self, initial_scale: float = 2**16, growth_factor: float = 2, backoff_factor: float = 0.5,
                 growth_interval: int = 1000, min_scale: Optional[float] = None, max_scale: Optional[float] = None,
                 hysteresis: int = 2):
        assert growth_factor > 1 and 0 < backoff_factor < 1 and hysteresis >= 0
        self._scale = float(initial_scale)
        self._min_scale, self._max_scale = min_scale, max_scale
        self._growth_factor, self._backoff_factor = growth_factor, backoff_factor
        self._growth_interval = growth_interval
        self._growth_step = 0
        self._hysteresis = hysteresis
        self._hysteresis_step = 0

    @property
    def scale(self) -> float:
        return self._scale

    @property
    def inv_scale(self) -> float:
        return 1.0 / self._scale

    def update(self, overflow: bool) -> None:
        if overflow:
            self._hysteresis_step += 1
            self._growth_step = 0
            if self._hysteresis_step >= self._hysteresis:
                self._scale *= self._backoff_factor
                if self._min_scale:
                    self._scale = max(self._scale, self._min_scale)
                if gpc.is_rank_for_log():
                    logger.warning(f"Overflow occurs, the loss scale is adjusted to {self._scale}")
        else:
            self._growth_step += 1
            if self._growth_step == self._growth_interval:
                self._growth_step = 0
                self._hysteresis_step = 0
                self._scale *= self._growth_factor
                if self._max_scale:
                    self._scale = min(self._scale, self._max_scale)
                if gpc.is_rank_for_log():
                    logger.warning(f"No overflow for consecutive {self._growth_interval} steps, "
                                   f"the loss scale is adjusted to {self._scale}")

    def state_dict(self):
        return {"_scale": self._scale, "_growth_step": self._growth_step, "_hysteresis_step": self._hysteresis_step}

    def load_state_dict(self, state_dict):
        self._scale = float(state_dict["_scale"])
        self._growth_step = state_dict["_growth_step"]
        self._hysteresis_step = state_dict["_hysteresis_step"]


# ---- tensor-list helpers of the reference's optimizer utilities (``solver/optimizer/utils.py:42-330``).  The arena
# optimizer does not need them (its gradients are one flat buffer and the norm is one fused kernel); they are kept for
# user code and for optimizers plugged in through ``BaseOptimizer``.
import math as _math  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors  # noqa: E402


def flatten(tensors):
    return _flatten_dense_tensors(tensors)


def unflatten(flat, tensors):
    return _unflatten_dense_tensors(flat, tensors)


def split_half_float_double(tensor_list):
    """Bucket tensors by type string, preserving first-seen order."""
    buckets = {}
    for t in tensor_list:
        buckets.setdefault(t.type(), []).append(t)
    return list(buckets.values())


def reduce_tensor(tensor, dtype=None, dst_rank=None, parallel_mode=None):
    """Asynchronous averaging all-reduce (or reduce to ``dst_rank`` of the group) in the tensor's own dtype → handle."""
    from internevo_b200.core.context import ParallelMode

    assert dtype is None or dtype == tensor.dtype, "communication happens in the tensor's dtype"
    mode = ParallelMode.DATA if parallel_mode is None else parallel_mode
    group, n = gpc.get_group(mode), gpc.get_world_size(mode)
    if group is None or n <= 1:
        return None
    avg = dist.ReduceOp.AVG if tensor.is_cuda else dist.ReduceOp.SUM          # gloo has no AVG
    if not tensor.is_cuda:
        tensor.div_(n)
    if dst_rank is None:
        return dist.all_reduce(tensor, op=avg, group=group, async_op=True)
    return dist.reduce(tensor, dst=gpc.get_ranks_in_group(mode)[dst_rank], op=avg, group=group, async_op=True)


def has_inf_or_nan(tensor) -> bool:
    s = float(tensor.float().sum())
    return _math.isinf(s) or _math.isnan(s)


def release_param_grad(tensor_list):
    for t in tensor_list:
        t.grad = None


def sync_param(flat_tensor, tensor_list):
    """Re-point every tensor of ``tensor_list`` at its slice of ``flat_tensor``."""
    for p, view in zip(tensor_list, _unflatten_dense_tensors(flat_tensor, tensor_list)):
        p.data = view


def multi_tensor_l2norm_torch(tensor_list, per_tensor: bool):
    norms = torch.stack([t.float().norm(2) for t in tensor_list])
    return norms.norm(2).unsqueeze(0), (norms if per_tensor else norms.new_empty(0))


def calc_l2_norm(grads):
    """L2 norm of a tensor list; contiguous bf16 / fp32 CUDA tensors go through the fused sum-of-squares kernel."""
    if len(grads) == 0:
        return 0.0
    if all(g.is_cuda and g.is_contiguous() for g in grads):
        from internevo_b200.ops import sumsq_

        acc = torch.zeros(1, dtype=torch.float32, device=grads[0].device)
        for g in grads:
            sumsq_(g, acc)
        return acc.sqrt()
    return multi_tensor_l2norm_torch(grads, False)[0]


def calc_lp(grads, norm_type):
    total = 0.0
    for g in grads:
        total = total + g.float().norm(norm_type) ** norm_type
    return total


def get_norm(grads, norm_type, enable_cuda_kernels: bool = True):
    """``max|g|`` for the inf-norm, else ``sum ||g||_p ** p`` (the caller reduces over ranks and takes the root)."""
    if norm_type == _math.inf:
        return max(g.detach().abs().max() for g in grads)
    if norm_type == 2.0 and enable_cuda_kernels:
        return calc_l2_norm(grads) ** norm_type
    return calc_lp(grads, norm_type)


class BaseGradScaler:
    """Constant loss scale; ``DynamicGradScaler`` adds the growth / back-off policy."""

    def __init__(self, initial_scale: float):
        assert initial_scale > 0
        self._scale = float(initial_scale)

    @property
    def scale(self) -> float:
        return self._scale

    @property
    def inv_scale(self) -> float:
        return 1.0 / self._scale

    def state_dict(self):
        return {"scale": self._scale}

    def load_state_dict(self, state_dict) -> None:
        self._scale = float(state_dict["scale"])

    def update(self, overflow: bool) -> None:
        pass


# ---- list-based global gradient norm (reference ``solver/optimizer/utils.py:25-39,225-378``).  The arena optimizer computes the
# same quantity from its flat shard inside the reduce kernels (``HybridZeroOptimizer._group_sumsq``); this is the general form
# for arbitrary (gradient, parameter) lists.
def get_grad_accumulate_object(tensor):
    """The ``AccumulateGrad`` node of a leaf parameter (hooks registered on it fire when the parameter's gradient is complete).
    ``expand_as`` makes a temporary non-leaf whose first ``next_functions`` entry is that node."""
    if tensor.grad_fn is not None:
        raise RuntimeError("get_grad_accumulate_object() takes a leaf tensor; compute graphs are built from parameters")
    acc = tensor.expand_as(tensor).grad_fn.next_functions[0][0]
    assert acc is not None and "AccumulateGrad" in type(acc).__name__
    return acc


def _counts_toward_norm(p, model_mode) -> bool:
    """Does THIS rank contribute ``p``'s gradient to the global norm?  Sharded parameters (tensor / weight / expert shards)
    count everywhere - every rank holds a different piece; parameters replicated over the model-parallel group (norm weights,
    gates) count on that group's rank 0 only; pipeline-shared modules on the first rank of their sharing group."""
    from internevo_b200.utils import parallel as _par

    shared = getattr(p, "pipeline_shared_module_pg", None)
    if shared is not None:
        return dist.get_rank(shared) == 0
    if _par.is_replica_zero_parallel_parameter(p):
        return not gpc.is_initialized(model_mode) or gpc.get_local_rank(model_mode) == 0
    if (_par.is_tensor_data_parallel_parameter(p) or _par.is_tensor_zero_parallel_parameter(p)
            or _par.is_weight_zero_parallel_parameter(p) or _par.is_tensor_expert_data_parallel_parameter(p)):
        return True
    # untagged parameters are treated like replicas: one copy per model-parallel group
    return not gpc.is_initialized(model_mode) or gpc.get_local_rank(model_mode) == 0


def _model_mode():
    from internevo_b200.core.context import ParallelMode
    from internevo_b200.utils.parallel import is_using_isp

    return ParallelMode.WEIGHT if is_using_isp() else ParallelMode.TENSOR


def reduce_grads(gradients, parameters, weight_parallel_mode=None):
    """fp32 copies of the gradients this rank contributes to the global norm (see ``_counts_toward_norm``)."""
    mode = _model_mode() if weight_parallel_mode is None else weight_parallel_mode
    return [g.data.float() for g, p in zip(gradients, parameters) if _counts_toward_norm(p, mode)]


def _all_reduce_scalar(t, mode, op):
    if mode is not None and gpc.is_initialized(mode) and gpc.get_world_size(mode) > 1:
        dist.all_reduce(t, op=op, group=gpc.get_group(mode))


def compute_norm(gradients, parameters, norm_type=2, zero_mode=None):
    """``norm ** norm_type`` of the gradients of ONE parameter group over the whole job (take the root before use).  The local
    contribution is reduced over the model-parallel group the group's parameters are sharded over (tensor, or weight under
    ISP), the pipeline stages and the ZeRO group that shards the gradients; expert groups are additionally combined over the
    expert group, scaled by the data-parallel size.  ``-1`` stands for an infinite norm, ``-2`` for NaN (the step is skipped)."""
    from internevo_b200.core.context import ParallelMode
    from internevo_b200.utils import parallel as _par

    zero_mode = ParallelMode.ZERO1 if zero_mode is None else zero_mode
    gradients, parameters = list(gradients), list(parameters)
    device = gradients[0].device
    norm_type = float(norm_type)
    model_mode = _model_mode()
    first = parameters[0]
    if _par.is_tensor_data_parallel_parameter(first):
        # ISP embedding group: one full copy per rank of the tensor (sequence) group in this framework - nothing to add up
        # (the reference splits it along the hidden dim and sums over TENSOR, ``optimizer/utils.py:340-342``)
        shard_mode = None
    elif _par.is_tensor_zero_parallel_parameter(first):
        shard_mode = ParallelMode.TENSOR
    else:
        shard_mode = model_mode
    if norm_type == _math.inf:
        total = torch.stack([g.data.abs().max().float() for g in gradients]).max().reshape(1).to(device)
        _all_reduce_scalar(total, shard_mode, dist.ReduceOp.MAX)
        _all_reduce_scalar(total, ParallelMode.PIPELINE, dist.ReduceOp.MAX)
    else:
        mine = reduce_grads(gradients, parameters, model_mode)
        total = torch.zeros(1, dtype=torch.float32, device=device)
        if mine:
            total += torch.as_tensor(get_norm(mine, norm_type, enable_cuda_kernels=device.type != "cpu"),
                                     dtype=torch.float32, device=device).reshape(1)
        _all_reduce_scalar(total, shard_mode, dist.ReduceOp.SUM)
        _all_reduce_scalar(total, ParallelMode.PIPELINE, dist.ReduceOp.SUM)
        _all_reduce_scalar(total, zero_mode, dist.ReduceOp.SUM)
    if zero_mode == ParallelMode.EXPERT_DATA and gpc.is_initialized(ParallelMode.EXPERT):
        # expert gradients are not synchronised over the expert group: combine the per-rank norms there
        total = total / float(gpc.get_world_size(ParallelMode.DATA))
        _all_reduce_scalar(total, ParallelMode.EXPERT, dist.ReduceOp.SUM)
    value = float(total)
    if _math.isinf(value):
        return -1
    if _math.isnan(value):
        return -2
    return value
