"""Bucket / gradient / parameter book-keeping containers (reference ``internlm/solver/optimizer/store.py:13-322``).

``HybridZeroOptimizer`` here does not need them: parameters, gradients and optimizer states live in flat, range-interleaved
arenas (``hybrid_zero_optim._GroupState``) and the reduce / update kernels walk ranges of those.  The containers are kept as
small, self-contained utilities for code that builds its own reduction schedule on top of the framework (custom optimizers,
the FSDP adapter, user scripts written against the reference), and ``stores_of(optimizer)`` fills a set of them from a live
optimizer's arenas so such code can inspect ownership and reduced gradients through the familiar interface.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional

import torch
from torch import Tensor
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc


class BaseStore:
    """Knows the size of, and this rank's position in, the group gradients are reduced over."""

    def __init__(self, dp_parallel_mode=ParallelMode.DATA):
        on = gpc.is_initialized(dp_parallel_mode)
        self._world_size = gpc.get_world_size(dp_parallel_mode) if on else 1
        self._local_rank = gpc.get_local_rank(dp_parallel_mode) if on else 0

    @property
    def world_size(self) -> int:
        return self._world_size

    @property
    def local_rank(self) -> int:
        return self._local_rank


class BucketStore(BaseStore):
    """Gradients (and their parameters) waiting for one reduction, keyed by the rank they are reduced to (``None``: all-reduce)."""

    def __init__(self, group_id, dp_parallel_mode):
        super().__init__(dp_parallel_mode)
        self._group_id, self._dp_parallel_mode = group_id, dp_parallel_mode
        self.reset()

    def reset(self) -> None:
        self._grads: Dict[Optional[int], List[Tensor]] = defaultdict(list)
        self._params: Dict[Optional[int], List[Tensor]] = defaultdict(list)
        self._numel: Dict[Optional[int], int] = defaultdict(int)

    def reset_by_rank(self, reduce_rank: Optional[int] = None) -> None:
        for d in (self._grads, self._params, self._numel):
            d.pop(reduce_rank, None)

    def get_param_group_id(self):
        return self._group_id

    def get_dp_parallel_mode(self):
        return self._dp_parallel_mode

    def num_elements_in_bucket(self, reduce_rank: Optional[int] = None) -> int:
        return self._numel[reduce_rank]

    def num_params_in_bucket(self, reduce_rank: Optional[int] = None) -> int:
        return len(self._params[reduce_rank])

    def add_num_elements_in_bucket(self, num_elements: int, reduce_rank: Optional[int] = None) -> None:
        self._numel[reduce_rank] += num_elements

    def add_grad(self, tensor: Tensor, reduce_rank: Optional[int] = None) -> None:
        self._grads[reduce_rank].append(tensor)

    def add_param(self, tensor: Tensor, reduce_rank: Optional[int] = None) -> None:
        self._params[reduce_rank].append(tensor)

    def get_grad(self, reduce_rank: Optional[int] = None) -> List[Tensor]:
        return self._grads[reduce_rank]

    def get_param(self, reduce_rank: Optional[int] = None) -> List[Tensor]:
        return self._params[reduce_rank]


class GradientStore(BaseStore):
    """Reduced (averaged) gradients per parameter group, plus the autograd accumulator nodes whose hooks must stay alive."""

    def __init__(self, *args):
        super().__init__(*args)
        self._averaged_gradients: Dict[int, List[Tensor]] = defaultdict(list)
        self._grad_acc_objs: list = []

    def add_accumulate_grad_object(self, obj) -> None:
        self._grad_acc_objs.append(obj)

    def get_averaged_gradients_by_group(self, group_id: int) -> List[Tensor]:
        return self._averaged_gradients[group_id]

    def add_average_gradient_by_group(self, group_id: int, tensor: Tensor) -> None:
        self._averaged_gradients[group_id].append(tensor)

    def reset_average_gradients_by_group(self, group_id: int) -> None:
        self._averaged_gradients[group_id] = []


class ParameterStore(BaseStore):
    """Which rank owns a parameter's optimizer state, the low-precision parameters (and their flat copy) per (rank, group), and
    which parameters have had their gradient reduced in the current step."""

    def __init__(self, dp_paralle_mode):
        super().__init__(dp_paralle_mode)
        self._param_to_rank: Dict[int, int] = {}
        self._fp16_params: Dict[tuple, List[Tensor]] = defaultdict(list)
        self._flat_fp16: Dict[tuple, Tensor] = {}
        self._is_param_reduced: Dict[int, bool] = {}
        self._keep: Dict[int, Tensor] = {}
        self._reduced_param: List[Tensor] = []
        self.reset_reduced_data_for_compute_norm()

    def set_param_to_rank(self, tensor: Tensor, rank: int) -> None:
        self._param_to_rank[id(tensor)] = rank
        self._keep[id(tensor)] = tensor

    def get_param_rank(self, tensor: Tensor) -> int:
        return self._param_to_rank[id(tensor)]

    def belongs_to_current_rank(self, tensor: Tensor) -> bool:
        return self.get_param_rank(tensor) == self.local_rank

    def add_fp16_param_list_by_rank_group(self, rank, group_id, tensor_list) -> None:
        self._fp16_params[(rank, group_id)].extend(tensor_list)

    def get_fp16_params_by_rank_group(self, rank, group_id) -> List[Tensor]:
        return self._fp16_params[(rank, group_id)]

    def add_flat_fp16_param_by_rank_group(self, rank, group_id, tensor) -> None:
        self._flat_fp16[(rank, group_id)] = tensor

    def get_flat_fp16_param_by_rank_group(self, rank, group_id) -> Tensor:
        return self._flat_fp16[(rank, group_id)]

    def is_param_reduced(self, tensor) -> bool:
        return self._is_param_reduced.get(id(tensor), False)

    def set_param_reduction_state(self, tensor, state: bool) -> None:
        self._is_param_reduced[id(tensor)] = state
        self._keep[id(tensor)] = tensor

    def get_param_reduction_states(self) -> Dict[Tensor, bool]:
        return {self._keep[k]: v for k, v in self._is_param_reduced.items()}

    def reset_previous_reduced_params(self) -> None:
        self._reduced_param = []

    def add_previous_reduced_param(self, tensor) -> None:
        self._reduced_param.append(tensor)

    def clear_grads_of_previous_reduced_params(self) -> None:
        for p in self._reduced_param:
            p.grad = None
        self.reset_previous_reduced_params()

    def add_reduced_param_for_compute_norm(self, param) -> None:
        gid = getattr(param, "group_id", 0)
        self._bucket_reduced_param[gid].append(param)
        self._bucket_reduced_grad[gid].append(param.grad)

    def get_reduced_param_for_compute_norm(self, group_id=0):
        return self._bucket_reduced_param[group_id], self._bucket_reduced_grad[group_id]

    def reset_reduced_data_for_compute_norm(self) -> None:
        self._bucket_reduced_param: Dict[int, list] = defaultdict(list)
        self._bucket_reduced_grad: Dict[int, list] = defaultdict(list)


class TensorBucket:
    """Collect tensors up to ``size`` elements, flatten them into one buffer for a collective, copy the result back."""

    def __init__(self, size: int):
        self._max_size = size
        self._unflatten_and_copy_flag = False
        self.empty()

    @property
    def max_size(self) -> int:
        return self._max_size

    @property
    def current_size(self) -> int:
        return self._current_size

    def is_full_or_oversized(self) -> bool:
        return self._current_size >= self._max_size

    def is_empty(self) -> bool:
        return not self._bucket

    def will_exceed_max_size(self, tensor_size: int) -> bool:
        return self._current_size + tensor_size > self._max_size

    def set_unflatten_and_copy_flag(self, flag: bool) -> None:
        self._unflatten_and_copy_flag = flag

    def get_unflatten_and_copy_flag(self) -> bool:
        return self._unflatten_and_copy_flag

    def add_to_bucket(self, tensor: Tensor, allow_oversize: bool = False) -> None:
        n = tensor.numel()
        if self.will_exceed_max_size(n) and not allow_oversize:
            raise RuntimeError(f"a tensor of {n} elements does not fit: {self._current_size} of {self._max_size} are taken")
        self._bucket.append(tensor)
        self._current_size += n

    def get_bucket(self) -> List[Tensor]:
        return self._bucket

    def get_flat_tensor(self) -> Optional[Tensor]:
        return self._flat_tensor

    def empty(self) -> None:
        self._bucket: List[Tensor] = []
        self._current_size = 0
        self._flat_tensor: Optional[Tensor] = None

    def flatten(self) -> Tensor:
        self._flat_tensor = _flatten_dense_tensors(self._bucket)
        return self._flat_tensor

    def unflatten_and_copy(self) -> None:
        if self._unflatten_and_copy_flag and self._flat_tensor is not None:
            for old, new in zip(self._bucket, _unflatten_dense_tensors(self._flat_tensor, self._bucket)):
                old.copy_(new)


def stores_of(optimizer):
    """``(param_store, grad_store, bucket_stores)`` describing a live ``HybridZeroOptimizer``: for every parameter group the rank
    (of its ZeRO group) whose shard holds the FIRST element of each parameter, the bf16 parameters and their flat arena, and -
    after a reduction - this rank's reduced gradient shard."""
    groups = [g for g in getattr(optimizer, "groups", []) if g.params]
    mode = groups[0].zero_mode if groups else ParallelMode.DATA
    params, grads = ParameterStore(mode), GradientStore(mode)
    buckets = []
    for g in groups:
        buckets.append(BucketStore(g.gid, g.dp_mode))
        for p in g.params:
            owner = g.owner_of(p) if hasattr(g, "owner_of") else 0
            params.set_param_to_rank(p, owner)
            params.add_fp16_param_list_by_rank_group(owner, g.gid, [p])
        params.add_flat_fp16_param_by_rank_group(params.local_rank, g.gid, g.param_arena)
        grads.add_average_gradient_by_group(g.gid, g.grad_shard())
    return params, grads, buckets


__all__ = ["BaseStore", "BucketStore", "GradientStore", "ParameterStore", "TensorBucket", "stores_of"]
