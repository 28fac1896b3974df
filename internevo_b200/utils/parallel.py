"""Parameter-class predicates and init-time parameter synchronisation (reference ``internlm/utils/parallel.py``)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from internevo_b200.core.context import (
    IS_REPLICA_ZERO_PARALLEL,
    IS_TENSOR_DATA_PARALLEL,
    IS_TENSOR_EXPERT_DATA_PARALLEL,
    IS_TENSOR_ZERO_PARALLEL,
    IS_WEIGHT_ZERO_PARALLEL,
    ParallelMode,
)
from internevo_b200.core.context import global_context as gpc


def is_using_sequence_parallel():
    par = gpc.config.get("parallel", None) if gpc.config is not None else None
    if not par or "tensor" not in par:
        return False
    return isinstance(par["tensor"], dict) and par["tensor"].get("mode", "mtp") != "mtp" and par["tensor"]["size"] > 1


def is_using_isp():
    par = gpc.config.get("parallel", None) if gpc.config is not None else None   # partial configs (data-only tests, tools)
    if not par or "tensor" not in par:
        return False
    return isinstance(par["tensor"], dict) and par["tensor"].get("mode", "mtp") == "isp"


def is_replica_zero_parallel_parameter(p):
    return getattr(p, IS_REPLICA_ZERO_PARALLEL, False)


def is_tensor_data_parallel_parameter(p):
    return gpc.is_initialized(ParallelMode.TENSOR) and is_using_isp() and getattr(p, IS_TENSOR_DATA_PARALLEL, False)


def is_tensor_zero_parallel_parameter(p):
    return gpc.is_initialized(ParallelMode.TENSOR) and not is_using_isp() and getattr(p, IS_TENSOR_ZERO_PARALLEL, False)


def is_weight_zero_parallel_parameter(p):
    return gpc.is_initialized(ParallelMode.WEIGHT) and is_using_isp() and getattr(p, IS_WEIGHT_ZERO_PARALLEL, False)


def is_tensor_expert_data_parallel_parameter(p):
    return gpc.is_initialized(ParallelMode.TENSOR) and getattr(p, IS_TENSOR_EXPERT_DATA_PARALLEL, False)


def is_expert_param(p):
    return getattr(p, "is_expert", False)


def _bcast(t: torch.Tensor, mode: ParallelMode):
    group = gpc.get_group(mode)
    if group is None or gpc.get_world_size(mode) <= 1:
        return
    dist.broadcast(t, src=gpc.get_ranks_in_group(mode)[0], group=group)


def sync_model_param(model):
    """Broadcast parameters from the first rank of each (weight-)data parallel group; expert params over EXPERT_DATA
    (reference ``utils/parallel.py:71-87``)."""
    dp_mode = ParallelMode.WEIGHT_DATA if is_using_isp() else ParallelMode.DATA
    for param in model.parameters():
        if is_expert_param(param):
            if gpc.is_initialized(ParallelMode.EXPERT_DATA):
                _bcast(param.data, ParallelMode.EXPERT_DATA)
        else:
            _bcast(param.data, dp_mode)


def sync_model_replica_param_group(model):
    """Broadcast replicated (norm / gate) parameters inside the TP (or WP) group (reference ``:90-106``)."""
    mode = ParallelMode.WEIGHT if is_using_isp() else ParallelMode.TENSOR
    for param in model.parameters():
        if is_replica_zero_parallel_parameter(param):
            _bcast(param.data, mode)
        elif is_expert_param(param) and not getattr(param, "expert_tp_sharded", False):
            # replicated experts (the default; `moe.expert_tensor_parallel` shards them instead): every tensor (sequence) rank holds the same
            # copy, which must start identical (the tensor-parallel RNG stream differs per rank) and stays identical
            # (HybridZeroOptimizer._reduce_replica_grads)
            _bcast(param.data, ParallelMode.TENSOR)


def get_parallel_log_file_name():
    if gpc.is_rank_for_log():
        fn_prefix = "main_"
    else:
        fn_prefix = ""
    return (
        f"{fn_prefix}dp={gpc.get_local_rank(ParallelMode.DATA)}_"
        f"tp={gpc.get_local_rank(ParallelMode.TENSOR)}_pp={gpc.get_local_rank(ParallelMode.PIPELINE)}"
    )
