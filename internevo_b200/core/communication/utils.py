"""Shape handshake for pipeline stages and the parameter-sync handler (reference
``internlm/core/communication/utils.py``)."""
from __future__ import annotations

from typing import List, Tuple, Union

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import get_current_device

TensorShape = Union[torch.Size, List[int], Tuple[int]]


def send_meta_helper(obj, next_rank, tensor_kwargs):
    send_shape = torch.tensor(obj.size(), **tensor_kwargs)
    send_ndims = torch.tensor(len(obj.size()), **tensor_kwargs)
    dist.send(send_ndims, next_rank)
    dist.send(send_shape, next_rank)


def send_obj_meta(obj, next_rank=None):
    """Tell the next stage the shape(s) it is about to receive (ndims then dims, int64)."""
    if gpc.is_pipeline_last_stage():
        return
    if next_rank is None:
        next_rank = gpc.get_next_global_rank(ParallelMode.PIPELINE)
    tensor_kwargs = {"dtype": torch.long, "device": get_current_device()}
    if isinstance(obj, torch.Tensor):
        dist.send(torch.tensor(1, **tensor_kwargs), next_rank)
        send_meta_helper(obj, next_rank, tensor_kwargs)
    else:
        dist.send(torch.tensor(len(obj), **tensor_kwargs), next_rank)
        for t in obj:
            send_meta_helper(t, next_rank, tensor_kwargs)


def recv_meta_helper(prev_rank, tensor_kwargs):
    recv_ndims = torch.empty((), **tensor_kwargs)
    dist.recv(recv_ndims, prev_rank)
    recv_shape = torch.empty(int(recv_ndims.item()), **tensor_kwargs)
    dist.recv(recv_shape, prev_rank)
    return recv_shape


def recv_obj_meta(prev_rank=None) -> torch.Size:
    if gpc.is_pipeline_first_stage():
        return None
    if prev_rank is None:
        prev_rank = gpc.get_prev_global_rank(ParallelMode.PIPELINE)
    tensor_kwargs = {"dtype": torch.long, "device": get_current_device()}
    recv_obj_nums = torch.empty((), **tensor_kwargs)
    dist.recv(recv_obj_nums, prev_rank)
    if recv_obj_nums.item() == 1:
        return torch.Size(recv_meta_helper(prev_rank, tensor_kwargs).tolist())
    return [torch.Size(recv_meta_helper(prev_rank, tensor_kwargs).tolist()) for _ in range(int(recv_obj_nums.item()))]


class ParamAsyncBcastHandler:
    """``overlap_sync_param``: the reference registers per-module pre-forward hooks that wait for that module's
    parameter broadcast (``communication/utils.py:134-235``).  With arena all-gather there is one handle per parameter
    group; the hook on the first sub-module waits for all of them."""

    def __init__(self, zero1_mode: ParallelMode, model, isp_communicator=None) -> None:
        self._optimizer = None
        modules = model if isinstance(model, torch.nn.ModuleList) else [model]

        self._extra_handles: list = []

        def _pre_forward(module, inputs):
            if self._optimizer is not None:
                self._optimizer.wait_param_sync()
            while self._extra_handles:
                _, h = self._extra_handles.pop()
                if h is not None:
                    h.wait()

        for m in modules:
            m.register_forward_pre_hook(_pre_forward)

    def bind(self, optimizer):
        self._optimizer = optimizer

    def get_rank_by_param(self, param) -> int:
        """Rank of the ZeRO group that owns (the start of) ``param``'s optimizer state."""
        g = self._optimizer._state_of(param) if self._optimizer is not None else None
        return g.owner_of(param) if g is not None else 0

    def add_bcast_handle(self, rank, handle) -> None:
        """Track an in-flight parameter redistribution started outside the optimizer; waited for before the next forward."""
        self._extra_handles.append((rank, handle))


def split_tensor_into_1d_equal_chunks(tensor: torch.Tensor, new_buffer: bool = False) -> torch.Tensor:
    """This tensor-parallel rank's 1/tp slice of the flattened tensor — the "scatter" half of scatter-gather pipeline
    p2p (reference ``core/communication/utils.py:96-114``); a view unless ``new_buffer``."""
    tp, r = gpc.get_world_size(ParallelMode.TENSOR), gpc.get_local_rank(ParallelMode.TENSOR)
    n = tensor.numel() // tp
    part = tensor.reshape(-1)[r * n: (r + 1) * n]
    return part.clone() if new_buffer else part


def gather_split_1d_tensor(tensor: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`split_tensor_into_1d_equal_chunks`: all-gather the slices over the TENSOR group (flat result)."""
    tp = gpc.get_world_size(ParallelMode.TENSOR)
    if tp <= 1:
        return tensor.reshape(-1)
    out = torch.empty(tensor.numel() * tp, dtype=tensor.dtype, device=tensor.device)
    with torch.no_grad():
        dist.all_gather_into_tensor(out, tensor.detach().contiguous().view(-1), group=gpc.get_group(ParallelMode.TENSOR))
    return out
