"""Pipeline p2p verbs on NCCL ``batch_isend_irecv`` (reference ``internlm/core/communication/p2p.py:89-584``).

Stays on NCCL p2p by design (activations cross a stage boundary once; there is no compute to fuse with).  Unlike the
reference there is no ``cuda.synchronize()`` after the batch: NCCL work is stream-ordered, ``wait()`` only inserts a
stream dependency, so the next compute kernel is queued immediately and the copy overlaps with whatever is still
running.  Asynchronous variants return the work handles so the interleaved scheduler can overlap a whole
forward/backward with the transfer.
"""
from __future__ import annotations

import operator
from functools import reduce
from typing import List, Tuple, Union

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import get_current_device

TensorShape = Union[torch.Size, List[int], Tuple[int]]


def _get_tensor_shape(tensor_shape: TensorShape, chunk_tensor: bool = False) -> Tuple[TensorShape, bool]:
    """With ``scatter_gather_tensors`` (mtp) only ``1/tp`` of the activation travels and is re-gathered on arrival."""
    if chunk_tensor:
        numel = reduce(operator.mul, tensor_shape, 1)
        tp = gpc.get_world_size(ParallelMode.TENSOR)
        if numel % tp == 0:
            return (numel // tp,), True
    return tuple(tensor_shape), False


def create_recv_buffer_with_shapes(recv_shapes, dtype, scatter_gather_tensors):
    if isinstance(recv_shapes, (torch.Size, tuple)) or (isinstance(recv_shapes, list) and recv_shapes and isinstance(recv_shapes[0], int)):
        shape, split = _get_tensor_shape(recv_shapes, scatter_gather_tensors)
        return torch.empty(shape, requires_grad=True, device=get_current_device(), dtype=dtype), split
    bufs, split = [], False
    for s in recv_shapes:
        shape, split = _get_tensor_shape(s, scatter_gather_tensors)
        bufs.append(torch.empty(shape, requires_grad=True, device=get_current_device(), dtype=dtype))
    return bufs, split


def _split_1d(t: torch.Tensor) -> torch.Tensor:
    tp, r = gpc.get_world_size(ParallelMode.TENSOR), gpc.get_local_rank(ParallelMode.TENSOR)
    flat = t.reshape(-1)
    n = flat.numel() // tp
    return flat[r * n: (r + 1) * n].contiguous()


def _gather_1d(t: torch.Tensor, shape) -> torch.Tensor:
    tp = gpc.get_world_size(ParallelMode.TENSOR)
    out = torch.empty(t.numel() * tp, dtype=t.dtype, device=t.device)
    with torch.no_grad():  # the receive buffer is a leaf that requires grad; the collective itself is not differentiated
        dist.all_gather_into_tensor(out, t.detach().contiguous(), group=gpc.get_group(ParallelMode.TENSOR))
    return out.view(shape).requires_grad_()


def process_object_to_send(obj, scatter_gather_tensors):
    if isinstance(obj, torch.Tensor):
        shape, split = _get_tensor_shape(obj.shape, scatter_gather_tensors)
        return _split_1d(obj) if split else obj.contiguous()
    return [process_object_to_send(o, scatter_gather_tensors) for o in obj]


def filling_ops_queue(obj, comm_op, comm_rank, ops_queue, group=None) -> None:
    """Append one ``P2POp`` per tensor of ``obj`` (a tensor or a list of tensors) to ``ops_queue`` - the building block of
    ``batch_isend_irecv`` exchanges (reference ``core/communication/p2p.py:79-86``)."""
    objs = obj if isinstance(obj, (list, tuple)) else [obj]
    ops_queue.extend(dist.P2POp(comm_op, o, comm_rank, group) for o in objs)


def _ops_for(obj, peer, send: bool, group=None):
    ops: list = []
    filling_ops_queue(obj, dist.isend if send else dist.irecv, peer, ops, group)
    return ops


def _communicate_async(object_send_next=None, object_send_prev=None, recv_prev=False, recv_next=False,
                       recv_prev_shape=None, recv_next_shape=None, prev_rank=None, next_rank=None, dtype=None,
                       scatter_gather_tensors=False):
    """Start the exchange; returns ``(recv_prev_buf, recv_next_buf, requests, post)`` where ``post`` finalises the
    received tensors (re-gather under scatter_gather) once the requests are waited on."""
    recv_prev_buf = recv_next_buf = None
    rp_split = rn_split = False
    if recv_prev:
        assert recv_prev_shape is not None
        recv_prev_buf, rp_split = create_recv_buffer_with_shapes(recv_prev_shape, dtype, scatter_gather_tensors)
    if recv_next:
        assert recv_next_shape is not None
        recv_next_buf, rn_split = create_recv_buffer_with_shapes(recv_next_shape, dtype, scatter_gather_tensors)
    if object_send_prev is not None or recv_prev:
        prev_rank = gpc.get_prev_global_rank(ParallelMode.PIPELINE) if prev_rank is None else prev_rank
    if object_send_next is not None or recv_next:
        next_rank = gpc.get_next_global_rank(ParallelMode.PIPELINE) if next_rank is None else next_rank
    # activations travel on the PIPELINE communicator, gradients on its twin (see ParallelContext.pipeline_bwd_group)
    fwd_group = gpc.get_group(ParallelMode.PIPELINE)
    bwd_group = gpc.pipeline_bwd_group if gpc.pipeline_bwd_group is not None else fwd_group
    fwd_ops, bwd_ops = [], []
    if object_send_next is not None:
        fwd_ops += _ops_for(process_object_to_send(object_send_next, scatter_gather_tensors), next_rank, True, fwd_group)
    if recv_prev_buf is not None:
        fwd_ops += _ops_for(recv_prev_buf, prev_rank, False, fwd_group)
    if object_send_prev is not None:
        bwd_ops += _ops_for(process_object_to_send(object_send_prev, scatter_gather_tensors), prev_rank, True, bwd_group)
    if recv_next_buf is not None:
        bwd_ops += _ops_for(recv_next_buf, next_rank, False, bwd_group)
    reqs = (dist.batch_isend_irecv(fwd_ops) if fwd_ops else []) + (dist.batch_isend_irecv(bwd_ops) if bwd_ops else [])

    def post(buf, split, shape):
        if buf is None:
            return None
        if split:
            if isinstance(buf, torch.Tensor):
                return _gather_1d(buf, shape)
            return [_gather_1d(b, s) for b, s in zip(buf, shape)]
        return buf

    def finish():
        for r in reqs:
            r.wait()
        return post(recv_prev_buf, rp_split, recv_prev_shape), post(recv_next_buf, rn_split, recv_next_shape)

    return finish


def _communicate(**kwargs):
    return _communicate_async(**kwargs)()


# --------------------------------------------------------------------------------------------------------------------
# synchronous verbs (stream-ordered; no host sync)
# --------------------------------------------------------------------------------------------------------------------
def recv_forward(input_tensor_shape, prev_rank=None, dtype=torch.float, scatter_gather_tensors=False):
    if gpc.is_pipeline_first_stage():
        return None
    return _communicate(recv_prev=True, recv_prev_shape=input_tensor_shape, prev_rank=prev_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)[0]


def recv_backward(output_grad_shape, next_rank=None, dtype=torch.float, scatter_gather_tensors=False):
    if gpc.is_pipeline_last_stage():
        return None
    return _communicate(recv_next=True, recv_next_shape=output_grad_shape, next_rank=next_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)[1]


def send_forward(output_tensor, next_rank=None, scatter_gather_tensors=False):
    if not gpc.is_pipeline_last_stage():
        _communicate(object_send_next=output_tensor, next_rank=next_rank, scatter_gather_tensors=scatter_gather_tensors)


def send_backward(input_tensor_grad, prev_rank=None, scatter_gather_tensors=False):
    if not gpc.is_pipeline_first_stage():
        _communicate(object_send_prev=input_tensor_grad, prev_rank=prev_rank,
                     scatter_gather_tensors=scatter_gather_tensors)


def send_forward_recv_backward(output_tensor, output_grad_shape, next_rank=None, dtype=torch.float,
                               scatter_gather_tensors=False):
    if gpc.is_pipeline_last_stage():
        return None
    return _communicate(object_send_next=output_tensor, recv_next=True, recv_next_shape=output_grad_shape,
                        next_rank=next_rank, dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)[1]


def send_backward_recv_forward(input_tensor_grad, input_tensor_shape, prev_rank=None, dtype=torch.float,
                               scatter_gather_tensors=False):
    if gpc.is_pipeline_first_stage():
        return None
    return _communicate(object_send_prev=input_tensor_grad, recv_prev=True, recv_prev_shape=input_tensor_shape,
                        prev_rank=prev_rank, dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)[0]


def send_forward_recv_forward(output_tensor, input_tensor_shape, prev_rank=None, next_rank=None, dtype=torch.float,
                              scatter_gather_tensors=False):
    return _communicate(object_send_next=output_tensor, recv_prev=input_tensor_shape is not None,
                        recv_prev_shape=input_tensor_shape, prev_rank=prev_rank, next_rank=next_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)[0]


def send_backward_recv_backward(input_tensor_grad, output_grad_shape, prev_rank=None, next_rank=None,
                                dtype=torch.float, scatter_gather_tensors=False):
    return _communicate(object_send_prev=input_tensor_grad, recv_next=output_grad_shape is not None,
                        recv_next_shape=output_grad_shape, prev_rank=prev_rank, next_rank=next_rank, dtype=dtype,
                        scatter_gather_tensors=scatter_gather_tensors)[1]


def send_forward_backward_recv_forward_backward(output_tensor, input_tensor_grad, input_tensor_shape,
                                                output_grad_shape, prev_rank=None, next_rank=None, dtype=torch.float,
                                                scatter_gather_tensors=False):
    return _communicate(object_send_next=output_tensor, object_send_prev=input_tensor_grad,
                        recv_prev=input_tensor_shape is not None, recv_next=output_grad_shape is not None,
                        recv_prev_shape=input_tensor_shape, recv_next_shape=output_grad_shape, prev_rank=prev_rank,
                        next_rank=next_rank, dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)


class AsynCommunicator:
    """Start an exchange now, collect the received tensor later (reference ``p2p.py:549-584``)."""

    def __init__(self, tensor_to_send=None, recv_shape=None, dtype=None, scatter_gather_tensors=False, forward=True):
        self._finish = None
        self.forward = forward
        self.kw = dict(dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)
        self.tensor_to_send, self.recv_shape = tensor_to_send, recv_shape

    @property
    def need_receive(self) -> bool:
        """Does ``wait_and_receive`` hand back a tensor (a receive shape was given) or only complete the send?"""
        return self.recv_shape is not None

    def start(self) -> None:
        if self.forward:  # send to next stage, receive from previous
            self._finish = _communicate_async(object_send_next=self.tensor_to_send, recv_prev=self.recv_shape is not None,
                                              recv_prev_shape=self.recv_shape, **self.kw)
        else:
            self._finish = _communicate_async(object_send_prev=self.tensor_to_send, recv_next=self.recv_shape is not None,
                                              recv_next_shape=self.recv_shape, **self.kw)

    def wait_and_receive(self):
        prev, nxt = self._finish()
        return prev if self.forward else nxt


# --------------------------------------------------------------------------------------------------------------------
# two-phase coroutines of the reference's interleaved scheduler (``p2p.py:429-546``): ``next(co)`` starts the exchange,
# the second ``next(co)`` waits for it and returns the received tensor.  ``AsynCommunicator`` is the object form.
# --------------------------------------------------------------------------------------------------------------------
def send_forward_and_recv_next_forward_async(output_tensor, recv_prev_shape=None, dtype=None, scatter_gather_tensors=False):
    finish = _communicate_async(object_send_next=output_tensor, recv_prev=recv_prev_shape is not None,
                                recv_prev_shape=recv_prev_shape, dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)
    yield
    yield finish()[0]


def send_backward_and_recv_next_backward_async(input_tensor, recv_next_shape=None, dtype=None, scatter_gather_tensors=False):
    finish = _communicate_async(object_send_prev=input_tensor, recv_next=recv_next_shape is not None,
                                recv_next_shape=recv_next_shape, dtype=dtype, scatter_gather_tensors=scatter_gather_tensors)
    yield
    yield finish()[1]
