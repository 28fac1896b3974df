"""Functional form of the parallel linears: ``fused_dense_func`` & co. for code that calls the linear as a function on its own
weight tensors instead of instantiating ``ColumnParallelLinear`` / ``RowParallelLinear`` / ``ISPLinear``.

Parity surface: reference ``internlm/model/utils.py:31-69`` (``ReduceScatterFunc`` / ``AllReduceFunc``), ``:220-226``
(``linear_bias_wgrad_torch``), ``:228-346`` (``FusedDenseFunc``: all-gather(x) -> GEMM, the input is re-gathered in
backward), ``:349-463`` (``MegatronFusedDenseFunc``: the gathered input is kept), ``:466-586`` (``ISPFusedDenseFunc``) and
the three ``*_fused_dense_func`` entry points ``:589-660``.

Nothing is re-implemented: the three classes bind (kind, mode) of the ONE autograd function behind the modules
(``linear._ParallelLinearFn`` / ``linear._ISPLinearFn``), so the functional path runs the same tcgen05 GEMMs and - with a
peer-memory heap - the same in-kernel collectives as the module path.  The reference's trailing ``is_using_cuda`` flag is
accepted and ignored (there is one backend here).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import linear as _linear
from .functional import all_gather_raw, all_reduce_raw, reduce_scatter_raw


def _ws(group) -> int:
    return 1 if group is None else dist.get_world_size(group)


class ReduceScatterFunc(torch.autograd.Function):
    """Reduce-scatter along ``reduce_dim`` in forward, all-gather of the gradient along the same dim in backward."""

    @staticmethod
    def forward(ctx, input_: Tensor, process_group, reduce_dim: int = 0) -> Tensor:
        ctx.group, ctx.dim = process_group, reduce_dim
        if _ws(process_group) <= 1:
            return input_
        x = input_.movedim(reduce_dim, 0).contiguous() if reduce_dim != 0 else input_.contiguous()
        out, _ = reduce_scatter_raw(x, process_group)
        return out.movedim(0, reduce_dim) if reduce_dim != 0 else out

    @staticmethod
    def backward(ctx, grad_output: Tensor):
        if _ws(ctx.group) <= 1:
            return grad_output, None, None
        g = grad_output.movedim(ctx.dim, 0).contiguous() if ctx.dim != 0 else grad_output.contiguous()
        out, _ = all_gather_raw(g, ctx.group)
        return (out.movedim(0, ctx.dim) if ctx.dim != 0 else out), None, None


class AllReduceFunc(torch.autograd.Function):
    """All-reduce(SUM) in forward, identity in backward (output of a row-parallel linear)."""

    @staticmethod
    def forward(ctx, input_: Tensor, process_group) -> Tensor:
        if _ws(process_group) <= 1:
            return input_
        out, _ = all_reduce_raw(input_.contiguous(), process_group)
        return out

    @staticmethod
    def backward(ctx, grad_output: Tensor):
        return grad_output, None


reduce_scatter = ReduceScatterFunc.apply
all_reduce = AllReduceFunc.apply


def linear_bias_wgrad_torch(my_input: Tensor, grad_output: Tensor, has_d_bias: bool):
    """``(dW, db)`` of ``y = x W^T + b`` from the 2-D input and output gradient (the wgrad GEMM is the tcgen05 MN-major x
    MN-major kernel on bf16 CUDA tensors, ``torch.matmul`` otherwise)."""
    assert my_input.dtype == grad_output.dtype
    if grad_output.is_cuda and grad_output.dtype == torch.bfloat16:
        from internevo_b200 import ops

        grad_weight = ops.matmul(grad_output.contiguous(), my_input.contiguous(), a_mn=True, b_mn=True)
    else:
        grad_weight = torch.matmul(grad_output.t(), my_input)
    grad_bias = grad_output.sum(dim=0) if has_d_bias else None
    return grad_weight, grad_bias


def _flatten_for(x: Tensor, gather_dim: int):
    """``x`` as 2-D ``[rows, features]`` with ``gather_dim`` outermost, plus what is needed to undo it on an output whose
    row count may have grown (gather) by the group size."""
    nd = x.dim()
    if nd == 2:
        assert gather_dim in (0, -2)
        return x, None
    gd = gather_dim % nd
    assert gd < nd - 1, "the feature dim cannot be the gather dim"
    xm = x.movedim(gd, 0) if gd != 0 else x
    return xm.reshape(-1, x.shape[-1]), (gd, tuple(xm.shape[:-1]))


def _restore(y2: Tensor, info, grew: int):
    if info is None:
        return y2
    gd, lead = info
    lead = (lead[0] * grew, *lead[1:])
    y = y2.reshape(*lead, y2.shape[-1])
    return y.movedim(0, gd) if gd != 0 else y


class _BoundDense:
    """Column-parallel linear as a function.  ``MODE`` is the tensor-parallel flavour the subclass stands for.  Not an
    ``autograd.Function`` itself: ``apply`` forwards to the one autograd function behind the modules."""

    MODE = "fsp"

    @classmethod
    def run(cls, x, weight, bias, return_residual, process_group, sequence_parallel, gather_dim):
        ws = _ws(process_group)
        mode = cls.MODE if (sequence_parallel and ws > 1) else "mtp"
        x2, info = _flatten_for(x, gather_dim)
        y2 = _linear._ParallelLinearFn.apply(x2.contiguous(), weight, bias, process_group, "column", mode)
        y = _restore(y2, info, ws if mode != "mtp" else 1)
        return (y, x) if return_residual else y

    # ``XxxFusedDenseFunc.apply(x, weight, bias, return_residual, process_group, sequence_parallel, gather_dim, is_using_cuda)``
    @classmethod
    def apply(cls, x, weight, bias=None, return_residual=False, process_group=None, sequence_parallel=True,
              gather_dim=0, is_using_cuda=True):  # noqa: ARG003  pylint: disable=arguments-differ
        return cls.run(x, weight, bias, return_residual, process_group, sequence_parallel, gather_dim)


class FusedDenseFunc(_BoundDense):
    """all-gather(x) -> GEMM; only the local shard of x is saved and it is gathered again for wgrad (flash-attn style,
    the reference's ``fsp``)."""

    MODE = "fsp"


class MegatronFusedDenseFunc(_BoundDense):
    """all-gather(x) -> GEMM with the gathered input kept for backward (Megatron style, ``msp``)."""

    MODE = "msp"


class ISPFusedDenseFunc:
    """Weight-parallel linear as a function: ``apply(x, weight, bias, module, communicator, return_residual, is_using_cuda)``."""

    @classmethod
    def apply(cls, x, weight, bias, module, communicator, return_residual=False, is_using_cuda=True):  # noqa: ARG003
        shape = x.shape
        y = _linear._ISPLinearFn.apply(x.reshape(-1, shape[-1]), weight, bias, module, communicator)
        y = y if len(shape) == 2 else y.reshape(*shape[:-1], y.shape[-1])
        return (y, x) if return_residual else y


def fused_dense_func(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, return_residual: bool = False,
                     process_group=None, sequence_parallel: bool = True, gather_dim: int = 0):
    return FusedDenseFunc.apply(x, weight, bias, return_residual, process_group, sequence_parallel, gather_dim)


def megatron_fused_dense_func(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, return_residual: bool = False,
                              process_group=None, sequence_parallel: bool = True, gather_dim: int = 0):
    return MegatronFusedDenseFunc.apply(x, weight, bias, return_residual, process_group, sequence_parallel, gather_dim)


def isp_fused_dense_func(x: Tensor, weight: Tensor, module, communicator, bias: Optional[Tensor] = None,
                         return_residual: bool = False):
    return ISPFusedDenseFunc.apply(x, weight, bias, module, communicator, return_residual)


__all__ = ["ReduceScatterFunc", "AllReduceFunc", "reduce_scatter", "all_reduce", "linear_bias_wgrad_torch", "FusedDenseFunc",
           "MegatronFusedDenseFunc", "ISPFusedDenseFunc", "fused_dense_func", "megatron_fused_dense_func",
           "isp_fused_dense_func"]
