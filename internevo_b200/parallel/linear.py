"""Tensor / sequence / weight parallel linear layers.

One autograd function (``_ParallelLinearFn``) covers the four tensor-parallel modes of the reference
(``internlm/model/ops/linear.py:205-396`` + ``internlm/model/utils.py:228-586``):

* ``mtp``  Megatron TP: column = [identity | dgrad all-reduce], row = [all-reduce | identity]
* ``msp``  Megatron SP: column = [all-gather(x) | reduce-scatter(dgrad)], row = [reduce-scatter | all-gather(dy)];
           the gathered input is kept for backward
* ``fsp``  like msp but only the local shard is kept and re-gathered in backward
* ``isp``  weight parallel: the weight shard is all-gathered for the GEMM, its gradient reduce-scattered (AVG)

Every GEMM is the hand-written tcgen05 kernel (``ops.matmul``); with ``fused=True`` and a peer-memory heap the
collective runs *inside* the GEMM kernel (``parallel/fused.py``), otherwise NCCL is issued asynchronously and overlapped
with the neighbouring GEMM exactly where the reference overlaps it.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from internevo_b200 import ops
from internevo_b200.ops.gemm import wgrad as _wgrad

from .functional import all_gather_raw, all_reduce_raw, reduce_scatter_raw

_fused_backend = None  # set by parallel.fused.enable()


def set_fused_backend(backend):
    global _fused_backend
    _fused_backend = backend


def _ws(group):
    return 1 if group is None else dist.get_world_size(group)


def _mm(x2, w, bias=None):
    if x2.is_cuda and x2.dtype == torch.bfloat16 and w.dtype == torch.bfloat16:
        return ops.matmul(x2, w, bias=bias)
    return F.linear(x2, w, bias)


def _mm_dgrad(dy2, w):
    if dy2.is_cuda and dy2.dtype == torch.bfloat16 and w.dtype == torch.bfloat16:
        return ops.matmul(dy2, w, b_mn=True)
    return dy2 @ w


def _mm_wgrad(dy2, x2, weight):
    if dy2.is_cuda and dy2.dtype == torch.bfloat16 and x2.dtype == torch.bfloat16:
        return _wgrad(dy2, x2, weight)
    return (dy2.t() @ x2).to(weight.dtype)


class _ParallelLinearFn(torch.autograd.Function):
    """kind ∈ {"column", "row"}; mode ∈ {"mtp", "msp", "fsp"}.  Tensors are ``[tokens, features]`` (2-D)."""

    @staticmethod
    def forward(ctx, x, weight, bias, group, kind, mode):
        ctx.group, ctx.kind, ctx.mode = group, kind, mode
        ctx.has_bias = bias is not None
        ws = _ws(group)
        sp = mode in ("msp", "fsp") and ws > 1
        fused = _fused_backend if (_fused_backend is not None and ws > 1 and x.is_cuda) else None
        if kind == "column":
            if sp:
                if fused is not None and bias is None and fused.supports(x.shape[0] * ws, x, weight):
                    y, x_full = fused.ag_gemm(x, weight, group, keep_gathered=(mode == "msp"))
                else:
                    x_full, _ = all_gather_raw(x, group)
                    y = _mm(x_full, weight, bias)
                ctx.save_for_backward(x_full if mode == "msp" else x, weight)
            else:
                y = _mm(x, weight, bias)
                ctx.save_for_backward(x, weight)
        else:  # row
            if fused is not None and bias is None and ws > 1 and fused.supports(x.shape[0], x, weight):
                y = fused.gemm_rs(x, weight, group, all_reduce=not sp)
            else:
                y = _mm(x, weight, bias)
                if ws > 1:
                    if sp:
                        y, _ = reduce_scatter_raw(y, group)
                    else:
                        y, _ = all_reduce_raw(y, group)
            ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        group, kind, mode = ctx.group, ctx.kind, ctx.mode
        ws = _ws(group)
        sp = mode in ("msp", "fsp") and ws > 1
        dy = dy.contiguous()
        dx = dw = db = None
        fused = _fused_backend if (_fused_backend is not None and ws > 1 and dy.is_cuda) else None
        if kind == "column":
            handle_x = None
            if sp and mode == "fsp":
                x_full, handle_x = all_gather_raw(x, group, async_op=True)
            else:
                x_full = x
            handle = None
            if ctx.needs_input_grad[0]:
                if fused is not None and fused.supports(dy.shape[0], dy, weight):
                    # dgrad with the reduce-scatter / all-reduce done by the GEMM epilogue over NVLink
                    dx = fused.gemm_rs(dy, weight, group, all_reduce=not sp, b_mn=True)
                else:
                    dx = _mm_dgrad(dy, weight)
                    if ws > 1:
                        if sp:
                            dx, handle = reduce_scatter_raw(dx, group, async_op=True)
                        else:
                            dx, handle = all_reduce_raw(dx, group, async_op=True)
            if handle_x is not None:
                handle_x.wait()
            if ctx.needs_input_grad[1]:
                dw = _mm_wgrad(dy, x_full, weight)  # overlaps with the in-flight dgrad collective
            if ctx.has_bias:
                db = dy.float().sum(0).to(dy.dtype)
            if handle is not None:
                handle.wait()
        else:  # row
            if sp and fused is not None and ctx.needs_input_grad[0] and fused.supports(dy.shape[0] * ws, dy, weight):
                # all-gather(dy) pushed by copy CTAs while the dgrad GEMM runs; the gathered dy feeds wgrad right away
                dx, dy = fused.ag_gemm(dy, weight, group, b_mn=True)
            else:
                if sp:
                    dy, _ = all_gather_raw(dy, group)
                if ctx.needs_input_grad[0]:
                    dx = _mm_dgrad(dy, weight)
            if ctx.needs_input_grad[1]:
                dw = _mm_wgrad(dy, x, weight)
            if ctx.has_bias:
                db = dy.float().sum(0).to(dy.dtype)
        return dx, dw, db, None, None, None


def _isp_fused(x2, weight, module, use_fused_comm=None):
    """The peer-memory backend for this weight-parallel linear, or ``None`` (NCCL prefetch path)."""
    if not (x2.is_cuda and x2.dtype == torch.bfloat16):
        return None
    from internevo_b200.core.context import global_context as gpc

    if gpc.config is None or not gpc.config.get("fused_comm", False):
        return None
    from . import fused

    be = fused.isp_backend(module.process_group)
    return be if be is not None and be.supports(x2, weight) else None


class _ISPLinearFn(torch.autograd.Function):
    """Weight-parallel linear: all-gather(W) → GEMM; backward: all-gather(W) → dgrad, wgrad → reduce-scatter(AVG).
    With a peer-memory heap the three collectives run inside the GEMM kernels (``fused.ISPFusedBackend``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, module, communicator):
        ctx.module, ctx.comm = module, communicator
        ctx.has_bias = bias is not None
        be = _isp_fused(x, weight, module)
        ctx.fused = be
        if be is not None:
            # the communicator does not prefetch (NCCL-gather) this module's weight where the GEMM gathers it itself
            module._b200_isp_fused = (True, be.prefer("dgrad"))      # (forward, backward)
            y = be.gather_gemm(x, weight)
            if bias is not None:
                y = y + communicator.all_gather_weight(module, bias, is_bias=True)
        else:
            w_full = communicator.all_gather_weight(module, weight, is_bias=False)
            b_full = communicator.all_gather_weight(module, bias, is_bias=True) if bias is not None else None
            y = _mm(x, w_full, b_full)
            communicator.release_weight(module)
        ctx.save_for_backward(x, weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        comm, module = ctx.comm, ctx.module
        dy = dy.contiguous()
        be = ctx.fused
        dw = db = None
        dx = None
        # the collective runs inside the GEMM where that form is the faster one at this group size (ISPFusedBackend.prefer);
        # otherwise the communicator's NCCL gather (prefetched) / reduce-scatter path is taken per operation
        if ctx.needs_input_grad[0]:
            if be is not None and be.prefer("dgrad"):
                dx = be.gather_gemm(dy, weight, b_mn=True)
            else:
                w_full = comm.all_gather_weight(module, weight, is_bias=False, backward=True)
                dx = _mm_dgrad(dy, w_full)
                comm.release_weight(module)
        if ctx.needs_input_grad[1]:
            if be is not None and be.prefer("wgrad"):
                buf = getattr(weight, "grad_buf", None)
                if buf is not None:
                    be.wgrad_rs(dy, x, buf, accumulate=getattr(weight, "grad_ready", False))
                    weight.grad_ready = True
                    hook = getattr(weight, "grad_hook", None)
                    if hook is not None:
                        hook(weight)
                else:
                    dw = torch.empty_like(weight)
                    be.wgrad_rs(dy, x, dw, accumulate=False)
            else:
                if dy.is_cuda and dy.dtype == torch.bfloat16:
                    dw_full = ops.matmul(dy, x, a_mn=True, b_mn=True)
                else:
                    dw_full = (dy.t() @ x).to(weight.dtype)
                dw = comm.reduce_scatter_grad(module, weight, dw_full)
        if ctx.has_bias:
            db_full = dy.float().sum(0).to(dy.dtype)
            db = comm.reduce_scatter_grad(module, bias, db_full, is_bias=True)
        return dx, dw, db, None, None


class ParallelLinearBase(nn.Module):
    def __init__(self, in_features, out_features, bias, device, dtype):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(out_features, device=device, dtype=dtype)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def extra_repr(self):
        return f"in={self.in_features}, out={self.out_features}, bias={self.bias is not None}"

    def _apply2d(self, fn, x, *a):
        shape = x.shape
        y = fn(x.reshape(-1, shape[-1]), *a)
        return y.reshape(*((-1,) if y.shape[0] != math.prod(shape[:-1]) else shape[:-1]), y.shape[-1]) \
            if len(shape) != 2 else y


class ColumnParallelLinear(ParallelLinearBase):
    """Output features sharded over ``process_group`` (weight ``[out/tp, in]``)."""

    def __init__(self, in_features, out_features, process_group=None, bias=True, sequence_parallel=False,
                 tp_mode: str = "mtp", multiple_of=1, device=None, dtype=None):
        ws = _ws(process_group)
        assert out_features % multiple_of == 0
        mult = out_features // multiple_of
        local_mult = mult // ws + int(dist.get_rank(process_group) < mult % ws if ws > 1 else 0)
        super().__init__(in_features, local_mult * multiple_of, bias, device, dtype)
        self.process_group, self.tp_mode = process_group, tp_mode
        self.sequence_parallel = sequence_parallel

    def forward(self, x, gather_dim=0):
        shape = x.shape
        y = _ParallelLinearFn.apply(x.reshape(-1, shape[-1]), self.weight, self.bias, self.process_group, "column",
                                    self.tp_mode)
        return y if len(shape) == 2 else y.reshape(*shape[:-2], -1, y.shape[-1])


class RowParallelLinear(ParallelLinearBase):
    """Input features sharded (weight ``[out, in/tp]``); bias only applied on rank 0 (reference ``linear.py:321``)."""

    def __init__(self, in_features, out_features, process_group=None, bias=True, sequence_parallel=False,
                 tp_mode: str = "mtp", multiple_of=1, device=None, dtype=None):
        ws = _ws(process_group)
        rank = dist.get_rank(process_group) if ws > 1 else 0
        assert in_features % multiple_of == 0
        mult = in_features // multiple_of
        local_mult = mult // ws + int(rank < mult % ws)
        super().__init__(local_mult * multiple_of, out_features, bias and rank == 0, device, dtype)
        self.process_group, self.tp_mode = process_group, tp_mode
        self.sequence_parallel = sequence_parallel

    def forward(self, x):
        shape = x.shape
        y = _ParallelLinearFn.apply(x.reshape(-1, shape[-1]), self.weight, self.bias, self.process_group, "row",
                                    self.tp_mode)
        return y if len(shape) == 2 else y.reshape(*shape[:-2], -1, y.shape[-1])


class ISPLinear(ParallelLinearBase):
    """Weight-parallel linear (weight ``[out/wp, in]``, sharded over the WEIGHT group); activations are untouched."""

    __communicator = None

    @staticmethod
    def register_communicator(communicator):
        ISPLinear.__communicator = communicator

    @staticmethod
    def communicator():
        return ISPLinear.__communicator

    def __init__(self, in_features, out_features, process_group=None, bias=True, sequence_parallel=False,
                 tp_mode: str = "isp", multiple_of=1, device=None, dtype=None):
        ws = _ws(process_group)
        assert out_features % (ws * multiple_of) == 0, "ISP requires out_features divisible by the weight group size"
        super().__init__(in_features, out_features // ws, bias, device, dtype)
        self.full_out_features = out_features
        self.process_group = process_group

    def forward(self, x):
        shape = x.shape
        comm = ISPLinear.__communicator
        if comm is None or _ws(self.process_group) <= 1:
            y = ops.linear(x.reshape(-1, shape[-1]), self.weight, self.bias)
        else:
            y = _ISPLinearFn.apply(x.reshape(-1, shape[-1]), self.weight, self.bias, self, comm)
        return y if len(shape) == 2 else y.reshape(*shape[:-1], y.shape[-1])


# reference class names kept as aliases so user code written against InternEvo keeps importing
ColumnParallelLinearTorch = MegatronColumnParallelLinearTorch = ColumnParallelLinear
RowParallelLinearTorch = MegatronRowParallelLinearTorch = RowParallelLinear


def get_linear_cls(tp_mode: str, parallel_mode: str):
    """``get_linear_cls("fsp", "column")`` → class whose instances run in that mode (reference ``linear.py:381-396``)."""
    if tp_mode == "isp":
        return ISPLinear
    base = ColumnParallelLinear if parallel_mode == "column" else RowParallelLinear

    class _Bound(base):  # binds the mode so callers use the reference's (in, out, group, bias, ...) signature
        def __init__(self, *args, **kwargs):
            kwargs.setdefault("tp_mode", tp_mode)
            super().__init__(*args, **kwargs)

    _Bound.__name__ = f"{base.__name__}_{tp_mode}"
    return _Bound


class BaseScaleColumnParallelLinear(nn.Module):
    """Vocab-parallel LM head base: weight ``[vocab/tp, hidden]``, gradient-scaling trick
    ``w * s + (1 - s) * w.detach()`` (reference ``linear.py:24-83``)."""

    def __init__(self, in_features, out_features, process_group=None, bias=False, device=None, dtype=None,
                 weight_scale: float = 1.0, norm_head: bool = False):
        super().__init__()
        ws = _ws(process_group)
        assert out_features % ws == 0, f"out_features ({out_features}) must be divisible by world_size ({ws})"
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features // ws, in_features, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(out_features // ws, device=device, dtype=dtype)) if bias else None
        self.process_group = process_group
        self.weight_scale = weight_scale
        self.norm_head = norm_head
        self.first_eval_flag = True
        self.tmp_weight = None
        nn.init.normal_(self.weight, std=0.02)

    def _scaled_weight(self):
        if self.weight_scale != 1:
            return self.weight * self.weight_scale + (1 - self.weight_scale) * self.weight.detach()
        return self.weight

    def _norm_weight(self, weight):
        """L2-normalise each vocabulary row; cached once in eval mode (reference ``linear.py:129-143``)."""
        if self.training:
            if not self.first_eval_flag:
                self.first_eval_flag, self.tmp_weight = True, None
            return F.normalize(weight)
        if self.first_eval_flag:
            self.first_eval_flag = False
            self.tmp_weight = F.normalize(weight)
        return self.tmp_weight


class _GatherWeightFn(torch.autograd.Function):
    """Weight-parallel parameter → full weight: all-gather forward, reduce-scatter (AVG) of the gradient backward
    (every rank of the WEIGHT group saw different tokens, like data parallel ranks)."""

    @staticmethod
    def forward(ctx, w, group):
        ctx.group = group
        ws = _ws(group)
        out = torch.empty(ws * w.shape[0], *w.shape[1:], dtype=w.dtype, device=w.device)
        dist.all_gather_into_tensor(out, w.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        ws = _ws(ctx.group)
        out = torch.empty(g.shape[0] // ws, *g.shape[1:], dtype=g.dtype, device=g.device)
        dist.reduce_scatter_tensor(out, g.contiguous(), group=ctx.group)
        return out.div_(ws), None


class ScaleColumnParallelLinear(BaseScaleColumnParallelLinear):
    """LM head. ``forward(x, gather_dim, tp_mode)`` accepts sequence-sharded ``x`` under msp/fsp; under ``isp`` the
    vocabulary shards (WEIGHT group) are gathered and the logits are full-vocabulary for the local sequence shard."""

    def forward(self, x, gather_dim=0, tp_mode: str = "mtp"):
        weight = self._scaled_weight()
        if self.norm_head:
            weight = self._norm_weight(weight)
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if tp_mode == "isp":
            if _ws(self.process_group) > 1:
                weight = _GatherWeightFn.apply(weight, self.process_group)
            y = ops.linear(x2, weight, self.bias) if x2.is_cuda else F.linear(x2, weight, self.bias)
            return y if len(shape) == 2 else y.reshape(*shape[:-2], -1, y.shape[-1])
        mode = tp_mode if tp_mode in ("msp", "fsp") else "mtp"
        y = _ParallelLinearFn.apply(x2, weight, self.bias, self.process_group, "column", mode)
        return y if len(shape) == 2 else y.reshape(*shape[:-2], -1, y.shape[-1])


ScaleColumnParallelLinearWithNormHead = ScaleColumnParallelLinear


class RewardModelLinear(BaseScaleColumnParallelLinear):
    """Replicated (non-sharded) scalar/low-dim head; rank 0's init is broadcast (reference ``linear.py:156-202``)."""

    def __init__(self, in_features, out_features, process_group=None, bias=True, device=None, dtype=None,
                 weight_scale: float = 1.0):
        nn.Module.__init__(self)
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(out_features, device=device, dtype=dtype)) if bias else None
        nn.init.normal_(self.weight, std=0.02)
        self.process_group, self.weight_scale, self.norm_head = process_group, weight_scale, False
        if _ws(process_group) > 1:
            src = dist.get_global_rank(process_group, 0)
            dist.broadcast(self.weight.data, src=src, group=process_group)
            if bias:
                dist.broadcast(self.bias.data, src=src, group=process_group)

    def forward(self, x):
        weight = self._scaled_weight()
        from .functional import copy_to_group

        x = copy_to_group(x, self.process_group)
        return F.linear(x, weight, self.bias)
