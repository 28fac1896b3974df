"""Vocab-parallel cross entropy: one-pass online-LSE forward, in-place backward (the gradient overwrites the logits).

Replaces flash-attn's ``xentropy_cuda_lib`` + the ``all_gather(lse)`` / ``all_reduce(loss)`` glue (reference
``internlm/model/losses/ce_loss.py:10-58``, ``third_party/flash-attention/flash_attn/losses/cross_entropy.py``).
With tensor parallelism each rank holds ``V / tp`` columns; the per-row (max, sumexp, target-logit) triples are combined
with three tiny all-reduces on ``[T]`` vectors.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from .gemm import _bump


class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, smoothing, ignore_index, group, inplace_backward):
        # logits: [rows, V_local] bf16 ; labels: [rows] int64 (global vocabulary ids)
        rows, V = logits.shape
        world = dist.get_world_size(group) if group is not None else 1
        rank = dist.get_rank(group) if group is not None else 0
        vocab_start = rank * V
        total = V * world
        dev = logits.device
        stats = torch.empty(4, rows, device=dev, dtype=torch.float32)
        torch.ops.b200.ce_fwd(logits, labels, vocab_start, stats[0], stats[1], stats[2], stats[3])
        _bump()
        mx, se, sx, tg = stats[0], stats[1], stats[2], stats[3]
        if world > 1:
            gmx = mx.clone()
            dist.all_reduce(gmx, op=dist.ReduceOp.MAX, group=group)
            se = se * torch.exp(mx - gmx)
            packed = torch.stack([se, sx, tg])
            dist.all_reduce(packed, group=group)
            se, sx, tg = packed[0], packed[1], packed[2]
            mx = gmx
        lse = mx + torch.log(se)
        valid = labels != ignore_index
        loss = lse - tg
        if smoothing > 0:
            loss = (1 - smoothing) * loss + smoothing * (lse - sx / total)
        loss = torch.where(valid, loss, torch.zeros_like(loss))
        ctx.save_for_backward(logits, labels, lse)
        ctx.cfg = (vocab_start, smoothing, total, ignore_index, inplace_backward)
        correct = (tg >= mx) & valid  # the target holds the row maximum <=> greedy prediction is right
        ctx.mark_non_differentiable(correct)
        return loss, correct

    @staticmethod
    def backward(ctx, dloss, _dcorrect=None):
        logits, labels, lse = ctx.saved_tensors
        vocab_start, smoothing, total, ignore_index, inplace = ctx.cfg
        g = logits if inplace else logits.clone()
        torch.ops.b200.ce_bwd(g, labels, lse, dloss.float().contiguous(), vocab_start, smoothing, total, ignore_index)
        _bump()
        return g, None, None, None, None, None


def _ce_ref(logits, labels, smoothing, ignore_index, group):
    if group is not None and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        parts = [torch.empty_like(logits) for _ in range(world)]
        # differentiable all-gather along vocab for the reference path
        from internevo_b200.parallel.functional import gather_forward_split_backward

        logits = gather_forward_split_backward(logits, group, dim=-1)
    loss = torch.nn.functional.cross_entropy(
        logits.float(), labels, reduction="none", label_smoothing=smoothing, ignore_index=ignore_index
    )
    with torch.no_grad():
        correct = (logits.argmax(-1) == labels) & (labels != ignore_index)
    return loss, correct


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, label_smoothing: float = 0.0, ignore_index: int = -100,
                  process_group: Optional[dist.ProcessGroup] = None, inplace_backward: bool = True,
                  return_correct: bool = False):
    """Per-token loss ``[rows]`` (fp32). ``logits`` may be a vocab shard when ``process_group`` is given.
    With ``return_correct`` also returns the boolean top-1 correctness per token (free: the kernel already has the row
    maximum and the target logit), which replaces the reference's separate argmax pass in ``AccPerplex``."""
    logits2 = logits.reshape(-1, logits.shape[-1])
    labels1 = labels.reshape(-1)
    if _lib.use_native(logits2) and logits2.dtype == torch.bfloat16 and logits2.stride(-1) == 1 and logits2.stride(0) % 8 == 0:
        loss, correct = _CrossEntropyFn.apply(logits2, labels1.long().contiguous(), label_smoothing, ignore_index,
                                              process_group, inplace_backward)
    else:
        loss, correct = _ce_ref(logits2, labels1.long(), label_smoothing, ignore_index, process_group)
    return (loss, correct) if return_correct else loss
