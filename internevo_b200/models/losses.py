"""Language-model loss (reference ``internlm/model/losses/ce_loss.py:10-58``): shifted labels come from the data
pipeline; with ``parallel_output`` the logits are a vocabulary shard and the loss kernel combines the per-rank
(max, sum-exp, target) statistics over the tensor group."""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn

from internevo_b200 import ops
from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.parallel import is_using_isp


class FlashGPTLMLoss(nn.Module):
    def __init__(self, parallel_output=True, label_smoothing=0):
        super().__init__()
        self.label_smoothing = label_smoothing if label_smoothing is not None else 0
        self.parallel_output = parallel_output
        self.last_per_token_loss = None  # consumed by the metric hook: no second pass over the logits
        self.last_correct = None
        self.last_labels = None

    def forward(self, *args):
        if len(args) == 3:
            logits, _, labels = args
        elif len(args) == 2:
            logits, labels = args
        else:
            raise RuntimeError(f"The number of criterion inputs are:{len(args)}")
        shift_logits = logits.reshape(-1, logits.size(-1))
        shift_labels = labels.reshape(-1)
        isp = is_using_isp() and gpc.get_world_size(ParallelMode.TENSOR) > 1
        if isp:
            # ISP keeps activations sequence-sharded through the head: take the matching label slice
            n, r = gpc.get_world_size(ParallelMode.TENSOR), gpc.get_local_rank(ParallelMode.TENSOR)
            valid_in_micro_batch = (shift_labels != -100).sum().clamp_min(1)      # every rank holds the full labels
            shift_labels = shift_labels.chunk(n)[r]
            group = None
        else:
            group = gpc.get_group(ParallelMode.TENSOR) if self.parallel_output else None
        per_tok, correct = ops.cross_entropy(shift_logits, shift_labels, label_smoothing=self.label_smoothing,
                                             process_group=group, inplace_backward=True, return_correct=True)
        self.last_per_token_loss = per_tok.detach()
        self.last_correct = correct
        self.last_labels = shift_labels
        valid = (shift_labels != -100).sum().clamp_min(1)
        loss = per_tok.sum() / valid
        if isp:
            # The reference gathers the sequence in front of the head and takes the mean over ALL valid tokens of the micro-batch
            # (``ops/linear.py:65-82``, ``model/utils.py:258-269``).  Here the head works on this rank's token shard, so the rank
            # contributes its share ``n * S_r / N`` (S_r: its summed token losses, N: valid tokens of the whole micro-batch): the
            # gradient reduction over the sequence group is a mean, and the mean of the shares is the reference's ``S / N`` -
            # shards with fewer valid labels no longer weigh their tokens higher.  The VALUE handed back is that global mean
            # (one scalar all-reduce), the gradient is the share's.
            loss = per_tok.sum() * (n / valid_in_micro_batch)
            with torch.no_grad():
                mean = loss.detach().clone()
                dist.all_reduce(mean, group=gpc.get_group(ParallelMode.TENSOR))
                mean /= n
            loss = loss + (mean - loss.detach())
        return loss
