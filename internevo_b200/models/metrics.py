"""Accuracy / perplexity / per-dataset loss metrics (reference ``internlm/model/metrics.py:55-375``).

The reference re-reads the ``[T, V]`` logits on the hot path (argmax + its own softmax) and issues five tensor-parallel
all-reduces per micro-batch.  Here everything is derived from by-products of the loss kernel — the
per-token loss and the top-1 correctness flag — so the metric costs a few ``[T]``-sized ops and no communication until
``get_metric`` is called at log time.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.common import SchedulerHook, get_current_device
from internevo_b200.utils.megatron_timers import megatron_timer as timer


def _isp_shards() -> int:
    """Under ISP every rank of the tensor (sequence) group scores its own token shard; 1 otherwise."""
    from internevo_b200.utils.parallel import is_using_isp

    return gpc.get_world_size(ParallelMode.TENSOR) if gpc.config is not None and is_using_isp() else 1


def _dp_sum(t: torch.Tensor):
    """Sum the counters over every rank that saw DIFFERENT tokens: the data-parallel group, and under ISP the sequence shards."""
    if gpc.is_initialized(ParallelMode.DATA) and gpc.get_world_size(ParallelMode.DATA) > 1:
        dist.all_reduce(t, group=gpc.get_group(ParallelMode.DATA))
    if _isp_shards() > 1:
        dist.all_reduce(t, group=gpc.get_group(ParallelMode.TENSOR))
    return t


class AccPerplex:
    """Token accuracy + perplexity, overall and per dataset type id."""

    def __init__(self, device=None, tp_pg=None, dp_pg=None, tokenizer=None, dataset_types: List[str] = None):
        self.device = device or get_current_device()
        self.dataset_types = dataset_types
        self.total_type_count = len(dataset_types) if dataset_types else 0
        self.tp_pg, self.dp_pg = tp_pg, dp_pg
        self.type_ids = None
        self._reset()

    def _reset(self):
        dev = self.device
        self.right = torch.zeros(1, device=dev, dtype=torch.float64)
        self.total = torch.zeros(1, device=dev, dtype=torch.float64)
        self.total_log_probs = torch.zeros(1, device=dev, dtype=torch.float64)
        n = max(1, self.total_type_count)
        self.ds_right = torch.zeros(n, device=dev, dtype=torch.float64)
        self.ds_tokens = torch.zeros(n, device=dev, dtype=torch.float64)
        self.ds_loss = torch.zeros(n, device=dev, dtype=torch.float64)

    def set_current_type_ids(self, type_ids: torch.Tensor):
        self.type_ids = type_ids.to(self.device) if type_ids is not None else None

    def set_cu_seqlens(self, cu_seqlens):
        pass

    def __call__(self, logits, labels):
        return self.update(logits, labels)

    def update(self, logits=None, labels=None, per_token_loss=None, correct=None, type_ids=None):
        """Preferred: pass ``per_token_loss`` / ``correct`` from the loss kernel. With only ``logits`` falls back to an
        explicit computation (validation of un-fused heads)."""
        with torch.no_grad():
            labels = labels.reshape(-1)
            if per_token_loss is None:
                lf = logits.reshape(-1, logits.shape[-1]).float()
                if self.tp_pg is not None and dist.get_world_size(self.tp_pg) > 1:
                    parts = [torch.empty_like(lf) for _ in range(dist.get_world_size(self.tp_pg))]
                    dist.all_gather(parts, lf, group=self.tp_pg)
                    lf = torch.cat(parts, -1)
                per_token_loss = torch.nn.functional.cross_entropy(lf, labels, reduction="none", ignore_index=-100)
                correct = (lf.argmax(-1) == labels)
            mask = labels != -100
            self.right += (correct & mask).sum()
            self.total += mask.sum()
            self.total_log_probs += (per_token_loss * mask).sum()
            tids = type_ids if type_ids is not None else self.type_ids
            if self.total_type_count > 0 and tids is not None:
                tids = tids.reshape(-1)
                if _isp_shards() > 1 and tids.numel() == labels.numel() * _isp_shards():
                    tids = tids.chunk(_isp_shards())[gpc.get_local_rank(ParallelMode.TENSOR)]    # this rank's token shard
                tids = tids[: labels.numel()].long().clamp_(0, self.total_type_count - 1)
                self.ds_right.index_add_(0, tids, (correct & mask).double())
                self.ds_tokens.index_add_(0, tids, mask.double())
                self.ds_loss.index_add_(0, tids, (per_token_loss * mask).double())

    def get_metric(self, reset=True):
        stats = torch.cat([self.right, self.total, self.total_log_probs, self.ds_right, self.ds_tokens, self.ds_loss])
        _dp_sum(stats)
        n = max(1, self.total_type_count)
        right, total, logp = stats[0].item(), stats[1].item(), stats[2].item()
        acc = right / max(total, 1)
        loss = logp / max(total, 1)
        res = {"acc": round(acc, 4), "perplexity": round(float(torch.exp(torch.tensor(min(loss, 20.0)))), 4),
               "loss_from_metric": round(loss, 4)}      # the reference's AccPerplex carries its LossWithTypeId result along
        if self.total_type_count > 0:
            dr, dt, dl = stats[3:3 + n], stats[3 + n:3 + 2 * n], stats[3 + 2 * n:3 + 3 * n]
            for i, name in enumerate(self.dataset_types):
                t = dt[i].item()
                res[f"acc/{name}"] = round(dr[i].item() / t, 4) if t > 0 else 0
                res[f"tokens/{name}"] = t
                res[f"loss/{name}"] = round(dl[i].item() / t, 4) if t > 0 else 0
                res[f"perplexity/{name}"] = round(float(torch.exp(torch.tensor(min(dl[i].item() / t, 20.0)))), 4) if t > 0 else 0
        if reset:
            self._reset()
        return res


class LossWithTypeId:
    """Per-dataset-type loss for validation (reference ``metrics.py:246-337``)."""

    def __init__(self, device=None, dp_pg=None, dataset_types: List[str] = None) -> None:
        self.device = device or get_current_device()
        self.dataset_types = dataset_types
        self.total_type_count = len(dataset_types) if dataset_types else 0
        self.loss = torch.zeros(1, device=self.device, dtype=torch.float64)
        self.token_num = torch.zeros(1, device=self.device, dtype=torch.float64)
        if self.total_type_count:
            self.ds_loss = torch.zeros(self.total_type_count, device=self.device, dtype=torch.float64)
            self.ds_token_num = torch.zeros(self.total_type_count, device=self.device, dtype=torch.float64)

    def update(self, logits, labels, type_ids=None):
        from internevo_b200 import ops

        with torch.no_grad():
            labels = labels.reshape(-1)
            group = gpc.get_group(ParallelMode.TENSOR) if gpc.config.model.get("parallel_output", True) else None
            if _isp_shards() > 1:       # full-vocabulary logits of this rank's token shard
                r, group = gpc.get_local_rank(ParallelMode.TENSOR), None
                if labels.numel() == logits.reshape(-1, logits.shape[-1]).shape[0] * _isp_shards():
                    labels = labels.chunk(_isp_shards())[r]
                    type_ids = type_ids.reshape(-1).chunk(_isp_shards())[r] if type_ids is not None else None
            loss = ops.cross_entropy(logits.reshape(-1, logits.shape[-1]).detach().clone(), labels, process_group=group)
            mask = labels != -100
            self.loss += (loss * mask).sum()
            self.token_num += mask.sum()
            if self.total_type_count and type_ids is not None:
                t = type_ids.reshape(-1).long().clamp_(0, self.total_type_count - 1)
                self.ds_loss.index_add_(0, t, (loss * mask).double())
                self.ds_token_num.index_add_(0, t, mask.double())

    def get_metric(self, reset=True):
        stats = torch.cat([self.loss, self.token_num])
        _dp_sum(stats)
        res = {"loss_from_metric": (stats[0] / stats[1].clamp_min(1)).item()}
        if self.total_type_count:
            ds = torch.cat([self.ds_loss, self.ds_token_num])
            _dp_sum(ds)
            n = self.total_type_count
            for i, name in enumerate(self.dataset_types):
                res[f"loss/{name}"] = (ds[i] / ds[n + i]).item() if ds[n + i] > 0 else 0
        if reset:
            self.loss.zero_()
            self.token_num.zero_()
            if self.total_type_count:
                self.ds_loss.zero_()
                self.ds_token_num.zero_()
        return res


class SchedulerMetricHook(SchedulerHook):
    """Timers around fwd / loss / bwd + metric update after the criterion (reference ``metrics.py:340-375``)."""

    def __init__(self, metric=None, skip: bool = False, criterion=None) -> None:
        self._post_func = metric
        self._skip = skip
        self._criterion = criterion

    def bind_criterion(self, criterion):
        self._criterion = criterion

    def before_forward(self, scheduler, inputs) -> None:
        if not self._skip:
            timer("fwd").start()

    def after_forward(self, scheduler, outputs) -> None:
        if not self._skip:
            timer("fwd").stop()

    def before_criterion(self, scheduler, outputs, label) -> None:
        if not self._skip:
            timer("cal_loss").start()
        self._labels = label

    def after_criterion(self, scheduler, loss) -> None:
        if not self._skip:
            timer("cal_loss").stop()
        crit = self._criterion
        if self._post_func is not None and crit is not None and getattr(crit, "last_per_token_loss", None) is not None:
            self._post_func.update(labels=crit.last_labels, per_token_loss=crit.last_per_token_loss,
                                   correct=crit.last_correct)

    def before_backward(self, scheduler, outputs, outputs_grad) -> None:
        if not self._skip:
            timer("bwd").start()

    def after_backward(self, scheduler, inputs_grad) -> None:
        if not self._skip:
            timer("bwd").stop()

    def post_helper_func(self, scheduler, outputs, label) -> None:
        # metric update is driven from the criterion's by-products (see train.pipeline.get_scheduler_hooks)
        pass


def broadcast(src: torch.Tensor, other: torch.Tensor, dim: int) -> torch.Tensor:
    """Expand an index vector so it can address ``other`` along ``dim`` (reference ``metrics.py:17-29``)."""
    if dim < 0:
        dim += other.dim()
    if src.dim() == 1:
        src = src.view([-1 if i == dim else 1 for i in range(other.dim())])
    while src.dim() < other.dim():
        src = src.unsqueeze(-1)
    return src.expand(other.size())


def vanilla_scatter(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None, dim_size=None, reduce=None):
    """``scatter(..., reduce="sum")`` of torch_scatter in plain torch: per-dataset-type sums of the metric code
    (reference ``metrics.py:32-52``)."""
    index = broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        size[dim] = dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)
