"""Optimizer adapter for FSDP (ZeRO-3) wrapped models (reference ``internlm/solver/optimizer/fsdp_optimizer.py:21-235``).

With ``parallel.zero1.fsdp=True`` the model is wrapped by ``torch.distributed.fsdp`` (``use_orig_params=True``): after
backward every rank holds the reduce-scattered gradient of ITS slice of every parameter, as a view into FSDP's flat shard.
This adapter adds what the training loop expects from an optimizer: loss scaling with overflow skip, global gradient norm
and clipping across the shard group (and the tensor-parallel group), fp32 master weights and AdamW on the shards through
the same fused kernel as ``HybridZeroOptimizer`` (``ops.adamw_`` with device-resident clip scalars), and a state dict.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist

from internevo_b200 import ops
from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.logger import get_logger

from .utils import DynamicGradScaler

logger = get_logger(__file__)


class FSDPadaptOptimizer:
    def __init__(self, optimizer, grad_scal_cfg=None, zero_cfg=None):
        self.param_groups: List[dict] = optimizer.param_groups
        cfg = zero_cfg or {}
        self._clip_grad_norm = float(cfg.get("clip_grad_norm", 0.0))
        params = [p for g in self.param_groups for p in g["params"]]
        self._dtype = params[0].dtype if params else torch.float32
        sc = grad_scal_cfg or {}
        fp16 = dict(sc.get("fp16", {}) or {})
        self.grad_scaler = DynamicGradScaler(
            initial_scale=1.0 if self._dtype == torch.float32 else fp16.get("initial_scale", 2 ** 16),
            min_scale=fp16.get("min_scale", 1), growth_factor=sc.get("growth_factor", 2),
            backoff_factor=sc.get("backoff_factor", 0.5), growth_interval=fp16.get("growth_interval", 1000),
            hysteresis=sc.get("hysteresis", 2), max_scale=sc.get("max_scale", 2 ** 24))
        # fp32 master copy + Adam moments of the LOCAL shard of every parameter (empty shards are skipped)
        self._state: Dict[int, dict] = {}
        for g in self.param_groups:
            for p in g["params"]:
                if p.numel() == 0:
                    continue
                self._state[id(p)] = dict(master=p.detach().float().clone().view(-1),
                                          exp_avg=torch.zeros(p.numel(), dtype=torch.float32, device=p.device),
                                          exp_avg_sq=torch.zeros(p.numel(), dtype=torch.float32, device=p.device))
        self._step = 0
        dev = params[0].device if params else torch.device("cpu")
        self._scalars = torch.zeros(4, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------------------------------------------------
    @property
    def loss_scale(self):
        return self.grad_scaler.scale

    def backward(self, loss, retain_graph=False):
        (loss * self.loss_scale).backward(retain_graph=retain_graph)

    def backward_by_grad(self, tensor, grad):
        torch.autograd.backward(tensors=tensor, grad_tensors=grad)

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    def _global_sumsq(self) -> torch.Tensor:
        """Σ grad² over every shard of every rank: shards partition the parameters, so a plain sum over the ZERO1 group
        counts each element once; tensor-parallel replicas (norm weights) are counted on tp rank 0 only."""
        self._sumsq.zero_()
        tp_rank = gpc.get_local_rank(ParallelMode.TENSOR) if gpc.is_initialized(ParallelMode.TENSOR) else 0
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                if tp_rank != 0 and getattr(p, "is_tp_replica", False):
                    continue
                ops.sumsq_(p.grad, self._sumsq)
        if gpc.is_initialized(ParallelMode.ZERO1) and gpc.get_world_size(ParallelMode.ZERO1) > 1:
            dist.all_reduce(self._sumsq, group=gpc.get_group(ParallelMode.ZERO1))
        if gpc.is_initialized(ParallelMode.TENSOR) and gpc.get_world_size(ParallelMode.TENSOR) > 1:
            dist.all_reduce(self._sumsq, group=gpc.get_group(ParallelMode.TENSOR))
        return self._sumsq

    def step(self, closure=None):
        """→ ``(success, {group_name: grad_norm})`` like ``HybridZeroOptimizer.step``."""
        ops.clip_scalars_(self._global_sumsq(), self._scalars, float(self.loss_scale), self._clip_grad_norm)
        host = self._scalars.cpu()
        found_inf = bool(host[1] != 0)
        self.grad_scaler.update(found_inf)
        names = [g.get("name", f"group{i}") for i, g in enumerate(self.param_groups)]
        if found_inf:
            if gpc.is_rank_for_log():
                logger.warning("Overflow occurs, please check it.")
            self.zero_grad()
            return False, {n: -1.0 for n in names}
        self._step += 1
        for g in self.param_groups:
            beta1, beta2 = g.get("betas", (0.9, 0.999))
            for p in g["params"]:
                st = self._state.get(id(p))
                if st is None or p.grad is None:
                    continue
                lp = p.data.view(-1) if p.dtype != torch.float32 else None
                ops.adamw_(st["master"], st["exp_avg"], st["exp_avg_sq"], p.grad.contiguous().view(-1), lp, g["lr"], beta1,
                           beta2, g.get("eps", 1e-8), g.get("weight_decay", 0.0), self._step, self._scalars)
                if lp is None:
                    p.data.view(-1).copy_(st["master"])
        self.zero_grad()
        return True, {n: float(host[2]) for n in names}

    def clip_grad_norm(self, model, max_norm):
        pass  # done inside step()

    # ------------------------------------------------------------------------------------------------------------
    def state_dict(self):
        order = [p for g in self.param_groups for p in g["params"]]
        return {"grad_scaler": self.grad_scaler.state_dict(), "step": self._step,
                "shards": [{k: v.cpu() for k, v in self._state[id(p)].items()} if id(p) in self._state else None
                           for p in order],
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, states):
        self.grad_scaler.load_state_dict(states["grad_scaler"])
        self._step = states["step"]
        order = [p for g in self.param_groups for p in g["params"]]
        assert len(order) == len(states["shards"]), "optimizer state was saved for a different model / sharding"
        for p, sh in zip(order, states["shards"]):
            if sh is None:
                continue
            st = self._state[id(p)]
            for k in st:
                st[k].copy_(sh[k])
            p.data.view(-1).copy_(st["master"])
        for g, sg in zip(self.param_groups, states["param_groups"]):
            g.update(sg)
