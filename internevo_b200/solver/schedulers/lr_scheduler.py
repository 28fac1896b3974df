"""LR schedule of the reference in closed form (``internlm/solver/schedulers/lr_scheduler.py:92-140``): zero for
``init_steps``, linear warm-up for ``total_steps * warmup_ratio`` steps, then cosine annealing to ``eta_min``."""
from __future__ import annotations

import json
import math


class FineTuneCosineAnnealingWarmupLR:
    def __init__(self, optimizer, total_steps: int, init_steps: int = 0, warmup_ratio: float = 0.0,
                 eta_min: float = 0.0, last_epoch: int = -1):
        self.optimizer = optimizer
        self.total_steps = int(total_steps)
        self._init_steps = int(init_steps)
        self._warmup_steps = int(total_steps * warmup_ratio)
        self.warmup_epochs = self._warmup_steps + self._init_steps
        self.eta_min = eta_min
        self.base_lrs = [g.setdefault("initial_lr", g["lr"]) for g in optimizer.param_groups]
        self.last_epoch = last_epoch
        self._last_lr = list(self.base_lrs)
        self.step()

    def _lr_at(self, epoch: int, base: float) -> float:
        if epoch >= self.warmup_epochs:
            t_max = max(1, self.total_steps - self.warmup_epochs)
            t = epoch - self.warmup_epochs
            return self.eta_min + (base - self.eta_min) * (1 + math.cos(math.pi * t / t_max)) / 2
        if epoch >= self._init_steps:
            return (epoch + 1 - self._init_steps) / max(1, self._warmup_steps) * base
        return 0.0

    def get_lr(self):
        return [self._lr_at(self.last_epoch, b) for b in self.base_lrs]

    def get_last_lr(self):
        return self._last_lr

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        self._last_lr = self.get_lr()
        for g, lr in zip(self.optimizer.param_groups, self._last_lr):
            g["lr"] = lr

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "base_lrs": self.base_lrs, "total_steps": self.total_steps,
                "_init_steps": self._init_steps, "_warmup_steps": self._warmup_steps, "eta_min": self.eta_min,
                "_last_lr": self._last_lr}

    def load_state_dict(self, state):
        for k in ("last_epoch", "base_lrs", "_last_lr"):
            if k in state:
                setattr(self, k, state[k])
        for g, lr in zip(self.optimizer.param_groups, self._last_lr):
            g["lr"] = lr

    def __str__(self):
        return json.dumps(self.state_dict(), indent=4, sort_keys=True)
