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
        if state.get("finished") and "after_scheduler_dict" in state:
            # the reference's layout (its file, or one written here for it): the warm-up wrapper stops counting at the end of the
            # warm-up and the wrapped cosine scheduler counts from there (``lr_scheduler.py:62-72``)
            self.last_epoch = int(state["last_epoch"]) + int(state["after_scheduler_dict"]["last_epoch"])
        for g, lr in zip(self.optimizer.param_groups, self._last_lr):
            g["lr"] = lr

    def __str__(self):
        return json.dumps(self.state_dict(), indent=4, sort_keys=True)


class WarmupScheduler:
    """Linear warm-up for ``warmup_epochs`` steps, then whatever ``after_scheduler`` says (any object with
    ``get_lr() / step() / state_dict() / load_state_dict()`` and a ``base_lrs`` list; reference ``:10-75``)."""

    def __init__(self, optimizer, warmup_epochs: int, after_scheduler, last_epoch: int = -1):
        self.optimizer = optimizer
        self.warmup_epochs = int(warmup_epochs)
        self.after_scheduler = after_scheduler
        self.base_lrs = [g.setdefault("initial_lr", g["lr"]) for g in optimizer.param_groups]
        self.last_epoch = last_epoch
        self._last_lr = list(self.base_lrs)
        self.step()

    @property
    def finished(self) -> bool:
        return self.last_epoch >= self.warmup_epochs

    def get_lr(self):
        if self.finished:
            return self.after_scheduler.get_lr()
        return [(self.last_epoch + 1) / self.warmup_epochs * lr for lr in self.base_lrs]

    def get_last_lr(self):
        return self._last_lr

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        if self.finished:
            self.after_scheduler.base_lrs = self.base_lrs
            self.after_scheduler.step(self.last_epoch - self.warmup_epochs)
            self._last_lr = list(self.after_scheduler.get_last_lr())
            return
        self._last_lr = self.get_lr()
        for g, lr in zip(self.optimizer.param_groups, self._last_lr):
            g["lr"] = lr

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "warmup_epochs": self.warmup_epochs, "base_lrs": self.base_lrs,
                "_last_lr": self._last_lr, "after_scheduler_type": type(self.after_scheduler).__name__,
                "after_scheduler_dict": self.after_scheduler.state_dict()}

    def load_state_dict(self, state):
        assert state.get("after_scheduler_type", type(self.after_scheduler).__name__) == type(self.after_scheduler).__name__
        for k in ("last_epoch", "warmup_epochs", "base_lrs", "_last_lr"):
            if k in state:
                setattr(self, k, state[k])
        self.after_scheduler.load_state_dict(state["after_scheduler_dict"])
        for g, lr in zip(self.optimizer.param_groups, self._last_lr):
            g["lr"] = lr


class CosineAnnealingWarmupLR(FineTuneCosineAnnealingWarmupLR):
    """Warm-up given in steps instead of a ratio and no zero-lr prefix (reference ``:78-89``)."""

    def __init__(self, optimizer, total_steps: int, warmup_steps: int = 0, eta_min: float = 0.0, last_epoch: int = -1):
        super().__init__(optimizer, total_steps, init_steps=0, warmup_ratio=0.0, eta_min=eta_min, last_epoch=-2)
        self._warmup_steps = int(warmup_steps)
        self.warmup_epochs = self._warmup_steps
        self.last_epoch = last_epoch
        self.step()
