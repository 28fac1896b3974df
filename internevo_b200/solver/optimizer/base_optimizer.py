"""Adapter that lets a plain ``torch.optim.Optimizer`` sit where ``Engine`` expects the framework's optimizer protocol
(``backward`` / ``backward_by_grad`` / ``clip_grad_norm`` next to ``step`` / ``zero_grad`` / ``state_dict``; reference
``internlm/solver/optimizer/base_optimizer.py:8-46``).  ``HybridZeroOptimizer`` and ``FSDPadaptOptimizer`` implement the
protocol on their own storage; everything a wrapped optimizer already provides is forwarded by ``__getattr__`` instead of
being re-declared method by method."""
from __future__ import annotations

import torch

_FORWARDED = ("param_groups", "defaults", "state", "add_param_group", "step", "zero_grad", "load_state_dict", "state_dict")


class BaseOptimizer:
    def __init__(self, optim: torch.optim.Optimizer):
        self.optim = optim

    def __getattr__(self, name):
        # only reached for attributes this adapter does not define itself
        if name in _FORWARDED:
            return getattr(self.__dict__["optim"], name)
        raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")

    # -- the part of the protocol a torch optimizer lacks ----------------------------------------------------------------
    def backward(self, loss, retain_graph: bool = False):
        loss.backward(retain_graph=retain_graph)

    def backward_by_grad(self, tensor, grad):
        torch.autograd.backward(tensors=tensor, grad_tensors=grad)

    def clip_grad_norm(self, *_, **__):
        """Clipping belongs to the concrete optimizer (``HybridZeroOptimizer.step``); nothing to do for a plain one."""
