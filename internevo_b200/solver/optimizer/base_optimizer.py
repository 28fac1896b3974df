"""Thin wrapper exposing a ``torch.optim.Optimizer`` behind the engine's optimizer protocol — ``backward`` /
``backward_by_grad`` / ``clip_grad_norm`` next to the usual ``step`` / ``zero_grad`` / ``state_dict``
(reference ``internlm/solver/optimizer/base_optimizer.py:8-46``).  ``HybridZeroOptimizer`` and ``FSDPadaptOptimizer``
implement the same protocol on their own storage; this class is for plugging a plain optimizer into ``Engine``."""
from __future__ import annotations

import torch


class BaseOptimizer:
    def __init__(self, optim: torch.optim.Optimizer):
        self.optim = optim

    @property
    def param_groups(self):
        return self.optim.param_groups

    @property
    def defaults(self):
        return self.optim.defaults

    def add_param_group(self, *args, **kwargs):
        return self.optim.add_param_group(*args, **kwargs)

    def step(self, *args, **kwargs):
        return self.optim.step(*args, **kwargs)

    def zero_grad(self, *args, **kwargs):
        self.optim.zero_grad(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.optim.load_state_dict(*args, **kwargs)

    def state_dict(self):
        return self.optim.state_dict()

    def backward(self, loss, retain_graph: bool = False):
        loss.backward(retain_graph=retain_graph)

    def backward_by_grad(self, tensor, grad):
        torch.autograd.backward(tensors=tensor, grad_tensors=grad)

    def clip_grad_norm(self):
        pass
