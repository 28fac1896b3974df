"""Engine: owns model / optimizer / criterion / schedulers and defines what one optimizer step is (reference
``internlm/core/engine.py:19-195``)."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.nn import Module
from torch.nn.modules.loss import _Loss

from internevo_b200.core.gradient_handler import BaseGradientHandler
from internevo_b200.utils.common import get_batch_size, move_to_device
from internevo_b200.utils.nvtx import nvtx_range


class Engine:
    def __init__(self, model: Module, optimizer, lr_scheduler=None, beta2_scheduler=None, criterion: Optional[_Loss] = None,
                 gradient_handlers: Optional[List[BaseGradientHandler]] = None, clip_grad_norm: float = 0.0):
        self._model = model
        self._optimizer = optimizer
        self._lr_scheduler = lr_scheduler
        self._beta2_scheduler = beta2_scheduler
        self._criterion = criterion
        self._clip_grad_norm = clip_grad_norm
        self.training = True
        self._gradient_handlers = gradient_handlers or []

    @property
    def model(self):
        return self._model

    @property
    def optimizer(self):
        return self._optimizer

    @property
    def criterion(self):
        return self._criterion

    def _all_reduce_gradients(self):
        for handler in self._gradient_handlers:
            handler.handle_gradient()

    def zero_grad(self):
        self.optimizer.zero_grad()

    def step(self):
        """gradient handlers → optimizer.step → (on success) lr / beta2 schedulers. Returns ``(success, grad_norms)``."""
        with nvtx_range("engine.step"):
            self._all_reduce_gradients()
            self.optimizer.clip_grad_norm(self.model, self._clip_grad_norm)
            success, group_norms = self.optimizer.step()
        if success and self._lr_scheduler is not None:
            self._lr_scheduler.step()
        if success and self._beta2_scheduler is not None:
            self._beta2_scheduler.step()
        return success, group_norms

    def train(self):
        self.training = True
        self._model.train()

    def eval(self):
        self.training = False
        self._model.eval()

    def backward(self, loss: torch.Tensor):
        with nvtx_range("backward"):
            return self.optimizer.backward(loss)

    def backward_by_grad(self, tensor, grad):
        with nvtx_range("backward"):
            return self.optimizer.backward_by_grad(tensor, grad)

    def __call__(self, *args, **kwargs):
        with nvtx_range("forward"):
            return self.model(*args, **kwargs)

    def load_batch(self, data_iter, to_gpu=True):
        """→ ``(batch_data, batch_size)`` moved to the device (pinned → async H2D)."""
        if data_iter is None:
            raise RuntimeError("Dataloader is not defined.")
        try:
            batch_data = next(data_iter)
        except TypeError:
            batch_data = data_iter
        if to_gpu:
            batch_data = move_to_device(batch_data)
        return batch_data, get_batch_size(batch_data)
