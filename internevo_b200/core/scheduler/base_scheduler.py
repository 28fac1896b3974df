"""Scheduler base class (reference ``internlm/core/scheduler/base_scheduler.py``)."""
from __future__ import annotations

import inspect
from abc import ABC, abstractmethod
from typing import Any, Callable, Dict, Iterable

import torch

_ACCEPTS_HINT: Dict[type, bool] = {}


def model_kwargs(target, data: Dict) -> Dict:
    """``data`` as keyword arguments for ``target`` (an Engine, a NaiveAMPModel or a bare module).  ``max_seqlen`` is a
    host-side hint this framework adds to packed batches (it saves the decoders a device sync); a user model whose
    ``forward`` neither names it nor takes ``**kwargs`` (e.g. the toy models of the reference's tests) does not get it."""
    if "max_seqlen" not in data:
        return data
    m = target
    while isinstance(getattr(m, "model", None), (torch.nn.Module, torch.nn.ModuleList)):
        m = m.model
    if isinstance(m, torch.nn.ModuleList) and len(m):
        m = m[0]
    fwd = getattr(m, "forward", None)
    ok = _ACCEPTS_HINT.get(type(m))
    if ok is None:
        try:
            params = inspect.signature(fwd).parameters.values()
            ok = any(p.kind is inspect.Parameter.VAR_KEYWORD or p.name == "max_seqlen" for p in params)
        except (TypeError, ValueError):
            ok = True
        _ACCEPTS_HINT[type(m)] = ok
    return data if ok else {k: v for k, v in data.items() if k != "max_seqlen"}


class BaseScheduler(ABC):
    def __init__(self, data_process_func: Callable = None):
        self.data_process_func = data_process_func
        # labels are un-packed with the ignore index as padding when the function can take one (``data.unpack_data``)
        self._label_pad = {}
        if data_process_func is not None:
            import inspect

            try:
                if "padding_v" in inspect.signature(data_process_func).parameters:
                    self._label_pad = {"padding_v": -100}
            except (TypeError, ValueError):
                pass

    @abstractmethod
    def pre_processing(self, engine):
        """actions before running the schedule"""

    def _load_micro_batch(self, data: Dict, label: torch.Tensor, offset: int, bsz_stride: int):
        """Slice rows ``[offset, offset + bsz_stride)`` out of every batch field (packed data: one row per micro-batch)."""
        assert isinstance(data, dict) and isinstance(label, torch.Tensor)
        micro = {k: v[offset: offset + bsz_stride] for k, v in data.items()}
        return micro, label[offset: offset + bsz_stride]

    @abstractmethod
    def forward_backward_step(self, engine, data_iter: Iterable, forward_only: bool, return_loss: bool = True,
                              return_output_label: bool = True):
        """one full batch: forward (+ backward) over all micro-batches"""

    @staticmethod
    def _call_engine(engine, inputs: Any):
        if isinstance(inputs, torch.Tensor):
            return engine(inputs)
        if isinstance(inputs, (list, tuple)):
            return engine(*inputs)
        if isinstance(inputs, dict):
            return engine(**model_kwargs(engine, inputs))
        raise TypeError(f"Expected engine inputs to be tensor, list, tuple or dict, got {type(inputs)}")

    @staticmethod
    def _call_engine_criterion(engine, outputs: Any, labels: Any):
        assert isinstance(outputs, (torch.Tensor, list, tuple, dict)), f"bad model output type {type(outputs)}"
        if isinstance(outputs, torch.Tensor):
            outputs = (outputs,)
        if isinstance(labels, torch.Tensor):
            labels = (labels,)
        if isinstance(outputs, (tuple, list)) and isinstance(labels, (tuple, list)):
            return engine.criterion(*outputs, *labels)
        if isinstance(outputs, (tuple, list)) and isinstance(labels, dict):
            return engine.criterion(*outputs, **labels)
        if isinstance(outputs, dict) and isinstance(labels, dict):
            return engine.criterion(**outputs, **labels)
        raise TypeError(f"unsupported (outputs, labels) types: {type(outputs)}, {type(labels)}")


def attach_host_max_seqlen(data: Dict) -> Dict:
    """Compute each micro-batch's longest segment on the HOST copy of ``cu_seqlens`` (before the H2D copy) so the model
    never calls ``.item()`` on a device tensor — the reference syncs once per forward
    (``internlm/model/modeling_internlm2.py:989``)."""
    cu = data.get("cu_seqlens", None)
    if cu is None or "max_seqlen" in data:
        return data
    if isinstance(cu, torch.Tensor) and cu.dim() == 2 and not cu.is_cuda:
        data["max_seqlen"] = (cu[:, 1:] - cu[:, :-1]).max(dim=1).values.to(torch.int32)
    elif isinstance(cu, (list, tuple)):
        data["max_seqlen"] = torch.tensor([int((c[1:] - c[:-1]).max()) for c in cu], dtype=torch.int32)
    return data
