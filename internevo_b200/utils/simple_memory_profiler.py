"""Static + dynamic memory accounting (reference ``internlm/utils/simple_memory_profiler.py:205-675``): parameter /
gradient / optimizer-state bytes by module tree, activation bytes by forward hooks, allocator peaks per step; dumps a
text summary (and a pyecharts sunburst when that package is importable) after ``stop_at`` steps."""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Any, Dict

import torch


def _nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


class SimpleMemState:
    """A node of the module tree with its own and its children's bytes."""

    def __init__(self, layer_name: str, layer_mem: int = 0) -> None:
        self.layer_name = layer_name
        self._layer_mem = layer_mem
        self._total_mem = layer_mem
        self.sub_model_stats: "OrderedDict[str, SimpleMemState]" = OrderedDict()

    @property
    def total_mem(self):
        return self._total_mem

    def add(self, path: str, mem: int):
        self._total_mem += mem
        if not path:
            self._layer_mem += mem
            return
        head, _, rest = path.partition(".")
        if head not in self.sub_model_stats:
            self.sub_model_stats[head] = SimpleMemState(head)
        self.sub_model_stats[head].add(rest, mem)

    def dump(self, prefix: str = "", depth: int = 3) -> str:
        s = f"{prefix}{self.layer_name}: {self._total_mem / 2**20:.2f} MB\n"
        if depth > 0:
            for c in self.sub_model_stats.values():
                s += c.dump(prefix + "  ", depth - 1)
        return s

    def to_json(self) -> Dict[str, Any]:
        return {"name": self.layer_name, "value": self._total_mem,
                "children": [c.to_json() for c in self.sub_model_stats.values()]}


class SimpleMemoryProfiler:
    def __init__(self, model: torch.nn.Module, optimizer, log_folder: str, total_steps: int = 5):
        self._model = model.model if hasattr(model, "model") else model
        self._optimizer = optimizer
        self._log_folder = log_folder
        self._remaining_steps = total_steps
        self._stoped = False
        self._step = 0
        self._peaks = []
        self._activation = SimpleMemState("activations")
        self._hooks = []
        self._param = SimpleMemState("parameters")
        self._grad = SimpleMemState("gradients")
        self._os = SimpleMemState("optimizer_states")
        for name, p in self._model.named_parameters():
            self._param.add(name, _nbytes(p))
            self._grad.add(name, _nbytes(p))
        for g in getattr(optimizer, "groups", []):
            for t in (g.master, g.exp_avg, g.exp_avg_sq):
                self._os.add(g.name, _nbytes(t))
        for name, m in self._model.named_modules():
            if len(list(m.children())) == 0:
                self._hooks.append(m.register_forward_hook(self._make_hook(name)))
        if torch.cuda.is_available():
            torch.cuda.reset_peak_memory_stats()

    def _make_hook(self, name):
        def hook(module, inputs, output):
            if self._stoped or not torch.is_grad_enabled():
                return
            outs = output if isinstance(output, (tuple, list)) else (output,)
            self._activation.add(name, sum(_nbytes(o) for o in outs if torch.is_tensor(o)))

        return hook

    def point(self, with_options: str = "", create_img: bool = False) -> None:
        os.makedirs(self._log_folder, exist_ok=True)
        with open(os.path.join(self._log_folder, f"memory_{self._step}.log"), "w", encoding="utf-8") as f:
            for st in (self._param, self._grad, self._os, self._activation):
                f.write(st.dump())
            if torch.cuda.is_available():
                f.write(f"allocator: allocated {torch.cuda.memory_allocated() / 2**30:.2f} GB, peak "
                        f"{torch.cuda.max_memory_allocated() / 2**30:.2f} GB, reserved "
                        f"{torch.cuda.memory_reserved() / 2**30:.2f} GB\n")
        if create_img:
            try:
                from pyecharts import options as opts
                from pyecharts.charts import Sunburst

                data = [s.to_json() for s in (self._param, self._grad, self._os, self._activation)]
                Sunburst().add("memory", data_pair=data).set_global_opts(
                    title_opts=opts.TitleOpts(title="memory")).render(os.path.join(self._log_folder, "memory.html"))
            except ImportError:
                pass

    def step(self) -> None:
        if self._stoped:
            return
        self._step += 1
        self._remaining_steps -= 1
        if torch.cuda.is_available():
            self._peaks.append(torch.cuda.max_memory_allocated())
        if self._remaining_steps <= 0:
            self.point(create_img=True)
            self._stoped = True
            for h in self._hooks:
                h.remove()
        else:
            self._activation = SimpleMemState("activations")
