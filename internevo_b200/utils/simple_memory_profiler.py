"""Static + dynamic memory accounting (reference ``internlm/utils/simple_memory_profiler.py:205-675``).

* static: parameter / gradient / optimizer-state bytes as a tree that follows the module names (the optimizer states are read
  from the Hybrid-ZeRO arenas: fp32 master + two Adam moments of the shard THIS rank owns);
* dynamic: activation bytes per leaf module from forward hooks, one tree per model chunk (interleaved pipeline stages hold
  several chunks); a tensor hook on every output hands its bytes back when autograd reaches it, so the profiler also knows the
  LIVE activation bytes at every point of the step and their peak - the number that decides whether a micro-batch fits;
* allocator: allocated / reserved / peak bytes per profiled step.

``step()`` is called once per training step; after ``total_steps`` steps the text summary ``memory_{step}.log`` and a
self-contained ``memory.html`` (nested, proportional bars; a pyecharts sunburst as well when that package is importable) are
written to ``log_folder`` and every hook is removed.
"""
from __future__ import annotations

import html
import os
from collections import OrderedDict
from typing import Any, Dict, List, Tuple

import torch


def _nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


def _tensor_bytes(obj) -> int:
    if torch.is_tensor(obj):
        return _nbytes(obj)
    if isinstance(obj, (tuple, list)):
        return sum(_tensor_bytes(o) for o in obj)
    if isinstance(obj, dict):
        return sum(_tensor_bytes(o) for o in obj.values())
    return 0


class SimpleMemState:
    """A node of the module tree with its own and its children's bytes."""

    def __init__(self, layer_name: str, layer_mem: int = 0) -> None:
        self.layer_name = layer_name
        self._layer_mem = layer_mem
        self._total_mem = layer_mem
        self.sub_model_stats: "OrderedDict[str, SimpleMemState]" = OrderedDict()

    @property
    def layer_mem(self) -> int:
        return self._layer_mem

    @property
    def total_mem(self) -> int:
        return self._total_mem

    def add(self, path: str, mem: int) -> None:
        """Account ``mem`` bytes to the node ``path`` (dot separated, relative to this node), creating it on the way."""
        self._total_mem += mem
        if not path:
            self._layer_mem += mem
            return
        head, _, rest = path.partition(".")
        if head not in self.sub_model_stats:
            self.sub_model_stats[head] = SimpleMemState(head)
        self.sub_model_stats[head].add(rest, mem)

    def delete(self, path: str) -> int:
        """Remove the node ``path`` with everything below it; the bytes leave every ancestor's total.  Returns them."""
        head, _, rest = path.partition(".")
        child = self.sub_model_stats.get(head)
        if child is None:
            return 0
        if rest:
            freed = child.delete(rest)
        else:
            freed = child.total_mem
            del self.sub_model_stats[head]
        self._total_mem -= freed
        return freed

    def update_total_memory(self) -> int:
        """Recompute the totals bottom-up (after children were edited directly)."""
        self._total_mem = self._layer_mem + sum(c.update_total_memory() for c in self.sub_model_stats.values())
        return self._total_mem

    def find_layer_state(self, path: str, create: bool = False):
        node = self
        for part in filter(None, path.split(".")):
            if part not in node.sub_model_stats:
                if not create:
                    return None
                node.sub_model_stats[part] = SimpleMemState(part)
            node = node.sub_model_stats[part]
        return node

    def dump(self, prefix: str = "", depth: int = 3) -> str:
        s = f"{prefix}{self.layer_name}: {self._total_mem / 2**20:.2f} MB\n"
        if depth > 0:
            for c in self.sub_model_stats.values():
                s += c.dump(prefix + "  ", depth - 1)
        return s

    def to_json(self, base: int = 1) -> Dict[str, Any]:
        return {"name": self.layer_name, "value": self._total_mem / base if base != 1 else self._total_mem,
                "children": [c.to_json(base) for c in self.sub_model_stats.values()]}


class ActivationMemState:
    """One activation tree per model chunk (reference ``:172-191``) plus the live / peak counters of the current step."""

    def __init__(self, num_chunks: int) -> None:
        self._num_chunks = num_chunks
        self.inited: List[bool] = [False] * num_chunks
        self.states: List[SimpleMemState] = [SimpleMemState(f"activations_{i}" if num_chunks > 1 else "activations")
                                             for i in range(num_chunks)]
        self.live = 0
        self.peak_live = 0

    @property
    def total_mem(self) -> int:
        return sum(s.total_mem for s in self.states)

    def produced(self, chunk: int, path: str, mem: int) -> None:
        self.inited[chunk] = True
        self.states[chunk].add(path, mem)
        self.live += mem
        self.peak_live = max(self.peak_live, self.live)

    def released(self, mem: int) -> None:
        self.live = max(0, self.live - mem)

    def dump(self, prefix: str = "") -> str:
        text = "".join(s.dump(prefix) for s in self.states)
        return text + f"{prefix}activations alive at the peak of the step: {self.peak_live / 2**20:.2f} MB\n"

    def to_json(self, base: int = 1) -> List:
        return [s.to_json(base) for s in self.states]


def _unpack_chunks(model) -> Tuple[List[torch.nn.Module], int]:
    """Model chunks without their mixed-precision wrapper."""
    chunks = list(model) if isinstance(model, (torch.nn.ModuleList, list, tuple)) else [model]
    chunks = [c.model if isinstance(getattr(c, "model", None), torch.nn.Module) else c for c in chunks]
    return chunks, len(chunks)


def _first_grad_tensor(obj):
    if torch.is_tensor(obj):
        return obj if obj.requires_grad else None
    if isinstance(obj, (tuple, list)):
        for o in obj:
            t = _first_grad_tensor(o)
            if t is not None:
                return t
    return None


def _html_tree(node: Dict[str, Any], total: float) -> str:
    frac = 100.0 * node["value"] / total if total else 0.0
    label = f'{html.escape(str(node["name"]))} &mdash; {node["value"] / 2**20:.2f} MB ({frac:.1f} %)'
    bar = f'<div class="bar" style="width:{max(frac, 0.2):.2f}%"></div>'
    if not node["children"]:
        return f"<li>{label}{bar}</li>"
    inner = "".join(_html_tree(c, total) for c in node["children"])
    return f"<li><details><summary>{label}</summary>{bar}<ul>{inner}</ul></details></li>"


class SimpleMemoryProfiler:
    def __init__(self, model, optimizer, log_folder: str, total_steps: int = 5):
        self._chunks, n_chunks = _unpack_chunks(model)
        self._model = self._chunks[0]
        self._optimizer = optimizer
        self._log_folder = log_folder
        self._remaining_steps = total_steps
        self._stoped = False
        self._step = 0
        self._peaks: List[Tuple[int, int]] = []
        self._activation = ActivationMemState(n_chunks)
        self._hooks = []
        self._param = SimpleMemState("parameters")
        self._grad = SimpleMemState("gradients")
        self._os = SimpleMemState("optimizer_states")
        for ci, chunk in enumerate(self._chunks):
            pre = f"chunk{ci}." if n_chunks > 1 else ""
            for name, p in chunk.named_parameters():
                self._param.add(pre + name, _nbytes(p))
                self._grad.add(pre + name, _nbytes(p))
        for g in getattr(optimizer, "groups", []):
            for t in (g.master, g.exp_avg, g.exp_avg_sq):
                self._os.add(g.name, _nbytes(t))
        for ci, chunk in enumerate(self._chunks):
            for name, m in chunk.named_modules():
                if len(list(m.children())) == 0:
                    self._hooks.append(m.register_forward_hook(self._make_forward_hook(ci, name)))
        if torch.cuda.is_available():
            torch.cuda.reset_peak_memory_stats()

    # Forward of a leaf module: what it returns stays alive for backward.  The bytes are handed back when autograd delivers
    # the gradient of that output - a TENSOR hook, not a module backward hook: module hooks wrap the outputs in an identity
    # function whose results must not be modified in place, and the attention block rotates the projection output in place.
    def _make_forward_hook(self, chunk: int, name: str):
        def hook(module, inputs, output):
            if self._stoped or not torch.is_grad_enabled():
                return
            mem = _tensor_bytes(output)
            act = self._activation
            act.produced(chunk, name, mem)
            t = _first_grad_tensor(output)
            if t is not None:
                t.register_hook(lambda grad, act=act, mem=mem: act.released(mem))

        return hook

    def _summary(self) -> str:
        text = "".join(st.dump() for st in (self._param, self._grad, self._os)) + self._activation.dump()
        if torch.cuda.is_available():
            text += (f"allocator: allocated {torch.cuda.memory_allocated() / 2**30:.2f} GB, peak "
                     f"{torch.cuda.max_memory_allocated() / 2**30:.2f} GB, reserved "
                     f"{torch.cuda.memory_reserved() / 2**30:.2f} GB\n")
            for i, (alloc, peak) in enumerate(self._peaks):
                text += f"  step {i + 1}: allocated {alloc / 2**30:.2f} GB at the step end, peak {peak / 2**30:.2f} GB\n"
        return text

    def _render_html(self) -> None:
        data = [self._param.to_json(), self._grad.to_json(), self._os.to_json(), *self._activation.to_json()]
        total = float(sum(d["value"] for d in data)) or 1.0
        body = "".join(_html_tree(d, total) for d in data)
        page = ("<!doctype html><meta charset='utf-8'><title>memory</title><style>body{font:13px monospace}"
                "ul{list-style:none;padding-left:18px}.bar{height:6px;background:#4a90d9;margin:2px 0 6px}</style>"
                f"<h3>memory of rank {os.environ.get('RANK', '0')} after step {self._step}: {total / 2**30:.2f} GB accounted</h3>"
                f"<ul>{body}</ul>")
        with open(os.path.join(self._log_folder, "memory.html"), "w", encoding="utf-8") as f:
            f.write(page)
        try:  # the reference's sunburst, when its plotting package is around
            from pyecharts import options as opts
            from pyecharts.charts import Sunburst

            Sunburst().add("memory", data_pair=data).set_global_opts(title_opts=opts.TitleOpts(title="memory")).render(
                os.path.join(self._log_folder, "memory_sunburst.html"))
        except ImportError:
            pass

    def point(self, with_options: str = "", create_img: bool = False) -> None:
        """Write the current accounting (``with_options``: free-text tag appended to the file name)."""
        os.makedirs(self._log_folder, exist_ok=True)
        tag = f"_{with_options}" if with_options else ""
        with open(os.path.join(self._log_folder, f"memory_{self._step}{tag}.log"), "w", encoding="utf-8") as f:
            f.write(self._summary())
        if create_img:
            self._render_html()

    def step(self) -> None:
        if self._stoped:
            return
        self._step += 1
        self._remaining_steps -= 1
        if torch.cuda.is_available():
            self._peaks.append((torch.cuda.memory_allocated(), torch.cuda.max_memory_allocated()))
            torch.cuda.reset_peak_memory_stats()
        if self._remaining_steps <= 0:
            self.point(create_img=True)
            self._stoped = True
            for h in self._hooks:
                h.remove()
            self._hooks = []
        else:
            self._activation = ActivationMemState(len(self._chunks))
