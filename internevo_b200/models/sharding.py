"""Re-sharding of full (unsharded) state dicts into tensor- / weight- / pipeline-parallel shards and back.

The reference can only load checkpoints written with the *same* tp/pp sizes (asserts in
``internlm/checkpoint/components.py:146-158``); these helpers make the layout explicit so checkpoints can be
re-partitioned and the distributed tests can start every layout from identical weights.
Sharding rules (InternLM2 keys; the other families map onto the same kinds):
  * ``wqkv``  ``[(kv_head, q_per_kv + 2, d), h]``  → split by kv head (dim 0, whole groups)
  * ``wq/wk/wv``, ``Wqkv``, ``w1``, ``w3``, ``output`` → dim 0;  ``Wqkv`` is split per (q|k|v) block
  * ``wo``, ``out_proj``, ``w2``                      → dim 1
  * ``tok_embeddings`` / ``embedding``: dim 1 (``embed_split_hidden``) or dim 0 (vocab parallel)
  * norms, biases of row-parallel layers: replicated
"""
from __future__ import annotations

import re
from typing import Dict

import torch

_COL = ("wq.weight", "wk.weight", "wv.weight", "w1.weight", "w3.weight", "fc1.weight", "output.weight", "head.weight",
        "wq.bias", "wk.bias", "wv.bias", "fc1.bias")
_ROW = ("wo.weight", "out_proj.weight", "w2.weight", "fc2.weight")


def _chunk(t, n, r, dim):
    assert t.shape[dim] % n == 0, (t.shape, n, dim)
    return t.chunk(n, dim=dim)[r].contiguous()


def shard_tensor(name: str, t: torch.Tensor, rank: int, size: int, embed_split_hidden: bool = True) -> torch.Tensor:
    if size == 1:
        return t
    if name.endswith("wqkv.weight") or name.endswith("wqkv.bias"):
        return _chunk(t, size, rank, 0)  # (kv_head gs d) order: a dim-0 chunk is a set of whole kv groups
    if name.endswith("Wqkv.weight") or name.endswith("Wqkv.bias"):
        three = t.reshape(3, t.shape[0] // 3, *t.shape[1:])
        return _chunk(three, size, rank, 1).reshape(-1, *t.shape[1:])
    if any(name.endswith(k) for k in _COL):
        return _chunk(t, size, rank, 0)
    if any(name.endswith(k) for k in _ROW):
        return _chunk(t, size, rank, 1)
    if name.endswith("tok_embeddings.weight") or name.endswith("embedding.weight"):
        return _chunk(t, size, rank, 1 if embed_split_hidden else 0)
    return t


def shard_state_dict(full: Dict[str, torch.Tensor], rank: int, size: int, embed_split_hidden: bool = True):
    return {k: shard_tensor(k, v, rank, size, embed_split_hidden) for k, v in full.items()}


def unshard_tensors(name: str, parts, embed_split_hidden: bool = True) -> torch.Tensor:
    if len(parts) == 1:
        return parts[0]
    if name.endswith("Wqkv.weight") or name.endswith("Wqkv.bias"):
        threes = [p.reshape(3, p.shape[0] // 3, *p.shape[1:]) for p in parts]
        return torch.cat(threes, dim=1).reshape(-1, *parts[0].shape[1:])
    if name.endswith("wqkv.weight") or name.endswith("wqkv.bias") or any(name.endswith(k) for k in _COL):
        return torch.cat(parts, 0)
    if any(name.endswith(k) for k in _ROW):
        return torch.cat(parts, 1)
    if name.endswith("tok_embeddings.weight") or name.endswith("embedding.weight"):
        return torch.cat(parts, 1 if embed_split_hidden else 0)
    return parts[0]


_LINEAR_W = _COL + _ROW + ("wqkv.weight", "Wqkv.weight", "wqkv.bias", "Wqkv.bias")


def shard_state_dict_isp(full: Dict[str, torch.Tensor], rank: int, size: int):
    """ISP / weight parallel: EVERY linear weight (column- and row-parallel alike) is split along the output dimension
    over the WEIGHT group (``ISPLinear`` all-gathers it right before use); embeddings and norms are replicated."""
    if size == 1:
        return dict(full)
    return {k: (_chunk(v, size, rank, 0) if any(k.endswith(e) for e in _LINEAR_W) else v) for k, v in full.items()}


def pipeline_slice(full: Dict[str, torch.Tensor], start: int, end: int, first: bool, last: bool,
                   layers_name: str = "layers", embed_name: str = "tok_embeddings", final_norm: str = "norm",
                   head: str = "output"):
    """Keys of global layers ``[start, end)`` renumbered from 0 (checkpoints index layers per stage)."""
    out = {}
    pat = re.compile(rf"^{layers_name}\.(\d+)\.(.*)$")
    for k, v in full.items():
        m = pat.match(k)
        if m:
            i = int(m.group(1))
            if start <= i < end:
                out[f"{layers_name}.{i - start}.{m.group(2)}"] = v
        elif k.startswith(embed_name + "."):
            if first:
                out[k] = v
        elif k.startswith(final_norm + ".") or k.startswith(head + "."):
            if last:
                out[k] = v
        else:
            out[k] = v
    return out
