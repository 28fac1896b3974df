"""Shared helpers for the checkpoint tools: read every ``model_tp{t}_pp{p}.pt`` of a training checkpoint folder and merge
them into ONE full state dict with global layer indices (pipeline stages are concatenated, tensor-parallel shards are
un-sharded with ``internevo_b200.models.sharding``)."""
import os
import re
import sys
from typing import Dict, List, Tuple

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from internevo_b200.models.sharding import shard_state_dict, unshard_tensors  # noqa: E402

_FN = re.compile(r"^model_tp(\d+)_pp(\d+)\.pt$")
_LAYER = re.compile(r"^(?:model\.)?(layers|blocks)\.(\d+)\.(.*)$")


def find_shards(folder: str) -> Tuple[int, int]:
    tp = pp = 0
    for fn in os.listdir(folder):
        m = _FN.match(fn)
        if m:
            tp, pp = max(tp, int(m.group(1)) + 1), max(pp, int(m.group(2)) + 1)
    assert tp and pp, f"no model_tp*_pp*.pt under {folder}"
    return tp, pp


def merge_pp(folder: str, tp_rank: int, pp_size: int) -> Dict[str, torch.Tensor]:
    """Concatenate the pipeline stages of one tensor-parallel rank; layer indices become global."""
    out: Dict[str, torch.Tensor] = {}
    shift = 0
    for pp in range(pp_size):
        sd = torch.load(os.path.join(folder, f"model_tp{tp_rank}_pp{pp}.pt"), map_location="cpu", weights_only=False)
        local_max = -1
        for k, v in sd.items():
            if k.endswith("inv_freq") or not torch.is_tensor(v):
                continue
            k = k[6:] if k.startswith("model.") else k
            m = _LAYER.match(k)
            if m:
                idx = int(m.group(2))
                local_max = max(local_max, idx)
                k = f"{m.group(1)}.{idx + shift}.{m.group(3)}"
            out[k] = v
        shift += local_max + 1
    return out


def load_full_state(folder: str, embed_split_hidden: bool = True) -> Dict[str, torch.Tensor]:
    tp, pp = find_shards(folder)
    per_tp: List[Dict[str, torch.Tensor]] = [merge_pp(folder, t, pp) for t in range(tp)]
    return {k: unshard_tensors(k, [s[k] for s in per_tp], embed_split_hidden) for k in per_tp[0]}


def save_sharded(full: Dict[str, torch.Tensor], folder: str, tp_size: int, embed_split_hidden: bool = True):
    os.makedirs(folder, exist_ok=True)
    for t in range(tp_size):
        sd = {k: v.clone() for k, v in shard_state_dict(full, t, tp_size, embed_split_hidden).items()}
        torch.save(sd, os.path.join(folder, f"model_tp{t}_pp0.pt"))


def load_model_config(folder: str) -> dict:
    fp = os.path.join(folder, "model_config.pt")
    return dict(torch.load(fp, map_location="cpu", weights_only=False)) if os.path.exists(fp) else {}
