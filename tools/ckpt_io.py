"""Shared helpers for the checkpoint tools: read every ``model_tp{t}_pp{p}.pt`` of a training checkpoint folder and merge
them into ONE full state dict with global layer indices (pipeline stages are concatenated, tensor-parallel shards are
un-sharded with ``internevo_b200.models.sharding``)."""
import os
import re
import sys
from typing import Dict, List, Tuple

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from internevo_b200.models.sharding import _LINEAR_W, shard_state_dict, shard_state_dict_isp, unshard_tensors  # noqa: E402

_FN = re.compile(r"^model_tp(\d+)_pp(\d+)\.pt$")
_FN_ISP = re.compile(r"^model_tp(\d+)_wp(\d+)_pp(\d+)\.pt$")
_EMBED = ("tok_embeddings.weight", "embedding.weight")
_HEAD = ("output.weight", "head.weight")
_LAYER = re.compile(r"^(?:model\.)?(layers|blocks)\.(\d+)\.(.*)$")


def find_shards(folder: str) -> Tuple[int, int]:
    tp = pp = 0
    for fn in os.listdir(folder):
        m = _FN.match(fn)
        if m:
            tp, pp = max(tp, int(m.group(1)) + 1), max(pp, int(m.group(2)) + 1)
    assert tp and pp, f"no model_tp*_pp*.pt under {folder}"
    return tp, pp


def find_isp_shards(folder: str):
    """``{(tp, wp, pp)}`` of the weight-parallel (isp) files of a folder, empty for the other layouts."""
    found = set()
    for fn in os.listdir(folder):
        m = _FN_ISP.match(fn)
        if m:
            found.add(tuple(int(g) for g in m.groups()))
    return found


def merge_pp(folder: str, tp_rank: int, pp_size: int, file_fmt: str = "model_tp{tp}_pp{pp}.pt") -> Dict[str, torch.Tensor]:
    """Concatenate the pipeline stages of one tensor-parallel rank; layer indices become global."""
    out: Dict[str, torch.Tensor] = {}
    shift = 0
    for pp in range(pp_size):
        sd = torch.load(os.path.join(folder, file_fmt.format(tp=tp_rank, pp=pp)), map_location="cpu", weights_only=False)
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


def load_full_state_isp(folder: str) -> Dict[str, torch.Tensor]:
    """Merge ``model_tp{t}_wp{w}_pp{p}.pt`` files: every linear weight is the concatenation of its weight-parallel row shards,
    the embedding of the tensor ranks' hidden slices, the head of the tensor ranks' vocabulary rows, the rest is replicated."""
    shards = find_isp_shards(folder)
    tp = max(s[0] for s in shards) + 1
    wp = max(s[1] for s in shards) + 1
    pp = max(s[2] for s in shards) + 1
    by_w = {w: next(t for (t, w2, _) in sorted(shards) if w2 == w) for w in range(wp)}       # the tensor rank stored with wp rank w
    per_w = [merge_pp(folder, by_w[w], pp, "model_tp{tp}_wp" + str(w) + "_pp{pp}.pt") for w in range(wp)]
    first_with_t = {t: next(w for w in range(wp) if by_w[w] == t) for t in range(tp)}
    full = {}
    for k in per_w[0]:
        if any(k.endswith(e) for e in _EMBED):
            full[k] = torch.cat([per_w[first_with_t[t]][k] for t in range(tp)], 1)
        elif any(k.endswith(e) for e in _HEAD):
            full[k] = torch.cat([per_w[first_with_t[t]][k] for t in range(tp)], 0)
        elif any(k.endswith(e) for e in _LINEAR_W):
            full[k] = torch.cat([per_w[w][k] for w in range(wp)], 0)
        else:
            full[k] = per_w[0][k]
    return full


def save_sharded_isp(full: Dict[str, torch.Tensor], folder: str, tp_size: int, wp_size: int):
    """Write ``full`` as a single-stage weight-parallel checkpoint (one file per weight rank, tensor rank ``w % tp``)."""
    os.makedirs(folder, exist_ok=True)
    for w in range(wp_size):
        t = w % tp_size
        sd = shard_state_dict_isp({k: v for k, v in full.items() if not any(k.endswith(e) for e in _EMBED + _HEAD)}, w, wp_size)
        for k, v in full.items():
            if any(k.endswith(e) for e in _EMBED):
                sd[k] = v.chunk(tp_size, 1)[t].clone()
            elif any(k.endswith(e) for e in _HEAD):
                sd[k] = v.chunk(tp_size, 0)[t].clone()
        torch.save({k: v.clone() for k, v in sd.items()}, os.path.join(folder, f"model_tp{t}_wp{w}_pp0.pt"))


def load_full_state(folder: str, embed_split_hidden: bool = True) -> Dict[str, torch.Tensor]:
    if find_isp_shards(folder):
        return load_full_state_isp(folder)
    tp, pp = find_shards(folder)
    per_tp: List[Dict[str, torch.Tensor]] = [merge_pp(folder, t, pp) for t in range(tp)]
    return {k: unshard_tensors(k, [s[k] for s in per_tp], embed_split_hidden) for k in per_tp[0]}


def _names_of(full: Dict[str, torch.Tensor]) -> dict:
    """Layer-list / embedding / final-norm / head names of the family the keys belong to (InternLM v1 and MoE use ``blocks``)."""
    if any(k.startswith("blocks.") for k in full):
        return dict(layers_name="blocks", embed_name="embedding", final_norm="norm", head="head")
    return dict(layers_name="layers", embed_name="tok_embeddings", final_norm="norm", head="output")


def pipeline_stages(full: Dict[str, torch.Tensor], pp_size: int) -> List[Dict[str, torch.Tensor]]:
    """Cut a full state dict (global layer indices) into ``pp_size`` stages the way the trainer does (``partition_uniform``):
    local layer numbering per stage, embedding on the first stage, final norm + head on the last."""
    from internevo_b200.models.sharding import pipeline_slice
    from internevo_b200.solver.pipeline_utils import partition_uniform

    names = _names_of(full)
    n_layers = 1 + max(int(m.group(2)) for m in (_LAYER.match(k) for k in full) if m)
    parts = partition_uniform(n_layers, pp_size, 1)
    return [pipeline_slice(full, parts[p][0][0], parts[p][0][1], first=p == 0, last=p == pp_size - 1, **names)
            for p in range(pp_size)]


def save_sharded(full: Dict[str, torch.Tensor], folder: str, tp_size: int, embed_split_hidden: bool = True, pp_size: int = 1):
    """Write ``full`` as ``model_tp{t}_pp{p}.pt`` files of a ``tp_size`` x ``pp_size`` layout."""
    os.makedirs(folder, exist_ok=True)
    stages = pipeline_stages(full, pp_size) if pp_size > 1 else [full]
    for p, stage in enumerate(stages):
        for t in range(tp_size):
            sd = {k: v.clone() for k, v in shard_state_dict(stage, t, tp_size, embed_split_hidden).items()}
            torch.save(sd, os.path.join(folder, f"model_tp{t}_pp{p}.pt"))


def load_model_config(folder: str) -> dict:
    fp = os.path.join(folder, "model_config.pt")
    return dict(torch.load(fp, map_location="cpu", weights_only=False)) if os.path.exists(fp) else {}
