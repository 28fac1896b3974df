"""Optimizer state between layouts: this framework's range-interleaved arena shards, the reference's per-rank flat buffers, and
a layout-free per-parameter form in between.

Why: ``HybridZeroOptimizer`` shards the fp32 master weights and the Adam moments ELEMENT-wise (every arena range is cut into
``zero_size`` equal sub-slices), the reference shards them PARAMETER-wise (whole parameters, largest first, to the emptiest rank;
``internlm/solver/optimizer/hybrid_zero_optim.py:238-262``) and stores one flat buffer per (group, rank).  Both write
``optimizer_tp{t}_pp{p}_zo{z}.pt``.  With the per-parameter form in the middle

* a run of the reference resumes here WITH its optimizer state (``named_from_reference``; the reference's own loader refuses a
  changed data-parallel size, ``hybrid_zero_optim.py:900``),
* a run of this framework resumes with another ZeRO / data-parallel size or bucket size (``named_from_arena_files`` of the old
  files, ``arena_states_from_named`` for the new layout),
* and a run can hand its state back to the reference (``reference_file_from_named``; ``ckpt.optimizer_ckpt_format =
  "reference"``).

The per-parameter form is ``{"grad_scaler": ..., "groups": {name: {"step", "hyper", "params": {parameter name: {"master",
"exp_avg", "exp_avg_sq"}}}}}`` with parameter names as in ``model.named_parameters()``.  Tensors in the reference's files are
named by the reference's state-dict keys (``w1`` / ``w3`` instead of the fused ``w13``, ...): the model's own state-dict hooks
translate, applied to fp32 stand-ins of the parameters (``file_layout_from_params`` / ``params_from_file_layout``), so whatever
the model checkpoint does to a weight is done to its master copy and moments as well.
"""
from __future__ import annotations

import re
from collections import OrderedDict
from contextlib import contextmanager
from typing import Dict, List, Optional, Sequence

import torch

KINDS = ("master", "exp_avg", "exp_avg_sq")
_FILE_KEY = {"master": "flat_fp32_weights", "exp_avg": "exp_avg", "exp_avg_sq": "exp_avg_sq"}
_EXPERT = re.compile(r"\.moe_layer\.experts\.wrapped_experts\.\d+\.")


# ---- arena geometry (pure functions of the numbers stored in the files) ------------------------------------------------------
def assemble_arena(shards: Sequence[torch.Tensor], ranges: Sequence[Sequence[int]], total: int) -> torch.Tensor:
    """Full ``[total]`` arena from the compact shards of ranks ``0 .. W-1``: range ``(lo, hi)`` is cut into ``W`` sub-slices of
    ``n = (hi - lo) / W`` elements and rank ``r`` keeps sub-slice ``r`` at ``[lo / W, lo / W + n)`` of its compact buffer."""
    W = len(shards)
    out = torch.empty(total, dtype=torch.float32)
    for lo, hi in ranges:
        n = (hi - lo) // W
        for r, sh in enumerate(shards):
            out[lo + r * n: lo + (r + 1) * n] = sh[lo // W: lo // W + n]
    return out


def carve_arena(full: torch.Tensor, ranges: Sequence[Sequence[int]], world: int, rank: int) -> torch.Tensor:
    """Inverse of ``assemble_arena`` for one rank."""
    out = torch.empty(full.numel() // world, dtype=full.dtype)
    for lo, hi in ranges:
        n = (hi - lo) // world
        out[lo // world: lo // world + n] = full[lo + rank * n: lo + (rank + 1) * n]
    return out


# ---- this framework's files -> per-parameter form -----------------------------------------------------------------------------
def named_from_arena_files(files: Sequence[dict]) -> dict:
    """``files``: the loaded ``optimizer_tp{t}_pp{p}_zo{z}.pt`` of ALL ZeRO ranks ``z = 0 .. W-1`` of one (tp, pp) coordinate."""
    first = files[0]
    named = {"grad_scaler": first["grad_scaler"], "groups": OrderedDict()}
    for gi, st in enumerate(first["groups"]):
        plan = first["zero_devide_optim_plan"][st["name"]]
        names = plan.get("names")
        assert names is not None and all(n is not None for n in names), (
            "this optimizer checkpoint was written without parameter names (older version): it can only be loaded with the "
            "parallel sizes / reduce_bucket_size it was saved with")
        W = st["zero_size"]
        assert len(files) == W, f"group {st['name']}: {W} ZeRO shards expected, {len(files)} files given"
        by_rank = sorted((f["groups"][gi] for f in files), key=lambda s: s["zero_rank"])
        assert [s["zero_rank"] for s in by_rank] == list(range(W))
        full = {k: assemble_arena([s[_FILE_KEY[k]] for s in by_rank], st["ranges"], st["total"]) for k in KINDS}
        params = OrderedDict()
        for name, (shape, off) in zip(names, plan["offsets"]):
            n = 1
            for d in shape:
                n *= d
            params[name] = {k: full[k][off: off + n].view(*shape).clone() for k in KINDS}
        named["groups"][st["name"]] = {"step": st["step"], "hyper": dict(st.get("hyper", {})), "params": params}
    return named


def arena_states_from_named(named: dict, optim) -> dict:
    """→ the dict ``optim.load_state_dict`` takes, in ``optim``'s OWN current layout (this rank's sub-slices)."""
    states = optim.state_dict()          # geometry + hyper-parameters of the live optimizer; tensors are replaced below
    states["grad_scaler"] = named["grad_scaler"]
    for g, st in zip(optim.groups, states["groups"]):
        src = named["groups"].get(g.name)
        assert src is not None, f"parameter group '{g.name}' is not in the optimizer checkpoint ({list(named['groups'])})"
        full = {k: torch.zeros(g.total, dtype=torch.float32) for k in KINDS}
        # alignment gaps of the master arena keep the live low-precision values (zeros) - they are never read
        for p in g.ordered:
            name = optim.param_name(p)
            assert name in src["params"], f"parameter '{name}' (group '{g.name}') is not in the optimizer checkpoint"
            o, n = g.offsets[id(p)], p.numel()
            for k in KINDS:
                t = src["params"][name][k]
                assert tuple(t.shape) == tuple(p.shape), f"{name}: checkpoint shape {tuple(t.shape)} != {tuple(p.shape)}"
                full[k][o: o + n] = t.reshape(-1).float()
        extra = set(src["params"]) - {optim.param_name(p) for p in g.ordered}
        assert not extra, f"group '{g.name}': the checkpoint has parameters the model does not: {sorted(extra)[:4]}"
        for k in KINDS:
            st[_FILE_KEY[k]] = carve_arena(full[k], g.ranges, g.zero_size, g.zero_rank)
        st["step"] = src["step"]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in src.get("hyper", {}):
                st.setdefault("hyper", {})[k] = src["hyper"][k]
    return states


# ---- parameter tensors <-> the model checkpoint's key space --------------------------------------------------------------------
@contextmanager
def _stand_ins(model, tensors: Optional[Dict[str, torch.Tensor]]):
    """Swap every parameter's storage for an fp32 CPU stand-in (given, or zeros) while the block runs.  The arena views come
    back afterwards untouched - nothing is written through them."""
    saved = []
    try:
        for name, p in model.named_parameters():
            saved.append((p, p.data))
            p.data = (tensors[name].detach().to(torch.float32).cpu().reshape(p.shape).clone() if tensors is not None
                      else torch.zeros(p.shape, dtype=torch.float32))
        yield
    finally:
        for p, data in saved:
            p.data = data


def _file_layout_hooks():
    from internevo_b200.core.context import ParallelMode, global_context as gpc  # noqa: F401
    from internevo_b200.utils.parallel import is_using_isp

    if not is_using_isp():
        return (lambda model, st: st), (lambda model, st: st)
    from .components import _isp_states_from_file_layout, _isp_states_to_file_layout

    return _isp_states_to_file_layout, _isp_states_from_file_layout


def file_layout_from_params(model, tensors: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """``{parameter name: fp32 tensor}`` → ``{model-checkpoint key: fp32 tensor}`` through the model's state-dict hooks (and the
    ISP file layout), i.e. exactly the transformation ``save_model_checkpoint`` applies to the weights."""
    from .utils import get_shard_state_dict

    to_file, _ = _file_layout_hooks()
    with _stand_ins(model, tensors):
        states = to_file(model, get_shard_state_dict(model))
        return OrderedDict((k, v.detach().clone()) for k, v in states.items()
                           if torch.is_tensor(v) and v.is_floating_point() and not k.endswith("inv_freq"))


def params_from_file_layout(model, file_states: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of ``file_layout_from_params``: what ``load_model_checkpoint`` does to the weights, done to fp32 stand-ins."""
    from .utils import load_shard_state_dict

    _, from_file = _file_layout_hooks()
    with _stand_ins(model, None):
        missing, unexpected = load_shard_state_dict(model, from_file(model, dict(file_states)), strict=False)
        pnames = {n for n, _ in model.named_parameters()}
        missing = [k for k in missing if k in pnames]
        assert not missing and not unexpected, (f"optimizer state does not cover the model: missing {missing[:4]}, "
                                                f"unexpected {list(unexpected)[:4]}")
        return {n: p.data.clone() for n, p in model.named_parameters()}


# ---- the reference's files ---------------------------------------------------------------------------------------------------
def reference_partition(shapes: Sequence[Sequence[int]], world: int):
    """The reference's parameter-wise ZeRO split (``hybrid_zero_optim.py:238-262``): parameters sorted by element count
    (largest first, ties in registration order), each one to the rank holding the fewest elements so far (lowest rank on
    ties).  → ``(per-rank lists of indices into `shapes`, per-rank lists of plan ids "<sorted position>_<dim>_<dim>...")``."""
    numel = [int(torch.Size(s).numel()) for s in shapes]
    order = sorted(range(len(shapes)), key=lambda i: numel[i], reverse=True)        # stable, like the reference's sorted()
    load = [0] * world
    idx: List[List[int]] = [[] for _ in range(world)]
    ids: List[List[str]] = [[] for _ in range(world)]
    for pos, i in enumerate(order):
        r = load.index(min(load))
        idx[r].append(i)
        ids[r].append("_".join([str(pos)] + [str(d) for d in shapes[i]]))
        load[r] += numel[i]
    return idx, ids


def _reference_group_of(key: str, tensor: torch.Tensor, model_dtype, isp: bool, moe_group: Optional[str]) -> str:
    """Which optimizer group the reference puts a parameter in (``internlm/train/utils.py:31-79``)."""
    if isp and re.search(r"(^|\.)(tok_embeddings|embedding|output|head)\.", key):
        return "embed_head"
    if tensor.dtype == torch.float32:       # whatever the model dtype: an fp32 MODEL has all its parameters in this group
        return "fp32"
    if moe_group is not None and _EXPERT.search(key):
        return moe_group
    return "default"


def _group_keys(model_keys: "OrderedDict[str, torch.Tensor]", model_dtype, isp: bool, moe_group: Optional[str]):
    groups: "OrderedDict[str, List[str]]" = OrderedDict()
    for k, v in model_keys.items():
        if torch.is_tensor(v) and v.is_floating_point() and not k.endswith("inv_freq"):
            groups.setdefault(_reference_group_of(k, v, model_dtype, isp, moe_group), []).append(k)
    return groups


def named_from_reference(files_of_group, model_keys: "OrderedDict[str, torch.Tensor]", model, optim, model_dtype,
                         isp: bool = False, moe_group: Optional[str] = None) -> dict:
    """``files_of_group(group name) -> [loaded optimizer file of the group's ZeRO rank 0, 1, ...]``: the dense groups are split
    over the ZERO1 group (files ``..._zo0 .. zo{W-1}``), an expert group over the ranks that replicate THIS rank's experts.
    ``model_keys``: the model state dict of the same coordinate in the order of the reference's ``model.parameters()``, which
    the plan ids refer to.  Every plan id carries the parameter's shape, so a parameter order that does not match the files is
    an error, not a silent mix-up."""
    # a model file written by the reference is in registration order already; one written by this framework is put into it
    model_keys = OrderedDict((k, model_keys[k]) for k in _reference_order(list(model_keys)))
    first = files_of_group("default")[0]
    group_names = [pg.get("name", f"group{i}") for i, pg in enumerate(first["base_optim_states"]["param_groups"])]
    keys_of = _group_keys(model_keys, model_dtype, isp, moe_group)
    flat: Dict[str, Dict[str, torch.Tensor]] = {k: {} for k in KINDS}      # kind -> {reference key: tensor}
    meta = {}
    for gid, gname in enumerate(group_names):
        keys = keys_of.get(gname, [])
        if not keys:
            continue
        files = files_of_group(gname)
        shapes = [tuple(model_keys[k].shape) for k in keys]
        idx, ids = reference_partition(shapes, len(files))
        for r, f in enumerate(files):
            plan = f["zero_devide_optim_plan"][gid]
            assert len(plan) == len(files) and list(plan[r]) == ids[r], (
                f"group '{gname}', ZeRO rank {r} of {len(files)}: the checkpoint's plan {[list(x)[:2] for x in plan]} does not "
                f"match the model file's parameter order {[x[:2] for x in ids]}")
            if not idx[r]:
                continue
            pg = f["base_optim_states"]["param_groups"][gid]
            state = f["base_optim_states"]["state"][pg["params"][0]]
            bufs = {"master": f["flat_fp32_weights"][gid], "exp_avg": state["exp_avg"], "exp_avg_sq": state["exp_avg_sq"]}
            off = 0
            for i in idx[r]:
                n = int(torch.Size(shapes[i]).numel())
                for k in KINDS:
                    flat[k][keys[i]] = bufs[k].detach().reshape(-1)[off: off + n].view(*shapes[i]).float().cpu()
                off += n
            assert off == bufs["master"].numel(), (gname, r, off, bufs["master"].numel())
            meta[gname] = {"step": int(float(state["step"])), "hyper": {k: pg[k] for k in ("lr", "betas", "eps", "weight_decay")
                                                                         if k in pg}}
    per_kind = {k: params_from_file_layout(model, flat[k]) for k in KINDS}
    named = {"grad_scaler": first["grad_scaler"], "groups": OrderedDict()}
    for g in optim.groups:
        # step count and hyper-parameters are those of the reference group of the same name; where the two sides file a parameter
        # under different groups (an fp32 model: "fp32" there, "default" here) any group's will do - the reference steps and
        # schedules all its groups together
        src = meta.get(g.name) or (next(iter(meta.values())) if meta else None)
        assert src is not None or not g.params, f"parameter group '{g.name}' has no counterpart in the reference checkpoint"
        params = OrderedDict((optim.param_name(p), {k: per_kind[k][optim.param_name(p)] for k in KINDS}) for p in g.ordered)
        named["groups"][g.name] = {"step": src["step"] if src else 0, "hyper": src["hyper"] if src else {}, "params": params}
    return named


def reference_file_from_named(named: dict, model, model_dtype, layout: Dict[str, Sequence[int]], isp: bool = False,
                              moe_group: Optional[str] = None) -> dict:
    """→ THIS rank's reference-format optimizer state dict.  ``layout[group name] = (ZeRO world size, this rank's ZeRO rank)`` of
    the group (groups missing from it: ``layout["default"]``).  The parameter order inside a group is the reference's
    registration order: this framework's state-dict order with ``w1, w2, w3`` inside an MLP (the reference registers them in
    that order, ``internlm/model/modules/mlp.py:54-79``)."""
    tensors = {k: {} for k in KINDS}
    for g in named["groups"].values():
        for name, st in g["params"].items():
            for k in KINDS:
                tensors[k][name] = st[k]
    per_kind = {k: file_layout_from_params(model, tensors[k]) for k in KINDS}
    # the group of a key follows the dtype of the live parameter's checkpoint entry, as on the reference's side
    from .utils import get_shard_state_dict

    to_file, _ = _file_layout_hooks()
    live = to_file(model, get_shard_state_dict(model))
    keys_of = _group_keys(OrderedDict((k, live[k]) for k in _reference_order(list(per_kind["master"]))), model_dtype, isp,
                          moe_group)
    # group ids as the reference numbers them: default, [embed_head], fp32, [moe] - empty groups included
    order = ["default"] + (["embed_head"] if isp else []) + ["fp32"] + ([moe_group] if moe_group else [])
    out = {"grad_scaler": named["grad_scaler"], "flat_fp32_weights": {}, "zero_devide_optim_plan": [],
           "base_optim_states": {"state": {}, "param_groups": []}}
    packed = 0      # torch.optim packs parameter ids group after group; a reference group holds ONE flat tensor (or nothing)
    for gid, gname in enumerate(order):
        keys = keys_of.get(gname, [])
        world, rank = layout.get(gname, layout["default"])
        shapes = [tuple(per_kind["master"][k].shape) for k in keys]
        idx, ids = reference_partition(shapes, world)
        src = named["groups"].get(gname) or next(iter(named["groups"].values()), {"step": 0, "hyper": {}})
        out["zero_devide_optim_plan"].append([list(x) for x in ids])
        # torch's ``load_state_dict`` REPLACES the live group dicts by the saved ones, so every key the reference reads after its
        # constructor has to be here: ``name``, ``dtype`` (of the group's model parameters; ``None`` for an empty group,
        # ``hybrid_zero_optim.py:155,604``), ``moe``.  ``optimizer_mode`` (an enum of the reference's own module) is only read
        # while the optimizer is built and is left out.
        pg = {"name": gname, "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
              "differentiable": False, "fused": True, "decoupled_weight_decay": True,
              "dtype": live[keys[0]].dtype if keys else None, "params": [packed] if keys else []}
        if moe_group is not None and gname == moe_group:
            pg["moe"] = True
        pg.update({k: v for k, v in src.get("hyper", {}).items() if k in ("lr", "betas", "eps", "weight_decay", "initial_lr")})
        pg.setdefault("initial_lr", pg.get("lr"))
        out["base_optim_states"]["param_groups"].append(pg)
        if idx[rank]:
            cat = {k: torch.cat([per_kind[k][keys[i]].reshape(-1) for i in idx[rank]]) for k in KINDS}
            out["flat_fp32_weights"][gid] = cat["master"]
            out["base_optim_states"]["state"][packed] = {"step": torch.tensor(float(src["step"])), "exp_avg": cat["exp_avg"],
                                                         "exp_avg_sq": cat["exp_avg_sq"]}
        packed += 1 if keys else 0
    return out


def _reference_order(keys: List[str]) -> List[str]:
    """State-dict keys in the reference's registration order: ours, except ``w1, w2, w3`` inside one MLP."""
    rank = {"w1": 0, "w2": 1, "w3": 2}
    out, i = [], 0
    while i < len(keys):
        m = re.match(r"^(.*\.)(w[123])\.(weight|bias)$", keys[i])
        if m is None:
            out.append(keys[i])
            i += 1
            continue
        j = i
        while j < len(keys) and (mm := re.match(r"^(.*\.)(w[123])\.(weight|bias)$", keys[j])) and mm.group(1) == m.group(1):
            j += 1
        block = keys[i:j]
        out.extend(sorted(block, key=lambda k: (rank[re.match(r"^.*\.(w[123])\.(weight|bias)$", k).group(1)], block.index(k))))
        i = j
    return out


def reference_scheduler_state(scheduler, n_groups: int) -> dict:
    """``schedulder.pt`` the way the reference's ``FineTuneCosineAnnealingWarmupLR.state_dict()`` writes it
    (``solver/schedulers/lr_scheduler.py:28-36``): the warm-up wrapper's fields plus the wrapped torch ``CosineAnnealingLR``'s, with
    one ``base_lr`` per REFERENCE parameter group (it keeps its empty groups).  The wrapper's ``last_epoch`` stops at the end of
    the warm-up, from there on the cosine scheduler counts."""
    st = scheduler.state_dict()
    warm, epoch = scheduler.warmup_epochs, scheduler.last_epoch
    finished = epoch >= warm
    base, last = [scheduler.base_lrs[0]] * n_groups, [scheduler.get_last_lr()[0]] * n_groups
    torch_fields = {"_is_initial": False, "_get_lr_called_within_step": False}
    after_epoch = epoch - warm if finished else 0
    out = dict(st)
    out.update(torch_fields)
    out.update({
        "_init_steps": scheduler._init_steps, "_warmup_steps": scheduler._warmup_steps, "warmup_epochs": warm,
        "finished": finished, "base_lrs": base, "last_epoch": min(epoch, warm), "_step_count": min(epoch, warm) + 1,
        "_last_lr": last, "after_scheduler_type": "CosineAnnealingLR",
        "after_scheduler_dict": dict(torch_fields, T_max=scheduler.total_steps - warm, eta_min=scheduler.eta_min, base_lrs=list(base),
                                     last_epoch=after_epoch, _step_count=after_epoch + 1, _last_lr=list(last)),
    })
    return out
