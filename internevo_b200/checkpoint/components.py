"""Save / load of the individual checkpoint components (reference ``internlm/checkpoint/components.py``)."""
from __future__ import annotations

import copy
import json
import os
import re

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.parallel import is_using_isp
from internevo_b200.utils.storage_manager import get_fns, get_storage_manager, llm_load, llm_save, try_get_storage_backend

from .utils import get_model_topology, get_shard_state_dict, load_shard_state_dict

logger = get_logger(__file__)
_EXPERT_KEY = re.compile(r"^(.*)\.(\d+)\.(?:mlp|feed_forward)\.moe_layer\.experts\.wrapped_experts\.(\d+)\.(.*)$")


def _model_fn():
    tp, pp = gpc.get_local_rank(ParallelMode.TENSOR), gpc.get_local_rank(ParallelMode.PIPELINE)
    if is_using_isp():
        return f"model_tp{tp}_wp{gpc.get_local_rank(ParallelMode.WEIGHT)}_pp{pp}.pt"
    return f"model_tp{tp}_pp{pp}.pt"


def _should_save_model():
    """One replica of every shard writes.  mtp / msp / fsp: data-parallel rank 0 of each (tp, pp) coordinate.  isp: files are
    named by (tp, wp, pp); the first data-parallel replica covers every tensor (sequence) rank and the first weight-data
    replica every weight shard, so ``wdp_rank == 0 or dp_rank == 0`` writes each (tp, wp) combination that exists
    (reference ``components.py:245-252``)."""
    if is_using_isp():
        return gpc.get_local_rank(ParallelMode.WEIGHT_DATA) == 0 or gpc.get_local_rank(ParallelMode.DATA) == 0
    return gpc.get_local_rank(ParallelMode.DATA) == 0


def _chunk_layer_offsets(model) -> dict:
    """``{state-dict prefix of a pipeline chunk: global index of its first block}``.  Every stage numbers its blocks from 0,
    so per-expert file names take the chunk's ``start_layer_idx`` (reference ``components.py:53-92`` uses
    ``pp_rank * layers_per_stage + i``; interleaved chunks are covered too)."""
    offs = {}
    for name, mod in model.named_modules():
        if hasattr(mod, "start_layer_idx") and hasattr(mod, "spec"):
            offs[name] = int(mod.start_layer_idx)
    return offs or {"": 0}


def _layer_offset(prefix: str, offsets: dict) -> int:
    """``prefix`` is everything before ``.<local layer>.`` (e.g. ``model.layers`` or ``1.layers``): longest chunk name wins."""
    best, off = -1, 0
    for name, o in offsets.items():
        if (name == "" or prefix == name or prefix.startswith(name + ".")) and len(name) > best:
            best, off = len(name), o
    return off


def _expert_geometry():
    ep_rank = gpc.get_local_rank(ParallelMode.EXPERT) if gpc.is_initialized(ParallelMode.EXPERT) else 0
    n_local = gpc.config.model.get("num_experts", 1) // max(1, gpc.expert_parallel_size)
    return ep_rank, n_local


def _split_expert_states(states: dict, offsets: dict):
    """Pull expert tensors out into ``{(global layer, global expert): {key: tensor}}``.  Inside a file the key carries the
    GLOBAL expert id (``wrapped_experts.<global>``), so a checkpoint can be re-read under another expert-parallel size and
    exchanged with the reference (``try_save_moe_checkpoint``, reference ``components.py:53-92``)."""
    ep_rank, n_local = _expert_geometry()
    experts, rest = {}, {}
    for k, v in states.items():
        m = _EXPERT_KEY.match(k)
        if m is None:
            rest[k] = v
            continue
        prefix, layer, local_e, tail = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)
        glob_e = ep_rank * n_local + local_e
        gkey = k[: m.start(3)] + str(glob_e) + k[m.end(3):]
        assert gkey.endswith(tail)
        experts.setdefault((layer + _layer_offset(prefix, offsets), glob_e), {})[gkey] = v
    return rest, experts


def _expert_keys_to_local(st: dict, glob_e: int, local_e: int) -> dict:
    out = {}
    for k, v in st.items():
        m = _EXPERT_KEY.match(k)
        if m is None or int(m.group(3)) not in (glob_e, local_e):
            raise KeyError(f"unexpected key {k!r} in the file of expert {glob_e}")
        out[k[: m.start(3)] + str(local_e) + k[m.end(3):]] = v
    return out


# ---------------------------------------------------------------------------------------------------------------------
# ISP file layout.  In memory this framework keeps the token embedding whole on every rank of the sequence (tensor) group and
# shards the head's vocabulary rows over the WEIGHT group; the reference splits the embedding along the hidden dimension and
# the head's vocabulary rows over the TENSOR group (``modules/embedding.py:44-47``, ``ops/linear.py:24-83``).  Files are
# written - and understood - in the reference's layout so ``model_tp{t}_wp{w}_pp{p}.pt`` moves between the two code bases; the
# conversion costs one all-gather of the head per save / load.  Older files of this framework (whole embedding, weight-group
# head rows) are recognised by their shapes and loaded as they are.
# ---------------------------------------------------------------------------------------------------------------------
def _isp_embed_head_keys(model, states):
    emb, head = [], []
    for mod in model.modules():
        spec = getattr(mod, "spec", None)
        if spec is not None and hasattr(spec, "embed_name") and hasattr(spec, "head_name"):
            for k in states:
                if k.endswith(f"{spec.embed_name}.weight") and k not in emb:
                    emb.append(k)
                elif k.endswith(f"{spec.head_name}.weight") and k not in head:
                    head.append(k)
    return emb, head


def _gather_rows(t: torch.Tensor, mode: ParallelMode) -> torch.Tensor:
    n = gpc.get_world_size(mode)
    if n <= 1:
        return t
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else t.device
    src = t.detach().to(dev).contiguous()
    out = torch.empty(n * src.shape[0], *src.shape[1:], dtype=src.dtype, device=dev)
    dist.all_gather_into_tensor(out, src, group=gpc.get_group(mode))
    return out


def _my_slice(full: torch.Tensor, mode: ParallelMode, dim: int) -> torch.Tensor:
    n, r = gpc.get_world_size(mode), gpc.get_local_rank(mode)
    size = full.shape[dim] // n
    return full.narrow(dim, r * size, size).contiguous()


def _isp_states_to_file_layout(model, states: dict) -> dict:
    """Collective over the TENSOR and WEIGHT groups (every rank of a pipeline stage holds the same keys)."""
    emb, head = _isp_embed_head_keys(model, states)
    vocab, hidden = gpc.config.model.vocab_size, gpc.config.model.hidden_size
    out = dict(states)
    for k in emb:
        if tuple(states[k].shape) == (vocab, hidden):
            out[k] = _my_slice(states[k].detach(), ParallelMode.TENSOR, 1)
    for k in head:
        if states[k].shape[0] * gpc.get_world_size(ParallelMode.WEIGHT) == vocab:
            out[k] = _my_slice(_gather_rows(states[k], ParallelMode.WEIGHT), ParallelMode.TENSOR, 0)
    return out


def _isp_states_from_file_layout(model, states: dict) -> dict:
    emb, head = _isp_embed_head_keys(model, model.state_dict())
    vocab, hidden = gpc.config.model.vocab_size, gpc.config.model.hidden_size
    tp, wp = gpc.get_world_size(ParallelMode.TENSOR), gpc.get_world_size(ParallelMode.WEIGHT)
    for k in emb:
        if k in states and tuple(states[k].shape) == (vocab, hidden // tp) and tp > 1:
            parts = _gather_rows(states[k].t().contiguous(), ParallelMode.TENSOR)      # gather along hidden
            states[k] = parts.t().contiguous().cpu()
    for k in head:
        if k in states and states[k].shape[0] == vocab // tp and not (tp == wp):
            states[k] = _my_slice(_gather_rows(states[k], ParallelMode.TENSOR), ParallelMode.WEIGHT, 0).cpu()
        # tp == wp: the two groups are the same ranks and the row slices coincide; vocab // wp rows: this framework's older files
    return states


def save_model_checkpoint(folder, model):
    """``model_tp*_pp*.pt`` + topology json (+ one file per global expert)."""
    states = get_shard_state_dict(model)
    if is_using_isp():
        states = _isp_states_to_file_layout(model, states)
    states = {k: v.detach().clone().cpu() if torch.is_tensor(v) else v for k, v in states.items()}
    states, experts = (_split_expert_states(states, _chunk_layer_offsets(model))
                       if gpc.config.model.get("num_experts", 1) > 1 else (states, {}))
    if folder is None:
        return
    tp = gpc.get_local_rank(ParallelMode.TENSOR)
    if _should_save_model():
        fn = _model_fn()
        llm_save(os.path.join(folder, fn), saved_obj=states)
        topo = json.dumps(get_model_topology(model))
        topo_fn = fn.replace("model_", "topo_").replace(".pt", ".json")
        topo_path = os.path.join(folder, topo_fn)
        get_storage_manager()._client(topo_path)[0].upload_bytes(topo.encode(), try_get_storage_backend(topo_path)[1])
    _write_expert_files(folder, experts, tp)
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def _write_expert_files(folder, experts: dict, tp: int) -> None:
    # experts are replicated over EXPERT_DATA: rank 0 of that group writes
    if experts and (not gpc.is_initialized(ParallelMode.EXPERT_DATA) or gpc.get_local_rank(ParallelMode.EXPERT_DATA) == 0):
        for (layer, e), st in experts.items():
            llm_save(os.path.join(folder, f"model_moe_layer{layer}_expert{e}_tp{tp}.pt"), saved_obj=st)


def try_save_moe_checkpoint(folder, model, tp_rank=None, pp_rank=None) -> None:  # noqa: ARG001
    """Write ``model_moe_layer{L}_expert{E}_tp{t}.pt`` for every (global layer, global expert) this rank holds - the expert half
    of ``save_model_checkpoint`` on its own (reference ``checkpoint/components.py:53-92``).  ``pp_rank`` is accepted for
    signature parity: the global layer id comes from the chunk's ``start_layer_idx``, which also covers interleaved chunks."""
    if gpc.config.model.get("num_experts", 1) <= 1 or folder is None:
        return
    states = {k: v.detach().clone().cpu() if torch.is_tensor(v) else v for k, v in get_shard_state_dict(model).items()}
    _, experts = _split_expert_states(states, _chunk_layer_offsets(model))
    _write_expert_files(folder, experts, gpc.get_local_rank(ParallelMode.TENSOR) if tp_rank is None else tp_rank)


def try_load_moe_checkpoint(folder, model, state_dict: dict, tp_rank=None, pp_rank=None) -> dict:  # noqa: ARG001
    """Merge this rank's expert files (this stage's layers x this rank's experts, keys back in local numbering) into
    ``state_dict`` and return it (reference ``checkpoint/components.py:31-50``)."""
    if gpc.config.model.get("num_experts", 1) > 1:
        state_dict.update(_load_expert_files(folder, get_fns(folder), model))
    return state_dict


def _load_expert_files(folder, fns, model) -> dict:
    """This stage's layers x this rank's experts, keys mapped back to the local numbering of both."""
    tp = gpc.get_local_rank(ParallelMode.TENSOR)
    ep_rank, n_local = _expert_geometry()
    offsets = _chunk_layer_offsets(model)
    # (global layer) -> state-dict prefix + local layer, for every MoE block this rank holds
    where = {}
    for k in model.state_dict().keys():
        m = _EXPERT_KEY.match(k)
        if m is not None:
            where[int(m.group(2)) + _layer_offset(m.group(1), offsets)] = (m.group(1), int(m.group(2)))
    out = {}
    for fn in fns:
        m = re.match(rf"model_moe_layer(\d+)_expert(\d+)_tp{tp}\.pt$", fn)
        if m is None:
            continue
        layer, e = int(m.group(1)), int(m.group(2))
        if layer not in where or not ep_rank * n_local <= e < (ep_rank + 1) * n_local:
            continue
        prefix, local_layer = where[layer]
        st = _expert_keys_to_local(llm_load(os.path.join(folder, fn), map_location="cpu"), e, e - ep_rank * n_local)
        for k, v in st.items():   # the file was written under the saver's chunk prefix / local layer: re-home it
            km = _EXPERT_KEY.match(k)
            out[f"{prefix}.{local_layer}." + k[km.end(2) + 1:]] = v
    return out


def load_model_checkpoint(folder, model):
    """Loads this rank's shard; asserts that tp/pp sizes match the checkpoint (reference ``:146-158``)."""
    fns = get_fns(folder)
    tp_size, pp_size = gpc.get_world_size(ParallelMode.TENSOR), gpc.get_world_size(ParallelMode.PIPELINE)
    max_pp = max_tp = 0
    for fn in fns:
        if fn.startswith("model_t") and not fn.endswith(".md5") and "moe_layer" not in fn:
            segs = fn.replace(".pt", "").split("_")
            max_pp = max(max_pp, int(segs[-1][2:]))
            max_tp = max(max_tp, int(segs[1][2:]))
    assert pp_size == max_pp + 1, f"The weights are save for {max_pp + 1} pipelines, while current has {pp_size} pipelines"
    assert tp_size == max_tp + 1, f"The weights are save for {max_tp + 1} parallelism, while current has {tp_size}"
    fp = os.path.join(folder, _model_fn())
    states = llm_load(fp, map_location="cpu")
    if is_using_isp():
        states = _isp_states_from_file_layout(model, states)
    try_load_moe_checkpoint(folder, model, states)
    missing_k, unexpected_keys = load_shard_state_dict(model, states, strict=False)
    lost = [k for k in missing_k if _EXPERT_KEY.match(k)]
    assert not lost, f"expert weights missing from the checkpoint (they would keep their random init): {lost[:4]}..."
    if len(missing_k) != 0 and gpc.is_rank_for_log():
        logger.warning(f"Warning: missing keys {missing_k}")
    if len(unexpected_keys) != 0 and gpc.is_rank_for_log():
        logger.warning(f"Warning: unexpected keys {unexpected_keys}")
    del states
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def _optimizer_fn():
    tp, pp = gpc.get_local_rank(ParallelMode.TENSOR), gpc.get_local_rank(ParallelMode.PIPELINE)
    if is_using_isp():
        return (f"optimizer_tp{tp}_wp{gpc.get_local_rank(ParallelMode.WEIGHT)}_pp{pp}_"
                f"dp{gpc.get_local_rank(ParallelMode.DATA)}.pt")
    return f"optimizer_tp{tp}_pp{pp}_zo{gpc.get_local_rank(ParallelMode.ZERO1)}.pt"


def _optimizer_format() -> str:
    ck = gpc.config.get("ckpt", None) if gpc.config is not None else None
    fmt = ck.get("optimizer_ckpt_format", "internevo_b200") if ck is not None else "internevo_b200"
    assert fmt in ("internevo_b200", "reference"), f"ckpt.optimizer_ckpt_format: 'internevo_b200' or 'reference', got {fmt!r}"
    return fmt


def _convertible(optim) -> bool:
    """The layout converters need the arena optimizer with its model bound, outside ISP (whose optimizer files are per (tp, wp,
    pp, dp) coordinate and whose embedding / head change their sharding between memory and file)."""
    return hasattr(optim, "groups") and getattr(optim, "_model", None) is not None and not is_using_isp()


def _gather_named(optim) -> dict:
    """Per-parameter optimizer state of this (tp, pp) coordinate, assembled over the ZeRO group (every rank gets all of it;
    gathered range by range so the transient buffer is one range, not the arena)."""
    from collections import OrderedDict

    import torch.distributed as dist

    from .optimizer_interchange import _FILE_KEY, KINDS

    optim.flush_param_update()
    named = {"grad_scaler": optim.grad_scaler.state_dict(), "groups": OrderedDict()}
    for g in optim.groups:
        compact = {"master": g.master, "exp_avg": g.exp_avg, "exp_avg_sq": g.exp_avg_sq}
        full = {}
        for k in KINDS:
            if g.zero_size == 1:
                full[k] = compact[k].detach().float().cpu()
                continue
            out, group = torch.empty(g.total, dtype=torch.float32), gpc.get_group(g.zero_mode)
            for lo, hi in g.ranges:
                n = (hi - lo) // g.zero_size
                mine = compact[k][lo // g.zero_size: lo // g.zero_size + n].contiguous()
                parts = [torch.empty_like(mine) for _ in range(g.zero_size)]
                dist.all_gather(parts, mine, group=group)
                out[lo:hi] = torch.cat(parts).float().cpu()
            full[k] = out
        params = OrderedDict()
        for p in g.ordered:
            o = g.offsets[id(p)]
            params[optim.param_name(p)] = {k: full[k][o: o + p.numel()].view(p.shape).clone() for k in KINDS}
        named["groups"][g.name] = {"step": g.step, "hyper": {k: v for k, v in g.cfg.items() if k != "params"}, "params": params}
    assert set(_FILE_KEY) == set(KINDS)
    return named


def save_optimizer_checkpoint(optim, state_path):
    """One file per (tp, pp, zero) coordinate; ranks beyond the first ZeRO replica hold identical shards and skip.
    ``ckpt.optimizer_ckpt_format = "reference"`` writes the files in the reference's layout (whole parameters per ZeRO rank, one
    flat buffer per group; ``checkpoint/optimizer_interchange.py``) so that the reference can resume from them."""
    if optim is None or state_path is None:
        return
    zero_size = gpc.get_world_size(ParallelMode.ZERO1)
    dp_rank = gpc.get_local_rank(ParallelMode.WEIGHT_DATA if is_using_isp() else ParallelMode.DATA)
    if _optimizer_format() == "reference":
        assert _convertible(optim), "ckpt.optimizer_ckpt_format='reference' needs HybridZeroOptimizer (not FSDP) and no ISP"
        from .optimizer_interchange import reference_file_from_named

        named = _gather_named(optim)            # collective over the ZeRO groups: every rank takes part before anyone returns
        if dp_rank >= zero_size:
            return
        moe = gpc.expert_parallel_group_names[0] if gpc.expert_parallel_group_names else None
        layout = {g.name: (g.zero_size, g.zero_rank) for g in optim.groups}
        layout.setdefault("default", (zero_size, gpc.get_local_rank(ParallelMode.ZERO1)))
        mine = reference_file_from_named(named, optim._model, gpc.config.model.get("dtype", torch.float32), layout, moe_group=moe)
        llm_save(os.path.join(state_path, _optimizer_fn()), saved_obj=mine)
        return
    if dp_rank >= zero_size and not is_using_isp():
        return
    states = optim.state_dict()
    llm_save(os.path.join(state_path, _optimizer_fn()), saved_obj=states)
    if gpc.is_rank_for_log() and "zero_devide_optim_plan" in states:
        llm_save(os.path.join(state_path, optim.rank_unique_id), saved_obj=states["zero_devide_optim_plan"])


def _same_arena_layout(optim, states) -> bool:
    if not hasattr(optim, "groups") or "groups" not in states or len(states["groups"]) != len(optim.groups):
        return False
    return all(st.get("layout") == "range-interleaved" and st["total"] == g.total and st["zero_size"] == g.zero_size
               and st["zero_rank"] == g.zero_rank and [tuple(r) for r in st["ranges"]] == g.ranges
               for g, st in zip(optim.groups, states["groups"]))


def load_optimizer_checkpoint(folder, optim):
    """Loads this rank's optimizer shard.  Files written with the current layout are read directly; files of ANOTHER layout -
    a different ZeRO / data-parallel size or bucket size, or the reference's own parameter-wise files - are converted through
    the per-parameter form (every rank reads the ZeRO shards of its (tp, pp) coordinate; ``optimizer_interchange.py``)."""
    fns = get_fns(folder)
    max_tp = max_pp = max_zo = 0
    for fn in fns:
        if fn.startswith("optimizer_") and not fn.endswith(".md5"):
            if is_using_isp():
                continue
            _, tp, pp, zo = os.path.splitext(fn)[0].split("_")
            max_zo, max_tp, max_pp = max(max_zo, int(zo[2:])), max(max_tp, int(tp[2:])), max(max_pp, int(pp[2:]))
    if not is_using_isp():
        assert gpc.get_world_size(ParallelMode.PIPELINE) == max_pp + 1 and gpc.get_world_size(ParallelMode.TENSOR) == max_tp + 1
    mine = _optimizer_fn()
    probe = mine if mine in fns or is_using_isp() else mine[: mine.rindex("_zo")] + "_zo0.pt"
    states = llm_load(os.path.join(folder, probe), map_location="cpu")
    reference_format = "base_optim_states" in states
    if reference_format or (not is_using_isp() and not _same_arena_layout(optim, states)
                            and not gpc.config.get("only_load_lr", False)):
        assert _convertible(optim), (
            f"The optimizer states are saved for {max_zo + 1} zero parallel in another layout, while current has "
            f"{gpc.get_world_size(ParallelMode.ZERO1)} zero broadcast range, and this optimizer cannot convert them")
        from .optimizer_interchange import arena_states_from_named, named_from_arena_files, named_from_reference

        prefix, cache = mine[: mine.rindex("_zo")], {probe: states}

        def zo_file(z):
            fn = f"{prefix}_zo{z}.pt"
            if fn not in cache:
                cache[fn] = llm_load(os.path.join(folder, fn), map_location="cpu")
            return cache[fn]

        if reference_format:
            if gpc.config.get("only_load_lr", False):
                optim.grad_scaler.load_state_dict(states["grad_scaler"])
                lrs = {pg.get("name"): pg["lr"] for pg in states["base_optim_states"]["param_groups"] if "lr" in pg}
                for g in optim.groups:
                    if g.name in lrs:
                        g.cfg["lr"] = lrs[g.name]
                return
            model_keys = llm_load(os.path.join(folder, _model_fn()), map_location="cpu")
            moe = gpc.expert_parallel_group_names[0] if gpc.expert_parallel_group_names else None
            if moe is not None:      # the reference keeps expert weights in per-expert files: put them back in key order
                model_keys = _with_expert_keys(folder, fns, optim._model, model_keys)

            def files_of_group(name):
                if name == moe and gpc.is_initialized(ParallelMode.EXPERT_DATA):
                    # the replicas of this rank's experts: their files are named by their rank in the ZERO1 group
                    zero_ranks = gpc.get_ranks_in_group(ParallelMode.ZERO1)
                    return [zo_file(zero_ranks.index(r)) for r in gpc.get_ranks_in_group(ParallelMode.EXPERT_DATA)]
                return [zo_file(z) for z in range(max_zo + 1)]

            named = named_from_reference(files_of_group, model_keys, optim._model, optim,
                                         gpc.config.model.get("dtype", torch.float32), moe_group=moe)
            what = "the reference's parameter-wise layout"
        else:
            named = named_from_arena_files([zo_file(z) for z in range(max_zo + 1)])
            what = f"a {max_zo + 1}-way ZeRO layout"
        if gpc.is_rank_for_log():
            logger.info(f"optimizer checkpoint {folder}: converting from {what} to the current "
                        f"{gpc.get_world_size(ParallelMode.ZERO1)}-way arena layout")
        states = arena_states_from_named(named, optim)
        del cache, named
    optim.load_state_dict(states)
    del states
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def _with_expert_keys(folder, fns, model, model_keys: dict) -> dict:
    """The reference's model state dict (its key order untouched) followed by this rank's expert tensors in registration order
    (local numbering).  Experts form an optimizer group of their own, so only the order AMONG them matters."""
    from collections import OrderedDict

    from .optimizer_interchange import _reference_order

    experts = _load_expert_files(folder, fns, model)
    merged = OrderedDict(model_keys)
    for k in _reference_order([k for k in get_shard_state_dict(model).keys() if k in experts]):
        merged[k] = experts[k]
    return merged


def load_sampler(ckpt_path: str, sampler):
    sampler_states = llm_load(os.path.join(ckpt_path, "sampler.pt"))
    sampler.load_state_dict(sampler_states)
    if gpc.is_rank_for_log():
        pstate = copy.deepcopy(sampler_states)
        pstate.pop("indices", None)
        pstate.pop("rng_state", None)
        logger.info(f"reload sampler_states:{pstate}")


def load_context(ckpt_path: str, train_state):
    context_stuffs = llm_load(os.path.join(ckpt_path, "context.pt"))
    train_state.load_state_dict(context_stuffs)
    if gpc.is_rank_for_log():
        logger.info(f"reload train_state:{train_state}")


def load_scheduler(ckpt_path: str, lr_scheduler, optimizer, train_state):
    """Resume the LR schedule; the base LR may be overridden by the current config (reference ``:431-456``)."""
    learning_rate = train_state.lr
    scheduler_states = llm_load(os.path.join(ckpt_path, "schedulder.pt"))
    if learning_rate != scheduler_states["base_lrs"][0] and gpc.is_rank_for_log():
        logger.warning(f"Using new learning rate {learning_rate} to replace old learn rate {scheduler_states['base_lrs'][0]}.")
    base_lrs = copy.deepcopy(scheduler_states["base_lrs"])
    scheduler_states["base_lrs"] = [learning_rate] * len(scheduler_states["base_lrs"])
    if "after_scheduler_dict" in scheduler_states:
        scheduler_states["after_scheduler_dict"]["base_lrs"] = [learning_rate] * len(
            scheduler_states["after_scheduler_dict"]["base_lrs"])
    lr_scheduler.load_state_dict(scheduler_states)  # last_epoch == number of completed steps (closed-form schedule)
    del scheduler_states
    ratio = learning_rate / base_lrs[0] if base_lrs and base_lrs[0] else 1.0
    lr_scheduler._last_lr = [lr * ratio for lr in lr_scheduler.get_lr()] if ratio != 1.0 else lr_scheduler.get_lr()
    for g, lr in zip(optimizer.param_groups, lr_scheduler._last_lr):
        g["lr"] = lr
    if gpc.is_rank_for_log():
        logger.info(f"reload load_scheduler:{lr_scheduler}")
