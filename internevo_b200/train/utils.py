"""Optimizer parameter groups (reference ``internlm/train/utils.py:11-84``): ``default`` (bf16, ZeRO-sharded),
``embed_head`` (ISP: embedding/head reduce over DATA), ``fp32`` (fp32-tagged modules), one group per expert-parallel
size for MoE experts (reduce over EXPERT_DATA).  Empty groups are dropped (the arenas would be empty)."""
from __future__ import annotations

from typing import Dict, List

import torch

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.parallel import is_expert_param, is_tensor_data_parallel_parameter, is_using_isp


def split_params_into_different_groups_for_optimizer(param_groups) -> List[Dict]:
    if isinstance(param_groups, tuple):
        param_groups = list(param_groups)
    elif isinstance(param_groups, dict):
        param_groups = [param_groups]
    elif not isinstance(param_groups, list):
        raise ValueError(f"Unknown param group type of {type(param_groups)}")
    new_groups: Dict[str, Dict] = {}
    if is_using_isp():
        new_groups["embed_head"] = {"name": "embed_head", "params": [], "optimizer_mode": ParallelMode.DATA}
    new_groups["fp32"] = {"name": "fp32", "params": [], "optimizer_mode": ParallelMode.ZERO1}
    if gpc.config.model.get("num_experts", 1) > 1:
        key = f"moe_ep_size_{gpc.expert_parallel_size}"
        new_groups[key] = {"name": key, "moe": True, "params": [], "optimizer_mode": ParallelMode.EXPERT_DATA}
    for pgroup in param_groups:
        for ori_key in pgroup.keys():
            if ori_key not in ("name", "params"):
                for group in new_groups.values():
                    group[ori_key] = pgroup[ori_key]
        origin = []
        for param in pgroup["params"]:
            if not param.requires_grad:
                continue
            if is_using_isp() and is_tensor_data_parallel_parameter(param):
                new_groups["embed_head"]["params"].append(param)
            elif is_expert_param(param):
                new_groups[getattr(param, "group_name", f"moe_ep_size_{gpc.expert_parallel_size}")]["params"].append(param)
            elif param.dtype == torch.float32 and gpc.config.model.get("dtype", torch.float32) != torch.float32:
                new_groups["fp32"]["params"].append(param)
            else:
                origin.append(param)
        pgroup["params"] = origin
        pgroup["optimizer_mode"] = ParallelMode.ZERO1
    param_groups.extend(g for g in new_groups.values())
    return [g for g in param_groups if len(g["params"]) > 0 or g["name"] == "default"]


def create_param_groups(model, weight_decay):
    parameters = {"params": list(model.parameters()), "name": "default", "weight_decay": weight_decay}
    return split_params_into_different_groups_for_optimizer(parameters)
