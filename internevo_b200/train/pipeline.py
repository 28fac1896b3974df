"""Glue between config and runtime: model / optimizer / ISP communicator construction, batch loading, per-step metrics,
profiler (reference ``internlm/train/pipeline.py:157-633``)."""
from __future__ import annotations

import time
from typing import Callable, Iterable, List, Optional, Union

import torch
from torch import nn
from torch.utils.data import DataLoader

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.core.context.random import set_mode
from internevo_b200.core.naive_amp import NaiveAMPModel, set_fp32_attr_to_module
from internevo_b200.core.scheduler.base_scheduler import attach_host_max_seqlen
from internevo_b200.core.trainer import TrainState
from internevo_b200.models.metrics import SchedulerMetricHook
from internevo_b200.monitor import send_heartbeat, set_env_var
from internevo_b200.monitor.monitor import monitor_manager as mm
from internevo_b200.solver.optimizer import HybridZeroOptimizer
from internevo_b200.solver.schedulers import Beta2Scheduler, FineTuneCosineAnnealingWarmupLR
from internevo_b200.train.utils import create_param_groups
from internevo_b200.utils.common import DummyProfile, SchedulerHook
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.megatron_timers import megatron_timer as timer
from internevo_b200.utils.parallel import is_using_isp, sync_model_param, sync_model_replica_param_group
from internevo_b200.utils.registry import MODEL_INITIALIZER
from internevo_b200.utils.timeout import llm_timeout

logger = get_logger(__file__)


def set_fp32_attr_for_model(model: Union[nn.Module, nn.ModuleList]):
    """``use_fp32_norm``: tag every norm module so ``NaiveAMPModel`` keeps it in fp32 (reference ``:88-95``)."""
    from internevo_b200.ops.norm import LayerNorm, RMSNorm

    models = model if isinstance(model, nn.ModuleList) else [model]
    for m in models:
        for sub in m.modules():
            if isinstance(sub, (RMSNorm, LayerNorm, nn.LayerNorm)):
                set_fp32_attr_to_module(sub)


def set_parallel_attr_for_param_groups(model: Union[nn.Module, nn.ModuleList]):
    """(Re-)tag every parameter with its reduction class: replica (norms, MoE gates), tensor / weight ZeRO-parallel
    (linears), tensor-data parallel (ISP embedding / head), expert-data parallel (MoE experts) — reference
    ``internlm/train/pipeline.py:98-154``.  The decoders tag their own parameters at construction
    (``PackedDecoder._set_param_attrs``); this entry point serves user models assembled from the public modules."""
    from internevo_b200.core.context import (IS_REPLICA_ZERO_PARALLEL, IS_TENSOR_DATA_PARALLEL,
                                             IS_TENSOR_EXPERT_DATA_PARALLEL, IS_TENSOR_ZERO_PARALLEL,
                                             IS_WEIGHT_ZERO_PARALLEL)
    from internevo_b200.models.modules import Embedding1D, VocabParallelEmbedding
    from internevo_b200.models.moe import MoE, is_moe_param
    from internevo_b200.ops.norm import LayerNorm, RMSNorm
    from internevo_b200.parallel.linear import BaseScaleColumnParallelLinear
    from internevo_b200.utils.parallel import is_using_isp

    isp = gpc.config is not None and is_using_isp()
    all_tags = (IS_REPLICA_ZERO_PARALLEL, IS_TENSOR_DATA_PARALLEL, IS_TENSOR_EXPERT_DATA_PARALLEL, IS_TENSOR_ZERO_PARALLEL,
                IS_WEIGHT_ZERO_PARALLEL)

    def tag(p, name):
        for t in all_tags:
            if hasattr(p, t):
                delattr(p, t)
        setattr(p, name, True)

    models = model if isinstance(model, nn.ModuleList) else [model]
    for m in models:
        m = m.model if hasattr(m, "model") and isinstance(getattr(m, "model"), nn.Module) else m
        for p in m.parameters():                                   # default class: every linear
            if is_moe_param(p):
                tag(p, IS_TENSOR_EXPERT_DATA_PARALLEL)
            else:
                tag(p, IS_WEIGHT_ZERO_PARALLEL if isp else IS_TENSOR_ZERO_PARALLEL)
        for sub in m.modules():
            if isinstance(sub, (RMSNorm, LayerNorm, nn.LayerNorm)):
                for p in sub.parameters():
                    tag(p, IS_REPLICA_ZERO_PARALLEL)
            elif isinstance(sub, MoE):
                gate = getattr(sub.moe_layer, "gate", None) or getattr(sub.moe_layer, "wg", None)
                for p in (gate.parameters() if gate is not None else []):
                    tag(p, IS_REPLICA_ZERO_PARALLEL)
            elif isinstance(sub, (Embedding1D, VocabParallelEmbedding, BaseScaleColumnParallelLinear)):
                for p in sub.parameters():
                    tag(p, IS_TENSOR_DATA_PARALLEL if isp else IS_TENSOR_ZERO_PARALLEL)


@llm_timeout(func_name="initialize_model")
def initialize_model(pre_process_func: Optional[Callable] = None, post_process_func: Optional[Callable] = None):
    """Build this rank's model chunk(s) from the registry, wrap in ``NaiveAMPModel``, synchronise replicas."""
    if pre_process_func:
        pre_output = pre_process_func()
    import internevo_b200.models  # noqa: F401  registers the builders

    if gpc.config.get("fused_comm", False) and torch.cuda.is_available() and not is_using_isp():
        # tensor-parallel linears run as fused GEMM+collective kernels over peer memory (parallel/fused.py)
        from internevo_b200.parallel import fused

        fused.enable_tp(gpc.get_group(ParallelMode.TENSOR))
    kwargs = dict(gpc.config.model)
    kwargs.pop("output_to_fp32", None)
    for k in ("num_experts", "moe_use_residual", "moe_type"):
        if "MoE" not in gpc.config.model_type:
            kwargs.pop(k, None)
    model = MODEL_INITIALIZER.get_module(module_name=gpc.config.model_type)(**kwargs)
    if post_process_func:
        post_process_func(pre_output)
    if gpc.config.get("use_fp32_norm", False):
        set_fp32_attr_for_model(model)
    dtype = gpc.config.model.get("dtype", torch.half)
    # logits stay in the compute dtype: the loss kernel accumulates in fp32 (reference converts on the last stage)
    model = NaiveAMPModel(model=model, output_to_fp32=bool(gpc.config.model.get("output_to_fp32", False)), dtype=dtype,
                          sync_buffer=False)
    sync_model_param(model)
    sync_model_replica_param_group(model)
    set_mode(ParallelMode.WEIGHT_DATA if is_using_isp() else ParallelMode.DATA)
    if gpc.config.parallel.zero1.get("fsdp", False):
        model = wrap_FSDP_model(model)
    return model


def wrap_FSDP_model(model: Union[nn.Module, nn.ModuleList]):
    """Optional torch-FSDP (ZeRO-3) wrapping over the ZERO1 group (reference ``:217-250``)."""
    import functools

    from torch.distributed.fsdp import BackwardPrefetch
    from torch.distributed.fsdp import FullyShardedDataParallel as FSDP
    from torch.distributed.fsdp.fully_sharded_data_parallel import ShardingStrategy
    from torch.distributed.fsdp.wrap import transformer_auto_wrap_policy

    from internevo_b200.models.decoder import DecoderLayer

    policy = functools.partial(transformer_auto_wrap_policy, transformer_layer_cls={DecoderLayer})
    kw = {}
    if torch.cuda.is_available():
        kw["device_id"] = torch.cuda.current_device()
    return FSDP(module=model, process_group=gpc.get_group(ParallelMode.ZERO1), sharding_strategy=ShardingStrategy.FULL_SHARD,
                auto_wrap_policy=policy, forward_prefetch=True, backward_prefetch=BackwardPrefetch.BACKWARD_PRE,
                limit_all_gathers=True, use_orig_params=True, **kw)


def initialize_isp_communicator(model: Union[nn.Module, nn.ModuleList]):
    """ISP only: build the weight all-gather / grad reduce-scatter communicator and register it with ``ISPLinear``."""
    if not is_using_isp() or gpc.get_world_size(ParallelMode.WEIGHT) <= 1:
        return None
    from internevo_b200.core.communication.isp import ISPCommModelConfig, ISPCommunicator
    from internevo_b200.parallel.linear import ISPLinear

    cfg = ISPCommModelConfig(gpc.config.model.dtype,
                             torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"),
                             gpc.config.model.get("checkpoint", 0))
    comm = ISPCommunicator(model, cfg, gpc.config.parallel.weight.get("overlap", False),
                           gpc.config.parallel.weight.get("memory_pool", False), gpc.get_group(ParallelMode.WEIGHT))
    ISPLinear.register_communicator(comm)
    return comm


@llm_timeout(func_name="initialize_optimizer")
def initialize_optimizer(model: Union[nn.Module, nn.ModuleList], isp_communicator=None):
    """→ ``(optimizer, beta2_scheduler, lr_scheduler)``."""
    adam_cfg = gpc.config.adam
    params = create_param_groups(model, adam_cfg.weight_decay)
    for g in params:
        g.setdefault("lr", adam_cfg.lr)
        g["betas"] = (adam_cfg.adam_beta1, adam_cfg.adam_beta2)
        g["eps"] = adam_cfg.adam_eps

    class _ParamGroups:  # the fused update lives in HybridZeroOptimizer; this is the hyper-parameter carrier
        def __init__(self, groups):
            self.param_groups = groups

    naive = _ParamGroups(params)
    zero_cfg = gpc.config.hybrid_zero_optimizer
    if gpc.is_using_parallel_mode(ParallelMode.PIPELINE) and not gpc.is_pipeline_first_stage(ignore_virtual=True):
        zero_cfg.overlap_sync_grad = False
    if gpc.config.parallel.zero1.get("fsdp", False):
        from internevo_b200.solver.optimizer import FSDPadaptOptimizer

        optimizer = FSDPadaptOptimizer(naive, grad_scal_cfg=gpc.config.grad_scaler, zero_cfg=zero_cfg)
    else:
        optimizer = HybridZeroOptimizer(naive, grad_scal_cfg=gpc.config.grad_scaler, zero_cfg=zero_cfg,
                                        isp_communicator=isp_communicator)
    if hasattr(optimizer, "bind_model"):
        optimizer.bind_model(model)
    if hasattr(optimizer, "attach_model"):
        optimizer.attach_model(model)   # pre-forward hooks of the update / forward overlap (unsharded groups on CUDA)
    beta2_scheduler = Beta2Scheduler(optimizer=naive, **gpc.config.beta2_scheduler)
    lr_scheduler = FineTuneCosineAnnealingWarmupLR(optimizer, **gpc.config.lr_scheduler)
    return optimizer, beta2_scheduler, lr_scheduler


def get_scheduler_hooks(metric, zero_optim=None, isp_communicator=None, criterion=None) -> List[SchedulerHook]:
    hooks: List[SchedulerHook] = []
    if metric is not None:
        skip = (gpc.is_using_parallel_mode(ParallelMode.PIPELINE) and hasattr(gpc.config.model, "num_chunks")
                and gpc.config.model.num_chunks > 1 and gpc.config.parallel["pipeline"].get("interleaved_overlap", False))
        hooks.append(SchedulerMetricHook(metric=metric, skip=skip, criterion=criterion))
    if isp_communicator is not None:
        from internevo_b200.core.communication.isp import ISPCommunicatorSchedulerHook

        hooks.append(ISPCommunicatorSchedulerHook(isp_communicator, zero_optim))
    return hooks


@llm_timeout(func_name="load_new_batch")
def load_new_batch(train_dl: DataLoader, train_iter: Iterable, train_state: TrainState):
    """Next batch (restarting the iterator at epoch end) with the resume anchor advanced.  ``type_ids`` (dataset type of every
    token, for the per-dataset loss / accuracy metrics) stay in ``batch[0]`` - un-packed to ``[micro, micro_bsz, seq]`` for
    un-packed datasets - and are consumed by the caller with ``metric.set_current_type_ids(batch[0].pop("type_ids"))`` before
    the batch goes to the scheduler (reference ``train/pipeline.py:381-414``, ``train.py:217-218``)."""
    timer("batch-gen").start()
    try:
        batch = next(train_iter)
        if hasattr(train_state, "batch_sampler_iter") and train_state.batch_sampler_iter is not None:
            next(train_state.batch_sampler_iter)
    except StopIteration:
        train_iter = iter(train_dl)
        batch = next(train_iter)
        # The anchor sampler trails the loader's own copy by exactly one `next`: drive its suspended generator over the end
        # so it reshuffles (``get_indices``) once, like the copy did, before a fresh generator is started on the new epoch.
        old_anchor = getattr(train_state, "batch_sampler_iter", None)
        if old_anchor is not None:
            for _ in old_anchor:
                pass
        train_state.batch_sampler_iter = iter(train_state.batch_sampler)
        next(train_state.batch_sampler_iter)
        train_state.num_consumed_samples_in_epoch = 0
    timer("batch-gen").stop()
    if batch[0].get("type_ids", None) is not None and not gpc.config.data.get("use_packed_dataset", True):
        from internevo_b200.data.datasets import unpack_data

        batch[0]["type_ids"] = unpack_data(batch[0]["type_ids"], batch[0]["cu_seqlens"], is_type_ids=True)
    attach_host_max_seqlen(batch[0])
    return batch, train_iter


def initialize_llm_profile(profiling: bool = False, start_time: str = None):
    """``torch.profiler`` on dp0 ∧ tp0 ranks when ``--profiling`` (reference ``:417-459``), else a dummy."""
    if profiling:
        from internevo_b200.utils import nvtx

        nvtx.enable(True)  # layer / fwd / bwd / optimizer / fused-collective ranges show up in the trace
    if profiling and gpc.get_local_rank(ParallelMode.DATA) == 0 and gpc.get_local_rank(ParallelMode.TENSOR) == 0:
        acts = [torch.profiler.ProfilerActivity.CPU]
        if torch.cuda.is_available():
            acts.append(torch.profiler.ProfilerActivity.CUDA)
        return torch.profiler.profile(
            activities=acts, schedule=torch.profiler.schedule(skip_first=3, wait=1, warmup=1, active=1, repeat=1),
            on_trace_ready=torch.profiler.tensorboard_trace_handler(
                f"RUN/{gpc.config.JOB_NAME}/{start_time}/traces/rank{gpc.get_global_rank()}_"
                f"dp{gpc.get_local_rank(ParallelMode.DATA)}_wp{gpc.get_local_rank(ParallelMode.WEIGHT)}_"
                f"tp{gpc.get_local_rank(ParallelMode.TENSOR)}"),
            with_stack=True, with_modules=True, profile_memory=True)
    return DummyProfile()


def _batch_stats(batch):
    cu = batch[0].get("cu_seqlens", None)
    if cu is None:
        return 0, 0, 0, 0
    rows = cu if isinstance(cu, (list, tuple)) else list(cu)
    rows = [r.cpu() if torch.is_tensor(r) else torch.as_tensor(r) for r in rows]
    return (sum(len(b) - 1 for b in rows), max(int((b[1:] - b[:-1]).max()) for b in rows),
            max(len(b) - 1 for b in rows), min(len(b) - 1 for b in rows))


@llm_timeout(func_name="record_current_batch_training_metrics")
def record_current_batch_training_metrics(get_tflops_func, logger, writer, success_update, batch_count, batch,
                                          train_state, optimizer, beta2_scheduler, trainer, start_time, loss, moe_loss,
                                          grad_norm, metric, update_panel=False):
    """Log line + TensorBoard scalars with the reference's metric names (TGS variants, tflops, lr, loss_scale...)."""
    set_env_var(key="LAST_ACTIVE_TIMESTAMP", value=int(time.time()))
    timer.store_last_timers()
    if success_update in (0, True):
        train_state.num_consumed_tokens += batch[1].nelement() * gpc.get_world_size(ParallelMode.DATA)
    acc_perplex = metric.get_metric() if gpc.is_no_pp_or_last_stage() and metric is not None else {}
    if not (success_update and gpc.is_rank_for_log()):
        return None
    lr = optimizer.param_groups[0]["lr"]
    scaler = trainer.engine.optimizer.grad_scaler.scale if hasattr(trainer.engine.optimizer, "grad_scaler") else 1.0
    num_tokens_in_batch = batch[1].nelement()
    num_samples, max_len, max_samples, min_samples = _batch_stats(batch)
    time_cost = time.time() - start_time
    world = gpc.get_world_size(ParallelMode.GLOBAL)
    tk_per_gpu = round(num_tokens_in_batch * gpc.get_world_size(ParallelMode.DATA) / world, 4)
    st = train_state.tgs_statistic
    st["sum_step"] += 1
    for k in ("sum_tg", "sum_last_tg_10", "sum_last_tg_50", "SMA_tg_50"):
        st[k] += tk_per_gpu
    for k in ("sum_time", "sum_last_time_10", "sum_last_time_50", "SMA_time_50"):
        st[k] += time_cost
    st["SMA_tg_50_list"].append(tk_per_gpu)
    st["SMA_time_50_list"].append(time_cost)
    if st["sum_step"] > 50:
        st["SMA_tg_50"] -= st["SMA_tg_50_list"].popleft()
        st["SMA_time_50"] -= st["SMA_time_50_list"].popleft()
    last_tgs_1 = round(tk_per_gpu / time_cost, 2)
    st["sum_tgs"] += last_tgs_1
    if st["sum_step"] % 10 == 0:
        st["last_tgs_10"] = round(st["sum_last_tg_10"] / st["sum_last_time_10"], 2)
        st["sum_last_tg_10"] = st["sum_last_time_10"] = 0
    if st["sum_step"] % 50 == 0:
        st["last_tgs_50"] = round(st["sum_last_tg_50"] / st["sum_last_time_50"], 2)
        st["sum_last_tg_50"] = st["sum_last_time_50"] = 0
    tflops = get_tflops_func(time_cost)
    loss_v = float(loss) if loss is not None else float("nan")
    moe_v = float(moe_loss) if moe_loss is not None else None
    infos = {
        "tflops": tflops, "step": batch_count, "loss": loss_v - moe_v if moe_v is not None else loss_v,
        "tgs (tokens/gpu/second)": round(tk_per_gpu / time_cost, 2), "tgs/last_tgs_1": last_tgs_1,
        "tgs/tgs_all": round(st["sum_tg"] / st["sum_time"], 2), "tgs/tgs_avg": round(st["sum_tgs"] / st["sum_step"], 2),
        "tgs/tgs_SMA": round(st["SMA_tg_50"] / st["SMA_time_50"], 2), "tgs/last_tgs_10": st["last_tgs_10"],
        "tgs/last_tgs_50": st["last_tgs_50"], "lr": lr, "loss_scale": scaler, "grad_norm": grad_norm,
    }
    if moe_v is not None:
        infos["moe_loss"] = moe_v
    infos.update({
        "micro_num": len(batch[1]), "num_consumed_tokens": train_state.num_consumed_tokens,
        "inf_nan_skip_batches": train_state.inf_nan_skip_batches, "num_samples_in_batch": num_samples,
        "largest_length": max_len, "largest_batch": max_samples, "smallest_batch": min_samples,
        "adam_beta2": beta2_scheduler.get_beta2() if beta2_scheduler is not None else None,
        "fwd_bwd_time": round(timer("fwd-bwd").elapsed(), 2),
    })
    infos.update(acc_perplex)
    line = ""
    for key, value in infos.items():
        line += f"{key}={value} "
        if writer is not None:
            if isinstance(value, dict):
                writer.add_scalars(key=key, value=value, step=train_state.step_count)
            elif value is not None:
                writer.add_scalar(key=key, value=value, step=train_state.step_count)
    alert = gpc.config.monitor.alert
    if alert.get("light_monitor_address", None) and batch_count % 50 == 0:
        send_heartbeat("train_metrics", infos)
    logger.info(line)
    mm.monitor_loss_spike(alert_address=alert.get("feishu_alert_address", None), step_count=batch_count,
                          cur_step_loss=loss_v)
    return infos
