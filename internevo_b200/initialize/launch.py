"""Process launch: CLI, rendezvous, config defaults/validation, device + seed setup.

User surface identical to the reference (``internlm/initialize/launch.py``): ``get_default_parser``,
``launch_from_torch``, ``launch_from_slurm``, ``initialize_distributed_env``, ``args_sanity_check``.  The multi-backend
accelerator shim is gone: CUDA (NCCL) when a GPU is visible, CPU (gloo) otherwise.
"""
from __future__ import annotations

import argparse
import gc
import os
from pathlib import Path
from typing import Dict, Union

import torch

from internevo_b200.core.context import Config, ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.initialize.legacy.launch import auto_resume_sanity_check
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.timeout import llm_timeout

logger = get_logger(__file__)

_DTYPES = {
    "torch.bfloat16": torch.bfloat16,
    "torch.float16": torch.float16,
    "torch.half": torch.float16,
    "torch.float32": torch.float32,
    "torch.tf32": torch.float32,
}


def get_default_parser():
    """Same flags as the reference parser (``launch.py:40-68``)."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, help="path to the config file")
    parser.add_argument("--launcher", type=str, default="torch", choices=["slurm", "torch"], help="launcher")
    parser.add_argument("--host", type=str, help="the master address for distributed training")
    parser.add_argument("--port", type=int, default=8888, help="the master port for distributed training")
    parser.add_argument("--world_size", type=int, help="world size for distributed training")
    parser.add_argument("--rank", type=int, help="rank for the default process group")
    parser.add_argument("--local_rank", type=int, help="local rank on the node")
    parser.add_argument("--backend", type=str, default=None, help="backend for distributed communication")
    parser.add_argument("--seed", type=int, default=1024)
    parser.add_argument("--profiling", default=False, action="store_true", help="enable/disable profiling.")
    parser.add_argument("--enable_ali_topology", default=False, action="store_true", help="(ignored) ali topology")
    return parser


def _d(cfg: Config, key: str, value):
    if key not in cfg:
        cfg._add_item(key, value)


def get_config_value(config, key, defalut):
    """``config[key]`` or the default (reference ``launch.py:637-642``; the misspelt keyword is part of its signature)."""
    try:
        return config[key]
    except KeyError:
        return defalut


def args_sanity_check():
    """Fill config defaults and validate combinations (reference ``launch.py:71-445``)."""
    assert gpc.config is not None, "config is not loaded!"
    cfg = gpc.config
    log = gpc.is_rank_for_log()
    _d(cfg, "JOB_NAME", "AnonymousJob")
    _d(cfg, "model_type", "INTERNLM")

    # ---- parallel
    _d(cfg, "parallel", Config())
    par = cfg.parallel
    _d(par, "zero1", dict(size=-1, fsdp=False))
    if isinstance(par.zero1, int):
        par._add_item("zero1", dict(size=par.zero1, fsdp=False))
    _d(par, "pipeline", dict(size=1, interleaved_overlap=False))
    if isinstance(par.pipeline, int):
        par._add_item("pipeline", dict(size=par.pipeline, interleaved_overlap=False))
    _d(par, "tensor", 1)
    _d(par, "weight", dict(size=1, overlap=False, memory_pool=False))
    if isinstance(par.weight, int):
        par._add_item("weight", dict(size=par.weight, overlap=False, memory_pool=False))
    pp = par.pipeline.size
    _d(par.zero1, "fsdp", False)
    assert not (par.zero1.fsdp and pp > 1), "FSDP is not supported when pipeline size > 1"

    # ---- data
    data = cfg.data
    assert data.get("seq_len") is not None, "'seq_len' must be given a value"
    assert data.get("micro_bsz") is not None, "'micro_bsz' must be given a value"
    data._add_item("packed_length", data.seq_len * data.micro_bsz)
    _d(data, "type", "tokenized")
    _d(data, "micro_num", 1)
    if "gradient_accumulation" not in data:
        data._add_item("gradient_accumulation", data.micro_num)
    elif pp == 1:
        assert data.gradient_accumulation == data.micro_num, "for nopp 'gradient_accumulation' should equal 'micro_num'"
    data._add_item("batch_size", data.micro_num)
    for k, v in dict(min_length=0, train_folder=None, valid_folder=None, valid_micro_num=data.micro_num, valid_every=0,
                     empty_cache_and_diag_interval=50, diag_outlier_ratio=1.1, use_packed_dataset=True,
                     fixed_random_dataset_seqlen=False, pack_sample_into_one=False, rampup_batch_size=None,
                     total_steps=0, skip_batches="").items():
        _d(data, k, v)
    data.diag_outlier_ratio = max(1, data.diag_outlier_ratio)

    # ---- checkpoint
    _d(cfg, "ckpt", Config())
    ckpt = cfg.ckpt
    _d(ckpt, "enable_save_ckpt", False if "save_ckpt_folder" not in ckpt else True)
    if ckpt.enable_save_ckpt:
        assert "checkpoint_every" in ckpt and ckpt.checkpoint_every > 0, "enable_save_ckpt needs checkpoint_every > 0"
        assert "save_ckpt_folder" in ckpt, "enable_save_ckpt needs save_ckpt_folder"
        _d(ckpt, "async_upload", False)
        if ckpt.async_upload:
            if not any(ckpt.save_ckpt_folder.startswith(p) for p in ("boto3:", "volc:", "oss2:")):
                if log:
                    logger.warning("file-system checkpoints do not use asynchronous upload; falling back to sync save")
                ckpt.async_upload = False
            else:
                _d(ckpt, "async_upload_tmp_folder", "/dev/shm/internlm_tmp_ckpt/")
        if not ckpt.async_upload:
            ckpt._add_item("async_upload_tmp_folder", None)
        _d(ckpt, "oss_snapshot_freq", float("inf"))
    else:
        for k, v in dict(checkpoint_every=float("inf"), oss_snapshot_freq=float("inf"), save_ckpt_folder=None,
                         async_upload=False, async_upload_tmp_folder=None, snapshot_ckpt_folder=None).items():
            ckpt._add_item(k, v)
    _d(ckpt, "load_ckpt_folder", None)
    _d(ckpt, "stop_file_path", None)
    _d(ckpt, "auto_resume", auto_resume_sanity_check(ckpt))      # True unless an old-style config says load_given_ckpt

    # ---- tensorboard / misc
    _d(cfg, "enable_tb", True)
    _d(cfg, "tensorboard_folder", os.environ.get("tensorboard_folder"))
    _d(cfg, "resume_tb_folder", os.environ.get("resume_tb_folder"))
    torch.backends.cudnn.benchmark = cfg.get("cudnn_benchmark", False)
    torch.backends.cudnn.deterministic = cfg.get("cudnn_deterministic", False)

    # ---- model
    model = cfg.model
    if "dtype" not in model:
        model._add_item("dtype", torch.float16)
    elif isinstance(model.dtype, str):
        assert model.dtype in _DTYPES, f"unsupported model.dtype {model.dtype}"
        if model.dtype == "torch.tf32":
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
        model.dtype = _DTYPES[model.dtype]
    if "checkpoint" in model:
        if model.checkpoint is True:
            model.checkpoint = 1
        elif model.checkpoint is False:
            model.checkpoint = 0
        else:
            assert 0 <= model.checkpoint <= 1, f'model.checkpoint: "{model.checkpoint}" should >=0 and <=1'
    else:
        model._add_item("checkpoint", 0)
    _d(model, "use_flash_attn", True)
    assert model.use_flash_attn == data.use_packed_dataset, (
        "use_packed_dataset should be set same value as use_flash_attn"
    )
    if "MoE" in cfg.get("model_type", "INTERNLM"):
        _d(model, "num_experts", 1)
        _d(model, "moe_use_residual", False)
        _d(model, "moe_type", "GShard")

    # ---- tensor / sequence / weight parallel
    _d(par, "sequence_parallel", False)
    if isinstance(par["tensor"], int):
        par["tensor"] = dict(size=par["tensor"], mode="mtp")
    if par["tensor"].get("mode", None) is None:
        par["tensor"]["mode"] = "mtp"
    tmode = par["tensor"]["mode"]
    assert tmode in ("mtp", "msp", "fsp", "isp"), "invalid tensor parallel mode, only mtp/msp/fsp/isp are supported"
    if tmode == "isp":
        assert not par.zero1.fsdp, "FSDP does not support isp"
    if tmode in ("msp", "fsp", "isp"):
        par.sequence_parallel = True
    if par["weight"].get("overlap", None) is None:
        par["weight"]["overlap"] = False
    if par["weight"].get("memory_pool", None) is None:
        par["weight"]["memory_pool"] = False
    if tmode != "isp":
        assert par["weight"]["size"] <= 1, "weight parallel is only supported with isp"
    if model.get("num_chunks", 1) > 1:
        assert par["pipeline"].get("interleaved_overlap", False) is True, (
            "only support interleaved pipeline scheduler with overlap"
        )

    # ---- monitor
    for key, value in {
        "alert_address": None,
        "monitor": {"alert": {"enable_feishu_alert": False, "feishu_alert_address": None,
                              "light_monitor_address": None, "alert_file_path": None}},
        "tensorboard": {"queue_max_length": 1},
    }.items():
        _d(cfg, key, value)
    _d(cfg.monitor, "alert", dict(enable_feishu_alert=False, feishu_alert_address=None, light_monitor_address=None,
                                  alert_file_path=None))

    # ---- optimizer
    _d(cfg, "hybrid_zero_optimizer", Config())
    opt = cfg.hybrid_zero_optimizer
    if "zero_overlap_communication" in opt:
        opt._add_item("overlap_sync_grad", opt.zero_overlap_communication)
    _d(opt, "overlap_sync_grad", False)
    _d(opt, "overlap_sync_param", False)
    _d(opt, "reduce_bucket_size", 512 * 1024 * 1024)
    _d(opt, "clip_grad_norm", 0.0)
    _d(cfg, "batch_count", 0)
    _d(cfg, "loss", Config(dict(label_smoothing=0.0)))
    _d(cfg.loss, "moe_loss_coeff", 1.0)
    _d(cfg, "grad_scaler", Config(dict(fp16=dict(initial_scale=2 ** 16, min_scale=1, growth_interval=1000),
                                       growth_factor=2, backoff_factor=0.5, max_scale=2 ** 24, hysteresis=2)))
    _d(cfg, "adam", Config(dict(lr=1e-4, adam_beta1=0.9, adam_beta2=0.95, adam_beta2_c=0, adam_eps=1e-8,
                                weight_decay=0.01)))
    _d(cfg, "lr_scheduler", Config(dict(total_steps=data.total_steps, init_steps=0, warmup_ratio=0.01, eta_min=1e-5,
                                        last_epoch=-1)))
    _d(cfg, "beta2_scheduler", Config(dict(init_beta2=cfg.adam.get("adam_beta2", 0.95),
                                           c=cfg.adam.get("adam_beta2_c", 0), cur_iter=-1)))

    if model.get("num_experts", 1) > 1:
        assert not par.zero1.fsdp, "FSDP does not support num_experts > 1"
        assert not (opt.overlap_sync_grad and opt.overlap_sync_param), "not support overlap and moe at the same time"
        assert par.zero1.size in (-1, gpc.get_world_size(ParallelMode.DATA)), (
            "moe only support zero1, set zero1=dict(size=-1,...) can fix this"
        )
    if log:
        logger.info(f"parallel: {par.to_dict()}  data: seq_len={data.seq_len} micro_num={data.micro_num} "
                    f"micro_bsz={data.micro_bsz} packed_length={data.packed_length} "
                    f"dtype={model.dtype} ckpt={model.checkpoint}")


def launch(config: Union[str, Path, Config, Dict], rank: int, world_size: int, host: str, port: int,
           backend: str = None, local_rank: int = None, seed: int = 1024):
    """Load config → global process group → parallel groups → device → seed (reference ``launch.py:448-513``)."""
    assert isinstance(config, (Config, str, Path, dict)), f"bad config type {type(config)}"
    gpc.load_config(config if not isinstance(config, Path) else str(config))
    gpc.init_global_dist(rank, world_size, backend, host, port)
    gpc.init_parallel_groups()
    if torch.cuda.is_available():
        gpc.set_device(local_rank)
    gpc.detect_num_processes_on_current_node()
    gpc.set_seed(seed)
    from internevo_b200.utils.gputest import warmup_process_group

    warmup_process_group()
    if gpc.is_rank_for_log():
        logger.info(
            f"Distributed environment is initialized, data parallel size: {gpc.data_parallel_size}, "
            f"pipeline parallel size: {gpc.pipeline_parallel_size}, tensor parallel size: {gpc.tensor_parallel_size}, "
            f"weight parallel size: {gpc.weight_parallel_size}, zero1 size: {gpc.zero1_parallel_size}"
        )


def launch_from_slurm(config, host: str, port: int, backend: str = None, seed: int = 1024):
    try:
        rank = int(os.environ["SLURM_PROCID"])
        world_size = int(os.environ["SLURM_NPROCS"])
    except KeyError as e:
        raise RuntimeError(f"Could not find {e} in the SLURM environment") from e
    try_bind_numa(global_rank=rank, world_size=world_size)
    launch(config=config, rank=rank, world_size=world_size, host=host, port=port, backend=backend, seed=seed)


def launch_from_torch(config, backend: str = None, seed: int = 1024):
    try:
        rank = int(os.environ["RANK"])
        local_rank = int(os.environ["LOCAL_RANK"])
        world_size = int(os.environ["WORLD_SIZE"])
        host = os.environ["MASTER_ADDR"]
        port = int(os.environ["MASTER_PORT"])
    except KeyError as e:
        raise RuntimeError(f"Could not find {e} in the torch environment") from e
    try_bind_numa(global_rank=rank, world_size=world_size, local_rank=local_rank)
    launch(config=config, local_rank=local_rank, rank=rank, world_size=world_size, host=host, port=port,
           backend=backend, seed=seed)


@llm_timeout(func_name="initialize_distributed_env")
def initialize_distributed_env(config, launcher: str = "slurm", master_port: int = 8888, seed: int = 1024,
                               args_check: bool = True, backend: str = None):
    """Entry used by ``train.py`` (reference ``launch.py:588-634``)."""
    gc.disable()  # automatic GC off for the run; collected explicitly in empty_cache_and_diag
    if launcher == "torch":
        launch_from_torch(config=config, seed=seed, backend=backend)
    elif launcher == "slurm":
        from internevo_b200.utils.common import get_master_node

        launch_from_slurm(config=config, host=get_master_node(), port=master_port, seed=seed, backend=backend)
    else:
        raise ValueError(f"launcher only supports slurm or torch, got {launcher}")
    if args_check:
        args_sanity_check()
    alert = gpc.config.get("monitor", {}).get("alert", {}) if gpc.config.get("monitor") else {}
    if alert and alert.get("light_monitor_address"):
        from internevo_b200.monitor import initialize_light_monitor

        initialize_light_monitor(alert["light_monitor_address"])


def try_bind_numa(global_rank, world_size, local_rank=None):
    """Bind the process to the NUMA node of its GPU when ``numa`` + ``pynvml`` are importable (reference ``:645-684``)."""
    try:
        import numa
        import pynvml
        from numa import memory, schedule
    except (ImportError, ModuleNotFoundError):
        return
    try:
        pynvml.nvmlInit()
        numa_nodes = numa.info.get_max_node() + 1
        gpus = pynvml.nvmlDeviceGetCount()
        if local_rank is None:
            local_rank = global_rank % gpus
        if world_size % gpus != 0 or gpus % numa_nodes != 0:
            return
        per_node = gpus // numa_nodes
        node = local_rank // per_node
        schedule.run_on_nodes(node)
        memory.set_membind_nodes(node)
    except Exception as e:  # pragma: no cover
        logger.warning(f"numa bind failed: {e}")
