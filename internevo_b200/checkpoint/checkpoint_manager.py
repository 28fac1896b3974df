"""Checkpoint manager: periodic NORMAL checkpoints, alternating SNAPSHOT folders, operator stop-file, auto-resume
(reference ``internlm/checkpoint/checkpoint_manager.py:166-650``).

File layout is the reference's: ``{folder}/{step}/model_tp{t}_pp{p}.pt`` (+ ``topo_*.json``, per-expert
``model_moe_layer{L}_expert{E}_tp{t}.pt``), ``optimizer_tp{t}_pp{p}_zo{z}.pt``, ``sampler.pt``, ``context.pt``,
``schedulder.pt`` [sic], ``model_config.pt``, ``config_file.pt`` and the ``{step}.step`` completion marker.
"""
from __future__ import annotations

import os
from enum import Enum
from typing import Dict

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.core.trainer import TrainState
from internevo_b200.initialize.legacy.launch import auto_resume_sanity_check, ckpt_info_sanity_check
from internevo_b200.monitor import send_alert_message
from internevo_b200.utils.common import get_current_device
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.megatron_timers import megatron_timer as timer
from internevo_b200.utils.storage_manager import (
    get_fns,
    get_storage_manager,
    init_storage_manager,
    llm_load,
    llm_save,
    try_get_storage_backend,
)
from internevo_b200.utils.timeout import llm_timeout

from .components import (
    load_context,
    load_model_checkpoint,
    load_optimizer_checkpoint,
    load_sampler,
    load_scheduler,
    save_model_checkpoint,
    save_optimizer_checkpoint,
)
from .load_funcs import LOAD_FUNC_DICT
from .utils import process_load_info

logger = get_logger(__file__)


class CheckpointSaveType(Enum):
    NORMAL_CHECKPOINT = 1
    SNAPSHOT_CHECKPOINT = 2


class CheckpointLoadType(Enum):
    INTERNLM = "internlm"
    INTERNEVO = "internevo"


class CheckpointLoadContent:
    MODEL, SAMPLER, OPIMIZER, SCHEDULAER = "model", "sampler", "optimizer", "scheduler"


class CheckpointLoadMask:
    """``content=("model", "sampler", ...)`` or ``("all",)``."""

    LOAD_CONTENT_DICT = {"model": CheckpointLoadContent.MODEL, "sampler": CheckpointLoadContent.SAMPLER,
                         "optimizer": CheckpointLoadContent.OPIMIZER, "scheduler": CheckpointLoadContent.SCHEDULAER}

    def __init__(self, content: tuple) -> None:
        self.load_set = set(map(lambda x: x.lower(), content))
        if "all" in self.load_set:
            self.load_set = set(self.LOAD_CONTENT_DICT.values())
        else:
            self.load_set = set(map(lambda x: self.LOAD_CONTENT_DICT[x.lower()], content))

    def need_load(self, content):
        return content in self.load_set

    def not_only_load(self, content):
        return content in self.load_set and len(self.load_set) > 1

    def only_load(self, content):
        return set((content,)) == self.load_set

    def __str__(self) -> str:
        return f"{self.load_set}."


def get_shard_state_dict(model):
    return model.state_dict()


def try_load_internevo_ckpt(ckpt_mm, load_info, train_state: TrainState = None):
    """model → context → optimizer → scheduler → sampler, each only if requested (reference ``:61-142``)."""
    load_content_str, load_ckpt_folder, load_content = process_load_info(load_info)
    if load_content.need_load(CheckpointLoadContent.MODEL):
        load_model_checkpoint(folder=load_ckpt_folder, model=ckpt_mm.model)
        load_content_str += f"{CheckpointLoadContent.MODEL}, "
    if load_content.not_only_load(CheckpointLoadContent.MODEL):
        load_context(load_ckpt_folder, train_state)
        if load_content.need_load(CheckpointLoadContent.OPIMIZER):
            load_optimizer_checkpoint(load_ckpt_folder, ckpt_mm.optimizer)
            load_content_str += f"{CheckpointLoadContent.OPIMIZER}, "
        else:
            if gpc.is_rank_for_log():
                logger.warning("CheckpointManager has no 'optimizer', skip reload optim checkpoint!")
        if load_content.need_load(CheckpointLoadContent.SCHEDULAER):
            if ckpt_mm.lr_scheduler:
                load_scheduler(load_ckpt_folder, ckpt_mm.lr_scheduler, ckpt_mm.optimizer, train_state)
                load_content_str += f"{CheckpointLoadContent.SCHEDULAER}, "
        if not load_content.need_load(CheckpointLoadContent.OPIMIZER):
            if ckpt_mm.lr_scheduler and train_state:
                gpc.config.lr_scheduler.last_epoch = train_state.step_count
                ckpt_mm.lr_scheduler.step(train_state.step_count)
            if load_content.need_load(CheckpointLoadContent.SCHEDULAER) and ckpt_mm.optimizer is not None:
                # schedule without optimizer states: still take the loss-scale state and the per-group learning rates from
                # the optimizer file (reference ``only_load_lr``, ``checkpoint_manager.py:109-114``)
                gpc.config.only_load_lr = True
                try:
                    load_optimizer_checkpoint(load_ckpt_folder, ckpt_mm.optimizer)
                except (FileNotFoundError, AssertionError) as e:
                    if gpc.is_rank_for_log():
                        logger.warning(f"only_load_lr: no usable optimizer file in {load_ckpt_folder} ({e})")
                finally:
                    gpc.config.only_load_lr = False
            if load_content.need_load(CheckpointLoadContent.MODEL) and hasattr(ckpt_mm.optimizer, "reload_zero_fp32_buff"):
                # new weights but the old optimizer: the fp32 master must follow the weights, or the first step would write
                # the pre-load values back
                ckpt_mm.optimizer.reload_zero_fp32_buff()
        if load_content.need_load(CheckpointLoadContent.SAMPLER):
            if hasattr(train_state, "batch_sampler") and train_state.batch_sampler is not None:
                load_sampler(load_ckpt_folder, ckpt_mm.train_dl.batch_sampler)
                train_state.init_batch_sampler(ckpt_mm.train_dl.batch_sampler)
                load_content_str += f"{CheckpointLoadContent.SAMPLER}, "
            elif gpc.is_rank_for_log():
                logger.warning("CheckpointManager skip reload 'batch_sampler'")
            if ckpt_mm.train_dl is not None and hasattr(ckpt_mm.train_dl, "dataset"):
                train_state.num_consumed_samples_in_epoch = getattr(train_state, "num_consumed_samples_in_epoch", 0)
    elif ckpt_mm.optimizer is not None and hasattr(ckpt_mm.optimizer, "reload_zero_fp32_buff"):
        # model-only load: refresh the fp32 master copy from the new weights
        ckpt_mm.optimizer.reload_zero_fp32_buff()
    return load_content_str


class CheckpointLoadMethod:
    """Registry of ``ckpt_type`` → loader, so user code can add its own checkpoint format
    (reference ``checkpoint_manager.py:145-163``).  A loader is called as ``fn(ckpt_manager, load_info, train_state)`` and
    returns a string naming what it loaded."""

    LOAD_TYPE_FUNC = {"internevo": try_load_internevo_ckpt, "internlm": try_load_internevo_ckpt, **LOAD_FUNC_DICT}

    @staticmethod
    def register_ckpt_load_type(load_type, load_func):
        if load_type in CheckpointLoadMethod.LOAD_TYPE_FUNC:
            if gpc.is_rank_for_log():
                logger.warning(f"{load_type} has already been registered!")
            return
        CheckpointLoadMethod.LOAD_TYPE_FUNC[load_type] = load_func

    @staticmethod
    def get_ckpt_load_type_func(load_type):
        return CheckpointLoadMethod.LOAD_TYPE_FUNC[getattr(load_type, "value", load_type)]


def try_load_internlm_ckpt_func(ckpt_mm, load_info, *args, func=None, **kwargs):
    """Model-only load through ``func(folder=, model=)`` followed by the master-weight refresh
    (reference ``checkpoint_manager.py:201-219``)."""
    assert func is not None, "pass the loader as func="
    func(folder=load_info["path"], model=ckpt_mm.model)
    if ckpt_mm.optimizer is not None and hasattr(ckpt_mm.optimizer, "reload_zero_fp32_buff"):
        ckpt_mm.optimizer.reload_zero_fp32_buff()
    return f"{CheckpointLoadContent.MODEL}, "


class CheckpointManager:
    """StorageManager is a singleton; the checkpoint manager is created once by ``train.py``."""

    def __init__(self, ckpt_config, model, train_dl=None, optimizer=None, lr_scheduler=None, model_config=None,
                 model_config_file=None, feishu_address=None) -> None:
        self.enable_save_ckpt = ckpt_config.get("enable_save_ckpt", False)
        self.checkpoint_every = ckpt_config.get("checkpoint_every", 100)
        self.save_ckpt_folder = ckpt_config.get("save_ckpt_folder", None)
        self.oss_snapshot_freq: int = ckpt_config.get("oss_snapshot_freq", 50)
        self.stop_file_path = ckpt_config.get("stop_file_path", None)
        if self.save_ckpt_folder:
            self.snapshot_ckpt_folder = ckpt_config.get("snapshot_ckpt_folder",
                                                        os.path.join(self.save_ckpt_folder, "snapshot"))
            self.async_upload_tmp_folder = ckpt_config.get("async_upload_tmp_folder", "/dev/shm/internlm_tmp_ckpt/")
        else:
            self.snapshot_ckpt_folder = None
            self.async_upload_tmp_folder = None
        self.async_upload = ckpt_config.get("async_upload", False)
        self.feishu_address = feishu_address
        self.storage_manager = init_storage_manager(self.enable_save_ckpt, self.async_upload_tmp_folder, self.async_upload)
        self.lr_scheduler, self.optimizer = lr_scheduler, optimizer
        self.train_dl = train_dl
        self.model_config, self.model_config_file = model_config, model_config_file
        # strip the AMP wrapper so keys match the reference's state dict
        self.model = model.model if hasattr(model, "model") and not isinstance(model, torch.nn.ModuleList) else model
        self.load_ckpt_info = ckpt_config.get("load_ckpt_info", None)
        if self.load_ckpt_info is None:      # old-style keys: load_ckpt_folder / load_model_only_folder / load_optimizer
            self.load_ckpt_info = ckpt_info_sanity_check(ckpt_config)
        self.defalut_load_type_func = {CheckpointLoadType.INTERNLM: try_load_internevo_ckpt,
                                       CheckpointLoadType.INTERNEVO: try_load_internevo_ckpt}
        for ckpt_load_type, fn in CheckpointLoadMethod.LOAD_TYPE_FUNC.items():       # built-ins + user-registered types
            self.defalut_load_type_func.setdefault(ckpt_load_type, fn)
        if self.stop_file_path and gpc.get_global_rank() == 0:
            dir_path = os.path.dirname(self.stop_file_path)
            if dir_path not in ("", ".") and not os.path.exists(dir_path):
                os.makedirs(dir_path, exist_ok=True)
            open(self.stop_file_path, "a").close()
        self.ckpt_quit_signal_handled = False
        # auto-resume takes precedence over load_ckpt_info
        auto_resume = ckpt_config.get("auto_resume", None)
        if auto_resume is None:                  # old-style: load_given_ckpt=True pins the named folder
            auto_resume = auto_resume_sanity_check(ckpt_config)
        if auto_resume and self.save_ckpt_folder:
            latest = self.query_lastest_ckpt()
            if latest is not None:
                self.load_ckpt_info = dict(path=latest, content=("all",), ckpt_type="internevo")
                if gpc.is_rank_for_log():
                    logger.info(f"auto_resume: latest checkpoint is {latest}")
        if self.load_ckpt_info is not None:
            self.load_ckpt_info = self._normalise_load_info(self.load_ckpt_info)
        torch.cuda.empty_cache() if torch.cuda.is_available() else None

    @staticmethod
    def _normalise_load_info(info):
        info = dict(info)
        assert "path" in info and "content" in info, "load_ckpt_info needs 'path' and 'content'"
        info.setdefault("ckpt_type", "internevo")
        if isinstance(info["content"], str):
            info["content"] = (info["content"],)
        info["content"] = CheckpointLoadMask(tuple(info["content"]))
        t = info["ckpt_type"]
        info["ckpt_type"] = {"internlm": CheckpointLoadType.INTERNLM, "internevo": CheckpointLoadType.INTERNEVO}.get(t, t)
        return info

    # ------------------------------------------------------------------------------------------------------------
    def quit_signal_handler(self, train_state) -> bool:
        """Stop-file protocol (reference ``:331-377``): rank 0 reads an integer N from the stop file and broadcasts
        it; ``N > 0`` = save at step N and quit, ``N < 0`` = save at step |N| and continue, ``0`` = nothing."""
        now_break, now_save_ckpt, save_type = False, False, CheckpointSaveType.NORMAL_CHECKPOINT
        if self.stop_file_path is None:
            return now_break, now_save_ckpt, save_type
        signal = 0
        if gpc.get_global_rank() == 0:
            try:
                with open(self.stop_file_path, "r+", encoding="utf-8") as f:
                    txt = f.read().strip()
                    signal = int(txt) if txt not in ("",) else 0
            except (OSError, ValueError):
                signal = 0
        if gpc.is_distributed and gpc.get_world_size(ParallelMode.GLOBAL) > 1:
            t = torch.tensor([signal], device=get_current_device(), dtype=torch.int64)
            dist.broadcast(t, src=0)
            signal = int(t.item())
        # The request fires at exactly step |N| (reference ``:355-371``); rank 0 then rewrites the file to 0, which also
        # re-arms the protocol: a later request in the same run (e.g. "-100" then "200") is honoured again.
        if signal != 0 and train_state.step_count == abs(signal):
            now_save_ckpt = True
            now_break = signal > 0
            if gpc.get_global_rank() == 0:
                with open(self.stop_file_path, "w", encoding="utf-8") as f:
                    f.write("0")
                msg = "Stop file: saving a checkpoint" + (" and quitting" if now_break else "")
                logger.warning(msg + f" at step {train_state.step_count}")
                send_alert_message(address=self.feishu_address, message=msg)
        elif signal != 0 and train_state.step_count > abs(signal) and gpc.get_global_rank() == 0 \
                and not self.ckpt_quit_signal_handled:
            self.ckpt_quit_signal_handled = True   # warn once: the requested step has already passed
            logger.warning(f"Stop file asks for step {abs(signal)} but training is at step {train_state.step_count}; ignored")
        return now_break, now_save_ckpt, save_type

    def is_now_to_save_ckpt(self, train_state, force=False) -> (bool, CheckpointSaveType, bool):
        save_ckpts, save_type, now_break = False, CheckpointSaveType.NORMAL_CHECKPOINT, False
        if force:
            return True, save_type, now_break
        if self.oss_snapshot_freq not in (None, float("inf")) and self.oss_snapshot_freq > 1 and \
                train_state.step_count % self.oss_snapshot_freq == 0:
            save_ckpts, save_type = True, CheckpointSaveType.SNAPSHOT_CHECKPOINT
        if train_state.step_count % self.checkpoint_every == 0 or train_state.step_count == train_state.total_steps:
            save_ckpts, save_type = True, CheckpointSaveType.NORMAL_CHECKPOINT
        now_break, singal_save_ckpts, singal_save_type = self.quit_signal_handler(train_state)
        if save_ckpts is False:
            save_ckpts = singal_save_ckpts
            save_type = singal_save_type
        return save_ckpts, save_type, now_break

    def try_save_checkpoint(self, train_state, force=False):
        if not self.enable_save_ckpt:
            return False
        save_ckpts, save_type, now_break = self.is_now_to_save_ckpt(train_state, force=force)
        if save_ckpts:
            self.storage_manager.wait()  # previous asynchronous upload
            if save_type == CheckpointSaveType.SNAPSHOT_CHECKPOINT:
                self.snapshot_counter = (getattr(self, "snapshot_counter", -1) + 1) % 2
                save_ckpt_folder = os.path.join(self.snapshot_ckpt_folder, f"{self.snapshot_counter}")
            else:
                save_ckpt_folder = os.path.join(self.save_ckpt_folder, str(train_state.step_count))
            self.save_checkpoint(folder=save_ckpt_folder, model=self.model, optimizer=self.optimizer,
                                 scheduler=self.lr_scheduler, train_state=train_state, model_config=self.model_config,
                                 model_config_file=self.model_config_file)
        return now_break

    def wait_async_upload_finish(self):
        self.storage_manager.wait()
        if gpc.is_distributed:
            dist.barrier()

    # ------------------------------------------------------------------------------------------------------------
    def query_latest_snapshot_step_boto3(self):
        return self.query_latest_snapshot_step_local()

    def query_latest_snapshot_step_local(self):
        """→ ``(path, step)`` of the newest complete checkpoint among normal folders and the two snapshot folders."""
        best_path, best_step = None, -1
        backend, root = try_get_storage_backend(self.save_ckpt_folder)
        prefix = "" if backend == "local" else backend + ":"

        def scan(folder, take_folder_name):
            nonlocal best_path, best_step
            try:
                names = get_fns(prefix + folder)
            except Exception:
                return
            for n in names:
                if take_folder_name:
                    sub = os.path.join(folder, n)
                    if n == "snapshot":
                        continue
                    try:
                        fns = get_fns(prefix + sub)
                    except Exception:
                        continue
                else:
                    sub, fns = folder, names
                for fn in fns:
                    if fn.endswith(".step"):
                        step = int(fn.split(".")[0])
                        if step > best_step:
                            best_step, best_path = step, prefix + sub
                if not take_folder_name:
                    break

        scan(root, True)
        for i in (0, 1):
            scan(os.path.join(root, "snapshot", str(i)), False)
        return best_path, best_step

    def query_lastest_ckpt(self):
        latest = None
        if gpc.get_global_rank() == 0:
            latest, step = self.query_latest_snapshot_step_local()
            if latest is None:
                logger.warning(f"No checkpoint found under {self.save_ckpt_folder}; training starts from scratch")
        if gpc.is_distributed and gpc.get_world_size(ParallelMode.GLOBAL) > 1:
            obj = [latest]
            dist.broadcast_object_list(obj, src=0)
            latest = obj[0]
        return latest

    def try_resume_training(self, train_state: TrainState, current_time=""):
        if self.load_ckpt_info is None:
            if gpc.is_rank_for_log():
                logger.info(f"===========New Run {current_time} on host:{os.uname().nodename},rank={gpc.get_global_rank()},"
                            f"tp={gpc.get_local_rank(ParallelMode.TENSOR)},pp={gpc.get_local_rank(ParallelMode.PIPELINE)},"
                            f"dp={gpc.get_local_rank(ParallelMode.DATA)}===========")
            return
        load_path = self.load_ckpt_info["path"]
        load_type = self.load_ckpt_info["ckpt_type"]
        load_func = self.defalut_load_type_func[load_type]
        load_content_str = load_func(self, self.load_ckpt_info, train_state)
        # whatever loader ran (user-registered types included): after a weights-only load the optimizer's fp32 master follows
        # the new weights (idempotent; reference ``checkpoint_manager.py:553-557``)
        content = self.load_ckpt_info["content"]
        if content.only_load(CheckpointLoadContent.MODEL) and hasattr(self.optimizer, "reload_zero_fp32_buff"):
            self.optimizer.reload_zero_fp32_buff()
        if gpc.is_rank_for_log():
            logger.info(f"===========Resume training from `{load_path}` {current_time}, loaded: {load_content_str}===========")
            if train_state is not None:
                logger.info(f"resume at step {train_state.step_count}, tokens {train_state.num_consumed_tokens}")

    @llm_timeout(func_name="save_checkpoint")
    def save_checkpoint(self, folder, model, optimizer, scheduler, train_state: TrainState, model_config: Dict = None,
                        model_config_file: str = None):
        start = timer("save-model")
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if gpc.is_distributed:
            dist.barrier()
        if gpc.is_rank_for_log():
            logger.info(f"Saving checkpoint to `{folder}` at batch count:{train_state.step_count}...")
        timer("save-model").start()
        save_model_checkpoint(folder=folder, model=model)
        timer("save-model").stop()
        timer("save-optimizer").start()
        save_optimizer_checkpoint(optim=optimizer, state_path=folder)
        timer("save-optimizer").stop()
        if (hasattr(train_state, "data_state_dict") and gpc.get_local_rank(ParallelMode.TENSOR) == 0
                and gpc.get_local_rank(ParallelMode.PIPELINE) == 0):
            llm_save(os.path.join(folder, f"sampler_{gpc.get_local_rank(ParallelMode.DATA)}.pt"),
                     saved_obj=train_state.data_state_dict)
        if gpc.is_rank_for_log():
            if scheduler:
                states = scheduler.state_dict()
                if gpc.config.get("ckpt", {}).get("optimizer_ckpt_format", "internevo_b200") == "reference":
                    from internevo_b200.utils.parallel import is_using_isp

                    from .optimizer_interchange import reference_scheduler_state

                    # the reference keeps its empty groups: default, [embed_head], fp32, [experts]
                    n_groups = 2 + int(is_using_isp()) + len(gpc.expert_parallel_group_names)
                    states = reference_scheduler_state(scheduler, n_groups)
                llm_save(os.path.join(folder, "schedulder.pt"), saved_obj=states)
            if hasattr(train_state, "batch_sampler") and train_state.batch_sampler is not None:
                llm_save(os.path.join(folder, "sampler.pt"), saved_obj=train_state.batch_sampler.state_dict())
            llm_save(os.path.join(folder, "context.pt"), saved_obj=train_state.state_dict())
            if model_config is not None:
                cfg = dict(model_config)
                cfg["dtype"] = str(cfg.get("dtype"))
                llm_save(os.path.join(folder, "model_config.pt"), saved_obj=cfg)
            if model_config_file is not None:
                llm_save(os.path.join(folder, "config_file.pt"), saved_obj=model_config_file)
        if gpc.is_distributed:
            dist.barrier()
        marker = os.path.join(folder, f"{train_state.step_count}.step")
        if gpc.is_rank_for_log():
            if self.async_upload:
                get_storage_manager().set_pending_marker(marker)
            else:
                get_storage_manager()._client(marker)[0].upload_bytes(b"", try_get_storage_backend(marker)[1])
            logger.info(f"Step: {train_state.step_count}, rank 0 save ckpt use {timer('save-model').elapsed(False):.3f}s "
                        f"(+ optimizer {timer('save-optimizer').elapsed(False):.3f}s)")
        del start

    def set_save_folder(self, folder, step) -> None:
        """Tell the storage manager where the newest checkpoint lives (asynchronous uploads report against it)."""
        sm = get_storage_manager()
        sm.latest_save_folder, sm.latest_save_step = folder, step

    def try_ping_storage(self):
        if gpc.is_rank_for_log() and self.save_ckpt_folder:
            p = os.path.join(self.save_ckpt_folder, "ping.pt")
            llm_save(p, saved_obj={"ping": 1})
            assert llm_load(p)["ping"] == 1
            get_storage_manager().delete_obj(p)
