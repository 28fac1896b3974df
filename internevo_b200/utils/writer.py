"""TensorBoard writer with a step-batched queue (reference ``internlm/utils/writer.py:17-184``): one writer per rank,
folder name agreed by broadcast from rank 0; scalars are queued and flushed every ``queue_max_length`` steps."""
from __future__ import annotations

import logging
import os
import socket
import sys
import traceback
from functools import partial

import torch
import torch.distributed as dist

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc

try:
    from torch.utils.tensorboard import SummaryWriter
except Exception:  # pragma: no cover
    SummaryWriter = None


def tb_save_run_info(writer, config_lines, global_step=0):
    writer.add_text(tag="cmd", text_string=" ".join(sys.argv[:]), global_step=global_step)
    writer.add_text(tag="hostname", text_string=socket.gethostname(), global_step=global_step)
    writer.add_text(tag="config", text_string="  \n".join(config_lines), global_step=global_step)


def init_tb_writer(job_name: str, launch_time: str, file_name: str, tensorboard_folder: str, resume_tb_folder: str,
                   step_count: int, config: str, logger: logging.Logger):
    tb_log_file_name = file_name
    if not tensorboard_folder:
        tb_folder = os.path.join(job_name, launch_time, "tensorboards")
    else:
        tb_folder = tensorboard_folder
    if gpc.get_global_rank() == 0:
        if resume_tb_folder is not None and os.path.exists(resume_tb_folder) and resume_tb_folder != tb_folder:
            os.makedirs(tb_folder, exist_ok=True)
            os.system(f"cp -r {resume_tb_folder}/* {tb_folder}/")
            os.system(f"chmod -R +w {tb_folder}/")
        else:
            os.makedirs(tb_folder, exist_ok=True)
    if gpc.is_distributed and gpc.get_world_size(ParallelMode.GLOBAL) > 1:
        tb_folders = [tb_folder]
        dist.broadcast_object_list(tb_folders, src=0)
        tb_folder = tb_folders[0]
    if gpc.get_local_rank(ParallelMode.TENSOR) == 0 and gpc.get_local_rank(ParallelMode.PIPELINE) in (
        0, gpc.get_world_size(ParallelMode.PIPELINE) - 1
    ):
        tb_logdir = os.path.join(tb_folder, tb_log_file_name)
        writer = SummaryWriter(log_dir=tb_logdir, max_queue=5, purge_step=step_count, flush_secs=3)
        writer.add_text(tag="job_name", text_string=job_name, global_step=step_count)
        writer.add_text(tag="tensorboard_folder", text_string=tb_logdir, global_step=step_count)
        if config is not None:
            tb_save_run_info(writer, [f"{k}: {v}" for k, v in dict(config).items()], step_count)
    else:
        writer = None
        tb_logdir = tb_folder
    if gpc.is_rank_for_log():
        logger.info(f"Launch time: {launch_time}, tensorboard folder: {tb_folder}")
    return writer, tb_logdir


class Writer:
    def __init__(self, job_name: str = None, launch_time: str = None, file_name: str = None,
                 tensorboard_folder: str = None, resume_tb_folder: str = None, step_count: int = 0, config: str = None,
                 logger: logging.Logger = None, enable_tb: bool = True, queue_max_length: int = 1,
                 total_steps: int = None) -> None:
        self.enable_tb = enable_tb and SummaryWriter is not None
        self.tb_writer, self.tb_logdir = None, None
        if self.enable_tb:
            self.tb_writer, self.tb_logdir = init_tb_writer(job_name, launch_time, file_name, tensorboard_folder,
                                                            resume_tb_folder, step_count, config, logger)
        self.queue_max_length = max(1, queue_max_length)
        self.total_steps = total_steps
        self.queue = []

    def _flush(self):
        for fn in self.queue:
            fn()
        self.queue = []

    def _add(self, fn, step):
        self.queue.append(fn)
        if len(self.queue) >= self.queue_max_length or (self.total_steps and step >= self.total_steps - 1):
            self._flush()

    def add_scalar(self, key, value, step):
        try:
            if self.enable_tb and self.tb_writer is not None:
                if torch.is_tensor(value):
                    value = value.item()
                self._add(partial(self.tb_writer.add_scalar, tag=key, scalar_value=value, global_step=step), step)
        except Exception:  # pragma: no cover
            traceback.print_exc()

    def add_scalars(self, key, value, step):
        try:
            assert isinstance(value, dict)
            if self.enable_tb and self.tb_writer is not None:
                self._add(partial(self.tb_writer.add_scalars, main_tag=key, tag_scalar_dict=value, global_step=step), step)
        except Exception:  # pragma: no cover
            traceback.print_exc()

    def add_text(self, key, value, step):
        try:
            if self.enable_tb and self.tb_writer is not None:
                self.tb_writer.add_text(tag=key, text_string=value, global_step=step)
        except Exception:  # pragma: no cover
            traceback.print_exc()

    def close(self):
        self._flush()
        if self.tb_writer is not None:
            self.tb_writer.close()
