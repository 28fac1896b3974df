"""Trainer façade + TrainState (reference ``internlm/core/trainer.py:20-213``)."""
from __future__ import annotations

import json
from collections import deque
from typing import Iterable, Optional

from internevo_b200.core.engine import Engine
from internevo_b200.core.scheduler.base_scheduler import BaseScheduler
from internevo_b200.core.scheduler.no_pipeline_scheduler import NonPipelineScheduler
from internevo_b200.core.scheduler.pipeline_scheduler import InterleavedPipelineScheduler, PipelineScheduler


class TrainState:
    """Step / token counters, TGS statistics and the resumable sampler anchor."""

    def __init__(self, config, batch_sampler) -> None:
        self.batch_count = 0
        self.num_consumed_samples_in_epoch = 0
        self.num_consumed_tokens = 0
        self.inf_nan_skip_batches = 0
        self.step_count = 0
        self.total_steps: int = config.data.total_steps
        self.resume_tb_folder = config.get("resume_tb_folder", None)
        self.tensorboard_folder = config.get("tensorboard_folder", None)
        self.lr = config.adam.lr
        self.init_batch_sampler(batch_sampler)
        self.tgs_statistic = {"sum_step": 0, "sum_tg": 0, "sum_time": 0, "sum_last_tg_10": 0, "sum_last_time_10": 0,
                              "sum_last_tg_50": 0, "sum_last_time_50": 0, "SMA_tg_50": 0, "SMA_time_50": 0,
                              "SMA_tg_50_list": deque(), "SMA_time_50_list": deque(), "sum_tgs": 0, "last_tgs_10": 0,
                              "last_tgs_50": 0}

    def init_batch_sampler(self, batch_sampler):
        """A *copy* of the sampler is the resume anchor: it advances only with fully processed batches."""
        self.batch_sampler = batch_sampler.copy() if batch_sampler is not None else None
        self.batch_sampler_iter = iter(self.batch_sampler) if batch_sampler is not None else None

    def __str__(self) -> str:
        return json.dumps({"batch_count": self.batch_count, "num_consumed_samples_in_epoch":
                           self.num_consumed_samples_in_epoch, "num_consumed_tokens": self.num_consumed_tokens,
                           "inf_nan_skip_batches": self.inf_nan_skip_batches}, indent=4, sort_keys=True)

    def load_state_dict(self, other_stuffs):
        self.num_consumed_samples_in_epoch = other_stuffs["num_consumed_samples_in_epoch"]
        self.num_consumed_tokens = other_stuffs["num_consumed_tokens"]
        self.inf_nan_skip_batches = other_stuffs["inf_nan_skip_batches"]
        self.batch_count = other_stuffs["batch_count"] + 1  # resume from the next batch
        # step_count counts COMPLETED optimizer steps and the checkpoint is written after it was incremented
        # (reference trainer.py:114-118): restore it as saved
        self.step_count = other_stuffs.get("step_count", other_stuffs["batch_count"] + 1)
        if self.resume_tb_folder is None:
            self.resume_tb_folder = other_stuffs.get("tensorboard_folder", None)

    def state_dict(self):
        return {"batch_count": self.batch_count, "num_consumed_samples_in_epoch": self.num_consumed_samples_in_epoch,
                "num_consumed_tokens": self.num_consumed_tokens, "inf_nan_skip_batches": self.inf_nan_skip_batches,
                "step_count": self.step_count, "tensorboard_folder": self.tensorboard_folder}


class Trainer:
    def __init__(self, engine: Engine, schedule: Optional[BaseScheduler] = None):
        self._engine = engine
        if schedule is None:
            schedule = NonPipelineScheduler()
        assert isinstance(schedule, BaseScheduler)
        self._schedule = schedule
        self._schedule.pre_processing(self._engine)

    @property
    def engine(self):
        return self._engine

    @property
    def schedule(self):
        return self._schedule

    @property
    def uses_pipeline(self):
        return isinstance(self._schedule, (PipelineScheduler, InterleavedPipelineScheduler))

    def train(self):
        self._engine.train()

    def eval(self):
        self._engine.eval()

    def zero_grad(self):
        self._engine.zero_grad()

    def step(self):
        return self._engine.step()

    def execute_schedule(self, data_iter: Iterable, **kwargs):
        """→ ``(output, label, loss)`` (+ moe_loss for MoE models)."""
        return self._schedule.forward_backward_step(self._engine, data_iter, **kwargs)
