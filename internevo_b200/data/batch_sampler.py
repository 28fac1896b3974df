"""Batch samplers (reference ``internlm/data/tokenized/batch_sampler.py``): rank-strided static batches with batch-size
ramp-up and a resumable state; ``DataParallelSampler`` for validation."""
from __future__ import annotations

import math
import random
from typing import Iterator, TypeVar

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Sampler

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc

T_co = TypeVar("T_co", covariant=True)


class DataParallelSampler(Sampler):
    """Shards a dataset over the DATA group, optional epoch-seeded shuffle (reference ``:21-107``)."""

    def __init__(self, dataset: Dataset, shuffle: bool = False, seed: int = 0, drop_last: bool = False) -> None:
        self.dataset = dataset
        self.num_replicas = gpc.get_world_size(ParallelMode.DATA)
        self.rank = gpc.get_local_rank(ParallelMode.DATA)
        self.epoch = 0
        self.drop_last = drop_last
        if self.drop_last and len(self.dataset) % self.num_replicas != 0:
            self.num_samples = math.ceil((len(self.dataset) - self.num_replicas) / self.num_replicas)
        else:
            self.num_samples = math.ceil(len(self.dataset) / self.num_replicas)
        self.total_size = self.num_samples * self.num_replicas
        self.shuffle = shuffle
        self.seed = seed

    def __iter__(self) -> Iterator[T_co]:
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            indices = torch.randperm(len(self.dataset), generator=g).tolist()
            self.epoch += 1
        else:
            indices = list(range(len(self.dataset)))
        if not self.drop_last:
            pad = self.total_size - len(indices)
            if pad <= len(indices):
                indices += indices[:pad]
            else:
                indices += (indices * math.ceil(pad / len(indices)))[:pad]
        else:
            indices = indices[: self.total_size]
        assert len(indices) == self.total_size
        indices = indices[self.rank: self.total_size: self.num_replicas]
        assert len(indices) == self.num_samples
        return iter(indices)

    def __len__(self) -> int:
        return self.num_samples

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch


class StaticBatchSampler:
    """Fixed micro-batch size, optional ``"start incr every"`` batch-size ramp-up, resumable (reference ``:110-287``)."""

    def __init__(self, datasets, batch_size=192, rampup_batch_size="6 2 8", micro_bsz=2, seed=0, drop_last=True,
                 data_rank=0, data_world_size=1):
        assert drop_last is True, "Currently only support drop last"
        if rampup_batch_size:
            start_bsz, bsz_incre, incre_every = map(int, rampup_batch_size.split())
        else:
            start_bsz, bsz_incre, incre_every = batch_size, batch_size, 1
        self.raw_rampup_batch_size = rampup_batch_size
        self.start_bsz, self.bsz_incre, self.incre_every = start_bsz, bsz_incre, incre_every
        if gpc.is_initialized(ParallelMode.PIPELINE):
            assert (batch_size - start_bsz) % bsz_incre == 0
            assert batch_size % micro_bsz == 0 and start_bsz % micro_bsz == 0 and bsz_incre % micro_bsz == 0
        self.batch_size = batch_size
        self.epoch = 0
        self.seed = seed
        self.rng = np.random.RandomState(seed)
        self.batch_count = 0
        self.micro_bsz = micro_bsz
        self.data_rank = data_rank
        self.data_world_size = data_world_size
        self.num_consumed_samples_in_epoch = 0
        self.datasets = datasets
        self.num_samples = sum(len(ds) for ds in datasets)
        self.get_indices()

    def _rampup_samples(self):
        ramp_steps = (self.batch_size - self.start_bsz) // self.bsz_incre
        if self.batch_count < ramp_steps * self.incre_every:
            return ramp_steps, sum((i * self.bsz_incre + self.start_bsz) * self.incre_every for i in range(ramp_steps))
        return ramp_steps, None

    def get_indices(self, old_indices=None):
        if old_indices is not None:
            assert len(old_indices) <= self.num_samples
        else:
            old_indices = np.array([])
        indices = np.arange(len(old_indices), self.num_samples)
        self.rng_state = self.rng.get_state()
        self.rng.shuffle(indices)
        _, ramp = self._rampup_samples()
        per = self.batch_size * self.data_world_size
        if ramp is not None:
            assert ramp * self.data_world_size <= self.num_samples, "Too much rampup samples"
            num_samples = (self.num_samples - ramp * self.data_world_size) // per * per + ramp * self.data_world_size
        else:
            num_samples = self.num_samples // per * per
        indices = np.concatenate([old_indices, indices]).astype(int)[:num_samples]
        self.indices = indices
        assert len(self.indices) >= self.batch_size, "The number of samples should be larger than batch_size"
        self.num_consumed_samples_in_epoch = 0

    def set_epoch(self, epoch):
        self.epoch = epoch
        self.rng = np.random.RandomState(self.seed + self.epoch)

    def __len__(self):
        ramp_steps, ramp = self._rampup_samples()
        if ramp is not None:
            n = (self.num_samples - ramp * self.data_world_size) // self.batch_size
            return n // self.data_world_size + self.incre_every * ramp_steps
        return self.num_samples // self.batch_size // self.data_world_size

    def __iter__(self):
        indices = self.indices[self.data_rank:: self.data_world_size]
        while self.num_consumed_samples_in_epoch < len(indices):
            cur = min((self.batch_count // self.incre_every) * self.bsz_incre + self.start_bsz, self.batch_size)
            batch = indices[self.num_consumed_samples_in_epoch: self.num_consumed_samples_in_epoch + cur]
            self.num_consumed_samples_in_epoch += len(batch)
            self.batch_count += 1
            yield batch
        self.get_indices()

    def state_dict(self):
        return {
            "batch_size": self.batch_size, "raw_rampup_batch_size": self.raw_rampup_batch_size,
            "rng_state": self.rng_state, "epoch": self.epoch, "seed": self.seed,
            "data_world_size": self.data_world_size,
            "num_consumed_samples_in_epoch": self.num_consumed_samples_in_epoch, "batch_count": self.batch_count,
            "indices": self.indices,
        }

    def load_state_dict(self, states):
        for name in ("data_world_size", "raw_rampup_batch_size", "seed"):
            assert states[name] == getattr(self, name), (name, states[name], getattr(self, name))
        self.rng.set_state(states["rng_state"])
        self.get_indices(old_indices=None)
        self.epoch = states["epoch"]
        self.batch_count = states["batch_count"]
        self.num_consumed_samples_in_epoch = states["num_consumed_samples_in_epoch"]

    def copy(self):
        c = StaticBatchSampler(self.datasets, self.batch_size, self.raw_rampup_batch_size, self.micro_bsz, self.seed,
                               drop_last=True, data_rank=self.data_rank, data_world_size=self.data_world_size)
        c.load_state_dict(self.state_dict())
        return c


def get_dpsampler_dataloader(dataset, shuffle=False, seed=1024, add_sampler=True, drop_last=False, pin_memory=False,
                             num_workers=0, **kwargs):
    """DataLoader over a ``DataParallelSampler`` with deterministic worker seeding (reference ``:290-355``)."""
    _kwargs = kwargs.copy()
    if add_sampler and gpc.is_initialized(ParallelMode.DATA) and gpc.get_world_size(ParallelMode.DATA) > 1:
        sampler = DataParallelSampler(dataset, shuffle=shuffle, drop_last=drop_last)
    else:
        sampler = None

    def seed_worker(worker_id):
        np.random.seed(seed)
        torch.manual_seed(seed)
        random.seed(seed)

    if sampler is None:
        return DataLoader(dataset, worker_init_fn=seed_worker, shuffle=shuffle, drop_last=drop_last,
                          pin_memory=pin_memory, num_workers=num_workers, **_kwargs)
    return DataLoader(dataset, sampler=sampler, worker_init_fn=seed_worker, drop_last=drop_last,
                      pin_memory=pin_memory, num_workers=num_workers, **_kwargs)
