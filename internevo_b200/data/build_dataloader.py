"""Train / validation dataloader construction (reference ``internlm/data/build_dataloader.py:88-158``)."""
from __future__ import annotations

from functools import partial

import torch
from torch.utils.data import ConcatDataset, DataLoader

from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.parallel import is_using_isp

from .batch_sampler import StaticBatchSampler, get_dpsampler_dataloader
from .collaters import jsonl_ds_collate_fn, packed_collate_fn
from .datasets import (
    get_dataset_type_ids_map,
    JsonlDataset,
    PackedDatasetWithCut,
    PackedDatasetWithoutCuSeqlen,
    RandomDataset,
    get_dataset_dict,
    get_packed_dataset_without_short_length,
)

logger = get_logger(__file__)


def _dp_rank_size():
    mode = ParallelMode.WEIGHT_DATA if is_using_isp() else ParallelMode.DATA
    # ISP shards the sequence over the TENSOR group: ranks of one TP group consume the same samples
    if is_using_isp():
        return gpc.get_local_rank(ParallelMode.DATA), gpc.get_world_size(ParallelMode.DATA)
    return gpc.get_local_rank(mode), gpc.get_world_size(mode)


def get_tokenized_train_loader_items(data_cfg):
    if data_cfg.get("train_folder", None) is None:
        train_ds = RandomDataset(num_samples=data_cfg.get("num_random_samples", 100000),
                                 max_len=data_cfg.seq_len, fixed_seqlen=data_cfg.fixed_random_dataset_seqlen)
        if data_cfg.pack_sample_into_one:
            train_ds = PackedDatasetWithoutCuSeqlen(train_ds, max_length_per_sample=data_cfg.seq_len,
                                                    packed_length=data_cfg.packed_length)
        else:
            train_ds = PackedDatasetWithCut(train_ds, max_length_per_sample=data_cfg.seq_len,
                                            packed_length=data_cfg.packed_length)
    else:
        train_ds = get_packed_dataset_without_short_length(
            folder=data_cfg.train_folder, packed_length=data_cfg.packed_length, max_length_per_sample=data_cfg.seq_len,
            show_progress=gpc.is_rank_for_log(), min_length=data_cfg.get("min_length", 0),
            min_length_dict=data_cfg.get("min_length_dict", None), pack_sample_into_one=data_cfg.pack_sample_into_one,
        )
    rank, size = _dp_rank_size()
    train_sampler = StaticBatchSampler(
        train_ds.datasets if isinstance(train_ds, ConcatDataset) else [train_ds],
        batch_size=data_cfg.micro_num, rampup_batch_size=data_cfg.rampup_batch_size, micro_bsz=data_cfg.micro_bsz,
        seed=data_cfg.get("seed", 1024), drop_last=True, data_rank=rank, data_world_size=size,
    )
    train_collate_fn = partial(packed_collate_fn, packed_length=data_cfg.packed_length)
    return train_ds, train_sampler, train_collate_fn


def get_tokenized_valid_loader_items(data_cfg):
    """→ ``({name: dataset}, collate_fn)``: one entry per sub-folder of ``valid_folder`` (reference ``:69-85``)."""
    if not data_cfg.get("valid_folder", None):
        valid_ds = RandomDataset(num_samples=gpc.get_world_size(ParallelMode.DATA) * 500, max_len=data_cfg.seq_len,
                                 fixed_seqlen=data_cfg.get("fixed_random_dataset_seqlen", False))
    else:
        # samples shorter than ``data.valid_min_length`` tokens are left out; 50 unless configured, as in the reference
        # (its ``JsonlDataset`` default, ``tokenized/dataset.py:9-56``)
        valid_ds = get_dataset_dict(folder=data_cfg.valid_folder, split="", min_length=data_cfg.get("valid_min_length", 50))
    if not isinstance(valid_ds, dict):
        valid_ds = {"val": valid_ds}
    return valid_ds, partial(jsonl_ds_collate_fn, max_length_per_sample=data_cfg.seq_len)


def build_train_loader_with_data_type():
    """→ ``(train_dataloader, dataset_types)``; only ``data.type == "tokenized"`` exists in this snapshot."""
    data_cfg = gpc.config.data
    assert data_cfg.type == "tokenized", f"unsupported data type {data_cfg.type}"
    train_ds, train_sampler, train_collate_fn = get_tokenized_train_loader_items(data_cfg)
    # names of the per-type training metrics: the sub-folders of the training folder in sorted order, the order of their type ids
    train_folder = data_cfg.get("train_folder", None)
    dataset_types = list(get_dataset_type_ids_map(train_folder).keys()) if train_folder else ["en", "cn", "code"]
    # ``data.num_worker`` loader processes pack the next batches while the GPU runs the current step: 4 unless configured, as in
    # the reference (``data/build_dataloader.py:107-110``); a CPU-only host (plumbing / demo runs) loads in the main process
    workers = data_cfg.get("num_worker", 4 if torch.cuda.is_available() else 0)
    train_dl = DataLoader(
        dataset=train_ds, batch_sampler=train_sampler, num_workers=workers, pin_memory=torch.cuda.is_available(),
        collate_fn=train_collate_fn, persistent_workers=workers > 0,
    )
    return train_dl, dataset_types


def build_valid_loader_with_data_type():
    """``{name: DataLoader}``; the batch is clamped to what every data-parallel rank can fill, sets too small for a single
    micro-batch per rank are skipped (reference ``:116-157``)."""
    data_cfg = gpc.config.data
    assert data_cfg.type == "tokenized", f"unsupported data type {data_cfg.type}"
    valid_ds, valid_collate_fn = get_tokenized_valid_loader_items(data_cfg)
    if valid_ds is None:
        return None
    val_dls = {}
    for name, ds in valid_ds.items():
        bsz = min(data_cfg.valid_micro_num * data_cfg.micro_bsz, len(ds) // gpc.get_world_size(ParallelMode.DATA))
        bsz = bsz // data_cfg.micro_bsz * data_cfg.micro_bsz
        if bsz == 0:
            if gpc.is_rank_for_log():
                logger.info(f"skip validate {name}.")
            continue
        val_dls[name] = get_dpsampler_dataloader(ds, shuffle=False, num_workers=data_cfg.get("num_worker", 0),
                                                 batch_size=bsz, collate_fn=valid_collate_fn, drop_last=True)
        if gpc.is_rank_for_log():
            logger.info(f"load validation dataset {name} with valid batch size {bsz} and samples {len(val_dls[name])}.")
    return val_dls
