from .batch_sampler import DataParallelSampler, StaticBatchSampler, get_dpsampler_dataloader
from .build_dataloader import build_train_loader_with_data_type, build_valid_loader_with_data_type
from .collaters import jsonl_ds_collate_fn, packed_collate_fn
from .datasets import (
    JsonlDataset,
    PackedDatasetWithCut,
    PackedDatasetWithoutCuSeqlen,
    RandomDataset,
    get_packed_dataset_without_short_length,
    unpack_data,
)

__all__ = [
    "DataParallelSampler", "StaticBatchSampler", "get_dpsampler_dataloader", "build_train_loader_with_data_type",
    "build_valid_loader_with_data_type", "jsonl_ds_collate_fn", "packed_collate_fn", "JsonlDataset",
    "PackedDatasetWithCut", "PackedDatasetWithoutCuSeqlen", "RandomDataset",
    "get_packed_dataset_without_short_length", "unpack_data",
]
