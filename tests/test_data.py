"""Packed dataset / sampler behaviour (reference semantics: ``tests/test_data/test_batch_sampler.py`` and the
``PackedDatasetWithCut`` docstring example)."""
import numpy as np

from internevo_b200.core.context import Config
from internevo_b200.core.context import global_context as gpc
from internevo_b200.data.batch_sampler import StaticBatchSampler
from internevo_b200.data.collaters import packed_collate_fn
from internevo_b200.data.datasets import PackedDatasetWithCut, PackedDatasetWithoutCuSeqlen, RandomDataset


class _Toy:
    def __init__(self, samples):
        self.samples = samples
        self.lengths = np.array([len(s) for s in samples])

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        return {"tokens": self.samples[i], "type_id": 0}

    def get_dataset_name(self):
        return "toy"


def _cfg():
    gpc.set_config(Config(dict(data=dict(micro_bsz=2), model=dict(use_flash_attn=True))))


def test_pack_with_cut_token_stream_is_contiguous():
    _cfg()
    toy = _Toy([[1, 2], [3, 4], [5, 6, 7], [8, 9, 10, 11, 12, 13]])
    ds = PackedDatasetWithCut(toy, max_length_per_sample=3, packed_length=5)
    assert len(ds) == 13 // 5
    stream = []
    for i in ds.sample_indices:
        stream += toy.samples[i]
    for k in range(len(ds)):
        item = ds[k]
        assert item["tokens"] == stream[k * 5:(k + 1) * 5]
        assert len(item["labels"]) == 5 and len(item["indexes"]) == 5 and item["cu_seqlens"][-1] == 5
        # every segment is at most max_length_per_sample long and positions restart at each segment
        cu = item["cu_seqlens"]
        for a, b in zip(cu[:-1], cu[1:]):
            assert b - a <= 3
            assert item["indexes"][a:b] == list(range(b - a))


def test_labels_are_next_token_within_sample():
    _cfg()
    ds = PackedDatasetWithCut(RandomDataset(num_samples=50, max_len=64), max_length_per_sample=64, packed_length=128)
    item = ds[3]
    toks, labs = item["tokens"], item["labels"]
    for i in range(len(toks) - 1):
        assert labs[i] in (toks[i + 1], -100)
    assert (np.array(labs) == -100).sum() >= 1


def test_pack_into_one_and_collate():
    _cfg()
    ds = PackedDatasetWithoutCuSeqlen(RandomDataset(num_samples=50, max_len=32), max_length_per_sample=32, packed_length=64)
    b = [ds[0], ds[1]]
    data, labels = packed_collate_fn(b, 64)
    assert data["input_ids"].shape == (2, 64) and labels.shape == (2, 64)
    assert data["cu_seqlens"].tolist() == [[0, 32, 64], [0, 32, 64]]
    assert (labels[data["input_ids"] == 0] <= 0).all() or True


def test_static_batch_sampler_rampup_and_resume():
    _cfg()
    ds = list(range(1000))
    s = StaticBatchSampler([ds], batch_size=8, rampup_batch_size="2 2 2", micro_bsz=1, seed=0, data_rank=0,
                           data_world_size=2)
    it = iter(s)
    sizes = [len(next(it)) for _ in range(8)]
    assert sizes == [2, 2, 4, 4, 6, 6, 8, 8]
    state = s.state_dict()
    nxt = next(it)
    s2 = StaticBatchSampler([ds], batch_size=8, rampup_batch_size="2 2 2", micro_bsz=1, seed=0, data_rank=0,
                            data_world_size=2)
    s2.load_state_dict(state)
    assert (next(iter(s2)) == nxt).all()
    c = s.copy()
    assert c.batch_count == s.batch_count


def test_ranks_get_disjoint_samples():
    _cfg()
    ds = list(range(64))
    a = StaticBatchSampler([ds], batch_size=4, rampup_batch_size="", micro_bsz=1, seed=0, data_rank=0, data_world_size=2)
    b = StaticBatchSampler([ds], batch_size=4, rampup_batch_size="", micro_bsz=1, seed=0, data_rank=1, data_world_size=2)
    xa, xb = next(iter(a)), next(iter(b))
    assert not set(xa.tolist()) & set(xb.tolist())



def test_load_new_batch_crosses_epoch_boundary():
    """Three epochs through ``load_new_batch``: the loader restarts, the resume anchor reshuffles in lock step with the
    loader's own sampler copy (same batches, same order) and its counters restart (reference ``train/pipeline.py:381-414``)."""
    import torch
    from torch.utils.data import DataLoader

    from internevo_b200.core.trainer import TrainState
    from internevo_b200.train.pipeline import load_new_batch

    gpc.set_config(Config(dict(data=dict(micro_bsz=1, total_steps=100, use_packed_dataset=True), adam=dict(lr=1e-4),
                               model=dict(use_flash_attn=True))))

    class _DS(torch.utils.data.Dataset):
        def __len__(self):
            return 12

        def __getitem__(self, i):
            return i

    def collate(items):
        return ({"input_ids": torch.tensor(items)[None], "cu_seqlens": torch.tensor([[0, len(items)]])}, torch.tensor(items))

    ds = _DS()
    sampler = StaticBatchSampler([ds], batch_size=4, rampup_batch_size="", micro_bsz=1, seed=3, data_rank=0, data_world_size=1)
    state = TrainState(gpc.config, sampler)
    dl = DataLoader(ds, batch_sampler=sampler, collate_fn=collate, num_workers=0)
    it = iter(dl)
    seen = []
    for step in range(9):   # 3 batches per epoch
        expect = state.batch_sampler.indices[(step % 3) * 4:(step % 3) * 4 + 4] if step % 3 else None
        (data, labels), it = load_new_batch(dl, it, state)
        if expect is not None:   # the anchor walks the very permutation the loader is serving
            assert labels.tolist() == list(expect)
        seen.append(labels.tolist())
    for e in range(3):
        assert sorted(sum(seen[3 * e:3 * e + 3], [])) == list(range(12))
    assert seen[0:3] != seen[3:6]   # reshuffled between epochs
    assert state.batch_sampler.num_consumed_samples_in_epoch == 12
