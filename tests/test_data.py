"""Packed dataset / sampler behaviour (reference semantics: ``tests/test_data/test_batch_sampler.py`` and the
``PackedDatasetWithCut`` docstring example)."""
import numpy as np
import torch

from common import run_distributed
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


# --------------------------------------------------------------------------------------------- un-packed (no flash-attention) mode
def test_unpack_data_rows_and_ignored_padding():
    """``build_unpack`` packs ``micro_bsz`` whole samples + padding into one row; ``unpack_data`` hands the samples back as
    ``[micro_bsz, seq_len]`` rows: short samples, a padding segment LONGER than ``seq_len`` and label padding with -100."""
    from internevo_b200.core.context import Config, global_context as gpc
    from internevo_b200.data.datasets import unpack_data

    gpc.set_config(Config(dict(data=dict(micro_bsz=2, seq_len=8, packed_length=16), model=dict(use_flash_attn=False))))
    ids = torch.arange(1, 17).reshape(1, 16)
    cu = torch.tensor([[0, 3, 5, 16]], dtype=torch.int32)            # samples of 3 and 2 tokens, then 11 padding positions
    rows = unpack_data(ids, cu)
    assert rows.shape == (2, 8)
    assert rows[0].tolist() == [1, 2, 3, 0, 0, 0, 0, 0] and rows[1].tolist() == [4, 5, 0, 0, 0, 0, 0, 0]
    lab = unpack_data(ids, cu, padding_v=-100)
    assert lab[1].tolist() == [4, 5] + [-100] * 6
    tids = unpack_data(torch.ones(3, 16, dtype=torch.long), cu.repeat(3, 1), is_type_ids=True)
    assert tids.shape == (3, 2, 8) and int(tids.sum()) == 3 * (3 + 2)


def _unpacked_vs_packed(rank, world, unpacked):
    """One step on the same two samples: packed with ``cu_seqlens`` (flash path) or un-packed ``[micro_bsz, seq_len]`` rows."""
    from common import build_trainer, tiny_config

    cfg = tiny_config(num_layers=2, micro_num=1, micro_bsz=2, seq_len=16, use_flash_attn=not unpacked)
    cfg["data"]["use_packed_dataset"] = not unpacked
    trainer, opt, model, _ = build_trainer(cfg)
    g = torch.Generator().manual_seed(5)
    s0, s1 = torch.randint(1, 100, (9,), generator=g), torch.randint(1, 100, (12,), generator=g)

    def labels_of(s):
        return torch.cat([s[1:], torch.tensor([-100])])

    if unpacked:     # what PackedDatasetWithCut.build_unpack + packed_collate_fn produce: samples, then zero padding
        pad = 32 - 21
        ids = torch.cat([s0, s1, torch.zeros(pad, dtype=torch.long)])[None]
        lab = torch.cat([labels_of(s0), labels_of(s1), torch.zeros(pad, dtype=torch.long)])[None]
        cu = torch.tensor([[0, 9, 21, 32]], dtype=torch.int32)
        idx = torch.cat([torch.arange(9), torch.arange(12), torch.arange(pad)])[None]
    else:            # packed row of exactly the two samples' tokens, padded by a third ignored segment
        pad = 32 - 21
        ids = torch.cat([s0, s1, torch.ones(pad, dtype=torch.long)])[None]
        lab = torch.cat([labels_of(s0), labels_of(s1), torch.full((pad,), -100)])[None]
        cu = torch.tensor([[0, 9, 21, 32]], dtype=torch.int32)
        idx = torch.cat([torch.arange(9), torch.arange(12), torch.arange(pad)])[None]
    data = {"input_ids": ids, "cu_seqlens": cu, "indexes": idx}
    trainer.zero_grad()
    out = trainer.execute_schedule((data, lab), forward_only=False, return_loss=True, return_output_label=False)
    ok, norms = trainer.step()
    assert ok
    return float(out[2]), {k: float(v) for k, v in norms.items()}


def test_unpacked_mode_matches_the_packed_loss_and_gradient_norm():
    packed = run_distributed(_unpacked_vs_packed, 1, False)[0]
    plain = run_distributed(_unpacked_vs_packed, 1, True)[0]
    assert abs(packed[0] - plain[0]) < 1e-4, (packed, plain)                       # same tokens, same targets, padding ignored
    for k in packed[1]:
        assert abs(packed[1][k] - plain[1][k]) < 1e-3 * max(1.0, packed[1][k]), (packed, plain)


# --------------------------------------------------------------------------------------------- native corpus scan / token parser
def test_native_dataio_matches_the_json_decoder(tmp_path):
    """``csrc/dataio.cpp`` through ctypes: the .meta table of a shard and every sample are what ``json`` gives; lines in another form
    (extra keys, floats, empty) are handed back to the JSON decoder instead of being mis-parsed."""
    import json

    from internevo_b200.data import _native
    from internevo_b200.data.datasets import JsonlDataset

    assert _native.lib() is not None, "internevo_b200/_dataio.so is built by csrc/build.py (g++ only)"
    rng = np.random.RandomState(0)
    rows = [rng.randint(-50000, 90000, size=int(n)).tolist() for n in rng.randint(0, 300, size=200)]
    rows[3] = []
    path = tmp_path / "part0.bin"
    with open(path, "wb") as f:
        for i, r in enumerate(rows):
            sep = (", ", ",", " , ")[i % 3]
            f.write(("{" + '"tokens": [' + sep.join(map(str, r)) + "]}" + ("\n" if i != len(rows) - 1 else "")).encode())
    table = _native.scan_jsonl(str(path))
    assert table.shape == (200, 2) and table[:, 1].tolist() == [len(r) for r in rows] and int(table[0, 0]) == 0
    raw = open(path, "rb").read()
    for i in (0, 3, 57, 199):
        end = raw.find(b"\n", int(table[i, 0]))
        line = raw[int(table[i, 0]): end if end >= 0 else len(raw)]
        assert json.loads(line)["tokens"] == rows[i] and _native.parse_tokens(line).tolist() == rows[i]
    ds = JsonlDataset(str(path), dataset_type_id=2, min_length=0)            # no .meta file: built by the native scan
    assert len(ds) == 200 and ds.num_tokens == sum(len(r) for r in rows)
    item = ds[57]
    assert item["tokens"] == rows[57] and item["length"] == len(rows[57]) and item["type_id"] == 2
    # other shapes fall back to json (parse) / to the Python scan (meta)
    for odd in (b'{"tokens": [1, 2], "type_id": 5}', b'{"tokens": [1.5]}', b'{"text": "x"}', b'', b'{"tokens": [1,]}',
                b'{"tokens": [12345678901234567890]}'):
        assert _native.parse_tokens(odd) is None, odd
    assert _native.parse_tokens(b' { "tokens" : [ -7 , 8 ] } \n').tolist() == [-7, 8]
    mixed = tmp_path / "mixed.bin"
    mixed.write_bytes(b'{"tokens": [1, 2, 3]}\n{"tokens": [4], "note": "extra key"}\n')
    assert _native.scan_jsonl(str(mixed)) is None
    ds2 = JsonlDataset(str(mixed), min_length=0)
    assert ds2.offsets[:, 1].tolist() == [3, 1] and ds2[1]["tokens"] == [4] and ds2[1]["note"] == "extra key"
