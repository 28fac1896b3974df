"""Datasets: synthetic ``RandomDataset``, mmap'd tokenised JSONL, and the two packers.

Behavioural parity with the reference (``internlm/data/tokenized/{dummy_dataset,single_dataset,packed_dataset}.py``):
identical sample order, pack boundaries, ``cu_seqlens`` / ``indexes`` / shifted ``labels``.  The implementation is
different: packs are assembled with vectorised numpy over a per-dataset token index instead of per-token python loops,
so a 4k-32k token pack costs microseconds of host time (the host must not be the bottleneck in front of a B200).
"""
from __future__ import annotations

import json
import mmap
import os
import threading
from typing import Dict, List

import numpy as np
import torch
from torch.utils.data import ConcatDataset, Dataset

from internevo_b200.core.context import global_context as gpc
from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)
DEFAULT_SEED = 1024


class RandomDataset(Dataset):
    """Synthetic periodic sequences, RNG seed 1999 (reference ``dummy_dataset.py:8-49``)."""

    def __init__(self, num_samples=10000, max_len=1024, fixed_seqlen: bool = False) -> None:
        super().__init__()
        rng = np.random.RandomState(1999)
        max_num = rng.randint(1, 30, size=(num_samples,))
        rep_num = rng.randint(10, 200, size=(num_samples,))
        data, lengths = [], []
        for n, r in zip(max_num, rep_num):
            n, r = int(n), int(r)
            if fixed_seqlen:
                while n * r < max_len:
                    r *= 2
            d = np.concatenate([[n, r], np.tile(np.arange(n), r)])[:max_len].astype(np.int64)
            data.append(d)
            lengths.append(len(d))
        self.data = data
        self.max_len = max_len
        self.lengths = np.array(lengths, dtype=int)

    def __getitem__(self, index):
        return {"tokens": list(self.data[index]), "type_id": 0}

    def get_dataset_name(self):
        return "dummy_path/dummy_lang/dummy_ds/train.bin"

    def __len__(self):
        return len(self.data)


class JsonlDataset(Dataset):
    """One ``.bin`` file of JSON lines ``{"tokens": [...]}`` + ``.bin.meta`` numpy ``(offset, length)`` table, mmap'd
    (reference ``single_dataset.py:18-117``)."""

    def __init__(self, path: str, dataset_type_id: int = 0, min_length=50):
        self.path = path
        self.threadlocal = threading.local()
        resolved = os.path.realpath(path)
        self.resolved_path = resolved
        self.meta = os.path.realpath(path + ".meta")
        self.type_id = dataset_type_id
        self.offsets = np.load(self.meta) if os.path.exists(self.meta) else self._build_meta(resolved)
        self.old_length = len(self.offsets)
        if min_length > 0:
            self.offsets = self.offsets[self.offsets[:, -1] >= min_length]
        self.new_length = len(self.offsets)
        self.num_tokens = int(self.offsets[:, -1].sum()) if len(self.offsets) else 0

    @staticmethod
    def _build_meta(path):
        from . import _native

        table = _native.scan_jsonl(path)          # one native pass over the file; None: no library / unusual lines
        if table is not None:
            return table
        offs, pos = [], 0
        with open(path, "rb") as f:
            for line in f:
                offs.append((pos, len(json.loads(line)["tokens"])))
                pos += len(line)
        return np.array(offs, dtype=np.int64).reshape(-1, 2)

    def __getitem__(self, idx):
        f = self._get_mmap()
        f.seek(int(self.offsets[idx][0]))
        raw = f.readline()
        from . import _native

        toks = _native.parse_tokens(raw)          # plain {"tokens": [...]} lines never reach the JSON decoder
        if toks is not None:
            return {"tokens": toks.tolist(), "length": int(toks.shape[0]), "type_id": self.type_id}
        item = raw.decode("utf-8")
        try:
            item = json.loads(item)
            item["length"] = len(item["tokens"])
            item["type_id"] = self.type_id
        except Exception as err:
            raise json.decoder.JSONDecodeError(
                doc=self.path, pos=int(self.offsets[idx][0]),
                msg=f"Error while loading JSONL line in file {self.path} at byte {self.offsets[idx][0]}: {err}",
            )
        return item

    def get_dataset_name(self):
        return self.path

    def _get_mmap(self):
        if not hasattr(self.threadlocal, "handles"):
            with open(self.resolved_path, "rb") as f:
                mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
                self.threadlocal.handles = [f, mm]
        return self.threadlocal.handles[-1]

    def __setstate__(self, state):
        self.__dict__ = state
        self.threadlocal = threading.local()

    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if k != "threadlocal"}

    def __del__(self):
        if hasattr(self.threadlocal, "handles"):
            for h in reversed(self.threadlocal.handles):
                try:
                    h.close()
                except Exception:
                    pass

    @staticmethod
    def exists(path):
        return os.path.exists(path)

    def __len__(self):
        return len(self.offsets)


def _lengths_of(dataset) -> np.ndarray:
    if hasattr(dataset, "lengths"):
        return np.asarray(dataset.lengths)
    if hasattr(dataset, "offsets"):
        return np.asarray(dataset.offsets[:, -1])
    return np.array([len(dataset[i]["tokens"]) for i in range(len(dataset))])


class PackedDataset(Dataset):
    def __init__(self, dataset, max_length_per_sample: int = 2048, packed_length: int = 4096):
        assert hasattr(dataset, "lengths") or hasattr(dataset, "offsets")
        self.dataset = dataset
        self.max_length_per_sample = max_length_per_sample
        self.lengths = _lengths_of(dataset)
        assert len(self.lengths) == len(dataset), "dataset lengths mismatch"
        self.packed_length = packed_length
        self.seed = DEFAULT_SEED
        self.path = dataset.get_dataset_name() if hasattr(dataset, "get_dataset_name") else "unknown"
        self.num_tokens = int(self.lengths.sum())

    def get_dataset_name(self):
        return self.path

    def _shuffled(self, seed):
        rng = np.random.RandomState(seed)
        idx = np.arange(len(self.lengths))
        rng.shuffle(idx)
        lens = self.lengths[idx]
        return idx, lens, np.cumsum(lens)

    def _gather(self, start: int, end: int, label_across_cut: bool = True):
        """tokens / labels / type_ids / per-token (sample ordinal, offset in sample) of stream range [start, end).  The last token
        of a fragment whose sample goes on in the next pack is labelled with its true successor (``PackedDatasetWithCut``,
        reference ``packed_dataset.py:318``) or, ``label_across_cut=False``, ignored like a sample end
        (``PackedDatasetWithoutCuSeqlen``, reference ``:190-196``)."""
        first = int(np.searchsorted(self.acm_len_samples, start, side="right"))
        last = int(np.searchsorted(self.acm_len_samples, end, side="left"))
        last = min(last, len(self.sample_indices) - 1)
        toks, labs, tids, offs, ords = [], [], [], [], []
        for pos in range(first, last + 1):
            s_begin = int(self.acm_len_samples[pos] - self.len_samples_shuffled[pos])
            lo = max(start, s_begin) - s_begin
            hi = min(end, int(self.acm_len_samples[pos])) - s_begin
            if hi <= lo:
                continue
            sample = self.dataset[int(self.sample_indices[pos])]
            t = np.asarray(sample["tokens"], dtype=np.int64)
            chunk = t[lo:hi]
            nxt = np.empty_like(chunk)
            nxt[:-1] = chunk[1:]
            nxt[-1] = t[hi] if hi < len(t) and label_across_cut else -100
            toks.append(chunk)
            labs.append(nxt)
            tids.append(np.full(len(chunk), sample.get("type_id", 0), dtype=np.int64))
            offs.append(np.arange(0, hi - lo))  # position restarts at every pack/sample fragment (reference semantics)
            ords.append(len(chunk))
        return toks, labs, tids, offs, ords

    def __getitem__(self, item: int) -> Dict:
        return self.build_pack(item)

    # ---- the stream arithmetic under the reference's names (``packed_dataset.py:122-131,243-282,333-340``) ---------------
    def accu_sample_len(self, seed=None):
        """``(shuffled sample ids, their lengths, running total of those lengths)`` for ``seed`` (default: dataset seed - 1)."""
        idx, lens, acc = self._shuffled(self.seed - 1 if seed is None else seed)
        return idx, lens.tolist(), acc.tolist()

    def find_offset(self, offset: int):
        """Token ``offset`` of the stream -> ``(position of its sample in the shuffled order, offset inside that sample)``."""
        pos = int(np.searchsorted(self.acm_len_samples, offset, side="right"))
        return pos, int(offset - (self.acm_len_samples[pos - 1] if pos > 0 else 0))

    def cal_map(self, carriage_idx: int = 0) -> int:
        """Shuffled position of the sample that holds the LAST token of pack ``carriage_idx``."""
        assert carriage_idx >= 0
        return int(np.searchsorted(self.acm_len_samples, (carriage_idx + 1) * self.packed_length, side="left"))

    def mapping(self, pack_idx: int = 0):
        """``(first sample position, token offset in it, last sample position, one-past-last token offset in it)`` of a pack."""
        pre_pos, pre_tok = self.find_offset(pack_idx * self.packed_length) if pack_idx > 0 else (0, 0)
        pos = self.cal_map(pack_idx)
        tok = int(self.len_samples_shuffled[pos] - (self.acm_len_samples[pos] - (pack_idx + 1) * self.packed_length))
        return pre_pos, pre_tok, pos, tok

    def cal_pos_unpack(self, index: int):
        """Sample range ``[pre_pos, pos)`` of un-packed micro-batch ``index``."""
        mb = gpc.config.data["micro_bsz"]
        return index * mb, (index + 1) * mb

    def pdebug(self, line) -> None:
        if getattr(self, "debug", False):
            print(line, flush=True)


class PackedDatasetWithCut(PackedDataset):
    """Concatenate shuffled samples into a token stream and cut it every ``packed_length`` tokens; fragments longer
    than ``max_length_per_sample`` are split into several attention segments (reference ``packed_dataset.py:206-389``).
    """

    def __init__(self, dataset, max_length_per_sample: int = 2048, packed_length: int = 4096):
        super().__init__(dataset, max_length_per_sample, packed_length)
        self.sample_indices, self.len_samples_shuffled, self.acm_len_samples = self._shuffled(self.seed)

    def __len__(self):
        return self.num_tokens // self.packed_length

    def _segments(self, frag_lens: List[int]):
        cu, idx = [0], []
        for n in frag_lens:
            full, left = divmod(n, self.max_length_per_sample)
            for _ in range(full):
                cu.append(cu[-1] + self.max_length_per_sample)
                idx.append(np.arange(self.max_length_per_sample))
            if left:
                cu.append(cu[-1] + left)
                idx.append(np.arange(left))
        return cu, (np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64))

    def build_pack(self, item: int):
        start, end = item * self.packed_length, (item + 1) * self.packed_length
        toks, labs, tids, _, ords = self._gather(start, end)
        cu, indexes = self._segments(ords)
        return {"tokens": np.concatenate(toks).tolist(), "cu_seqlens": cu, "indexes": indexes.tolist(),
                "labels": np.concatenate(labs).tolist(), "type_ids": np.concatenate(tids).tolist()}

    def build_unpack(self, index: int):
        """Non-packed mode: ``micro_bsz`` whole samples (truncated to ``max_length_per_sample``), zero padded."""
        mb = gpc.config.data["micro_bsz"]
        pack, labels, type_ids, indexes, cu = [], [], [], [], [0]
        for pos in range(index * mb, min((index + 1) * mb, len(self.dataset))):
            sample = self.dataset[int(self.sample_indices[pos])]
            chunk = list(sample["tokens"][: self.max_length_per_sample])
            pack.extend(chunk)
            labels.extend(chunk[1:] + [-100])
            type_ids.extend([sample.get("type_id", 0)] * len(chunk))
            cu.append(cu[-1] + len(chunk))
            indexes.extend(range(len(chunk)))
        if cu[-1] != self.packed_length:
            pad = self.packed_length - cu[-1]
            pack += [0] * pad
            labels += [0] * pad
            type_ids += [0] * pad
            indexes.extend(range(pad))
            cu.append(self.packed_length)
        return {"tokens": pack, "cu_seqlens": cu, "indexes": indexes, "labels": labels, "type_ids": type_ids}

    def __getitem__(self, item: int) -> Dict:
        if gpc.config is not None and gpc.config.get("model") is not None and not gpc.config.model.get(
            "use_flash_attn", True
        ):
            return self.build_unpack(item)
        return self.build_pack(item)


class PackedDatasetWithoutCuSeqlen(PackedDataset):
    """``pack_sample_into_one``: documents are glued without attention boundaries — a single segment per
    ``max_length_per_sample`` window, positions run through (reference ``packed_dataset.py:70-203``)."""

    def __init__(self, dataset, max_length_per_sample: int = 2048, packed_length: int = 4096, debug=False):
        super().__init__(dataset, max_length_per_sample, packed_length)
        assert packed_length % max_length_per_sample == 0
        assert len(getattr(dataset, "lengths", self.lengths)) == len(dataset)
        self.bsz = packed_length // max_length_per_sample
        self.packed_length = packed_length
        self.debug = debug
        self.sample_indices, self.len_samples_shuffled, self.acm_len_samples = self._shuffled(self.seed)
        self.cu_seqlens = list(range(0, packed_length + 1, max_length_per_sample))
        self.indexes = list(range(max_length_per_sample)) * self.bsz

    def __len__(self):
        return self.num_tokens // self.packed_length

    def build_pack(self, item: int):
        start, end = item * self.packed_length, (item + 1) * self.packed_length
        toks, labs, tids, _, _ = self._gather(start, end, label_across_cut=False)
        return {"tokens": np.concatenate(toks).tolist(), "cu_seqlens": list(self.cu_seqlens),
                "indexes": list(self.indexes), "labels": np.concatenate(labs).tolist(),
                "type_ids": np.concatenate(tids).tolist()}


DATASET_TYPE_IDS_MAP = {"en": 0, "cn": 1, "code": 2}


def get_dataset_dict(folder, split="valid", min_length: int = 50) -> Dict[str, Dataset]:
    """``{sub-folder name: ConcatDataset of its *.bin files whose name contains `split`}`` — one validation set per data
    source, walked in sorted order so every rank builds the same dict (reference ``data/tokenized/dataset.py:9-56``)."""
    assert os.path.exists(folder), f"folder `{folder}` not exists"
    if os.path.isfile(folder):
        return {os.path.basename(os.path.dirname(os.path.abspath(folder))) or "val": JsonlDataset(folder, min_length=0)}
    out: Dict[str, Dataset] = {}
    for root, dirs, files in os.walk(folder, followlinks=True):
        dirs.sort()
        bins = [os.path.join(root, f) for f in sorted(files) if f.endswith(".bin") and split in f]
        if bins:
            out[os.path.basename(os.path.normpath(root))] = ConcatDataset([JsonlDataset(b, min_length=min_length) for b in bins])
    return out


def get_dataset_type_ids_map(path):
    """``{sub-folder name: type id}`` in sorted order (reference ``data/utils.py:11-14``)."""
    return {name: i for i, name in enumerate(sorted(os.listdir(path)))}


def get_dataset_type_id(dataset_type_ids_map, path):
    import re

    matches = [v for k, v in dataset_type_ids_map.items()
               if re.search(rf"/[z_]*{re.escape(k)}/", path) or path.startswith(f"{k}/")]
    assert len(matches) == 1, f"{path} should match exactly one of {list(dataset_type_ids_map)}"
    return matches[0]


def _walk_like_rank0(folder):
    """``os.walk`` of the training folder as RANK 0 sees it, handed to every rank (reference ``packed_dataset.py:427-432``): the
    datasets are concatenated in that order, and the order is part of the data stream a checkpointed sampler position refers to.
    The reference materialises the walk before it sorts the sub-folders, so its order is the file system's; it is kept (a run
    that moves over from the reference, or back, sees the same stream on the same storage), files inside a folder are sorted."""
    import torch.distributed as dist

    triples = [list(os.walk(folder, followlinks=True))] if gpc.get_global_rank() == 0 else [None]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast_object_list(triples, src=0)
    if triples[0] is None:       # single process that is not "rank 0" (tools, tests)
        triples = [list(os.walk(folder, followlinks=True))]
    return triples[0]


def get_packed_dataset_without_short_length(folder, max_length_per_sample=2048, packed_length=4096, show_progress=False,
                                            min_length=50, min_length_dict=None, pack_sample_into_one=False):
    """Every ``*.bin`` file under ``folder`` becomes one packed dataset (short samples filtered: ``min_length``, or the value of
    the ``min_length_dict`` key that occurs in the file's path), all concatenated; a token's type id is the position of its
    top-level sub-folder in the sorted folder listing (reference ``packed_dataset.py:392-480``, ``data/utils.py:11-24``)."""
    assert os.path.exists(folder), f"{folder} does not exist."
    datasets = []
    type_ids_map = get_dataset_type_ids_map(folder)
    for root, _, files in _walk_like_rank0(folder):
        for fn in sorted(files):
            if not fn.endswith(".bin"):
                continue
            fp = os.path.join(root, fn)
            hits = [k for k in (min_length_dict or {}) if k in fp]
            assert len(hits) < 2, f"The file name `{fp}` matched the following resample keys:{hits}"
            ml = min_length_dict[hits[0]] if hits else min_length
            try:
                type_id = get_dataset_type_id(type_ids_map, fp)
            except AssertionError:      # a .bin directly under `folder`: no sub-folder names it
                type_id = 0
            ds = JsonlDataset(fp, type_id, min_length=ml)
            if len(ds) == 0:
                if gpc.is_rank_for_log():
                    logger.info(f"None of the data in `{fp}` is longer than {ml}")
                continue
            if ds.num_tokens < packed_length:
                if gpc.is_rank_for_log():
                    logger.warning(f"skip {fp}: fewer tokens than one pack")
                continue
            cls = PackedDatasetWithoutCuSeqlen if pack_sample_into_one else PackedDatasetWithCut
            datasets.append(cls(ds, max_length_per_sample, packed_length))
    assert datasets, f"no usable .bin dataset under {folder}"
    return ConcatDataset(datasets)


def unpack_data(input_ids, cu_seqlens, is_type_ids: bool = False, padding_v: int = 0):
    """Un-packed (no flash-attention) mode: a packed row ``[b, packed_length]`` built by ``build_unpack`` holds ``micro_bsz`` whole
    samples followed by padding up to ``packed_length``; give the samples back as rows ``[b, micro_bsz, seq_len]`` padded with
    ``padding_v`` (the trailing padding segment is dropped).  ``b == 1`` is squeezed away unless ``is_type_ids`` (reference
    ``internlm/data/utils.py:27-55``; labels should be unpacked with ``padding_v=-100`` so the padding never enters the loss)."""
    bsz = input_ids.shape[0]
    n_seq, max_len = gpc.config.data["micro_bsz"], gpc.config.data["seq_len"]
    out = torch.full((bsz, n_seq, max_len), padding_v, dtype=input_ids.dtype, device=input_ids.device)
    for i in range(bsz):
        cu = cu_seqlens[i] if torch.is_tensor(cu_seqlens) and cu_seqlens.dim() > 1 else (
            cu_seqlens[i] if not torch.is_tensor(cu_seqlens) and hasattr(cu_seqlens[0], "__len__") else cu_seqlens)
        for j in range(min(n_seq, len(cu) - 1)):
            a0, b0 = int(cu[j]), int(cu[j + 1])
            n = min(b0 - a0, max_len)
            out[i, j, :n] = input_ids[i, a0: a0 + n]
    return out.squeeze(0) if bsz == 1 and not is_type_ids else out
