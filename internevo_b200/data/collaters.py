"""Collate functions (reference ``internlm/data/tokenized/collaters.py``)."""
from __future__ import annotations

import torch


def packed_collate_fn(batch, packed_length):
    """→ ``({input_ids, cu_seqlens, indexes, type_ids}, labels)``; negative token ids mark loss-masked positions
    (|id| is fed to the model, label becomes -100)."""
    xs, ys, cu_seqlens, indexes, ts = [], [], [], [], []
    for b in batch:
        assert len(b["tokens"]) == packed_length, (len(b["tokens"]), packed_length)
        assert len(b["labels"]) == packed_length and len(b["type_ids"]) == packed_length
        tok = torch.as_tensor(b["tokens"], dtype=torch.long)
        lab = torch.as_tensor(b["labels"], dtype=torch.long)
        xs.append(tok.abs())
        ys.append(torch.where(lab > 0, lab, torch.full_like(lab, -100)))
        ts.append(torch.as_tensor(b["type_ids"], dtype=torch.long))
        cu_seqlens.append(torch.as_tensor(b["cu_seqlens"], dtype=torch.int32))
        indexes.append(torch.as_tensor(b["indexes"], dtype=torch.long))
    xs = torch.nn.utils.rnn.pad_sequence(xs, batch_first=True)
    ys = torch.nn.utils.rnn.pad_sequence(ys, batch_first=True, padding_value=-100)
    ts = torch.nn.utils.rnn.pad_sequence(ts, batch_first=True, padding_value=0)
    indexes = torch.stack(indexes, dim=0)
    if len(set(map(len, cu_seqlens))) == 1:
        cu_seqlens = torch.stack(cu_seqlens, dim=0)
    assert xs.shape[1] == packed_length, (xs.shape[1], packed_length)
    return {"input_ids": xs, "cu_seqlens": cu_seqlens, "indexes": indexes, "type_ids": ts}, ys


def jsonl_ds_collate_fn(batch, max_length_per_sample):
    xs, ys = [], []
    for x in batch:
        t = torch.as_tensor(x["tokens"][:max_length_per_sample], dtype=torch.long)
        lab = torch.where(t > 0, t, torch.full_like(t, -100))
        xs.append(t.abs())
        ys.append(torch.cat([lab[1:], lab.new_full((1,), -100)]))
    xs = torch.nn.utils.rnn.pad_sequence(xs, batch_first=True)
    ys = torch.nn.utils.rnn.pad_sequence(ys, batch_first=True, padding_value=-100)
    xs = torch.cat([xs, xs.new_zeros(len(xs), max_length_per_sample - xs.shape[1])], dim=-1)
    ys = torch.cat([ys, ys.new_full((len(ys), max_length_per_sample - ys.shape[1]), -100)], dim=-1)
    return {"input_ids": xs}, ys
