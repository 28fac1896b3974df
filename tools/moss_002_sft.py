"""Dialogue SFT data for the HF fine-tuning example (``tools/intern_moss_example.py``): MOSS-002 style samples
``{"plain_text": "...<eoh> ...<eoa>...", "num_turns": n, "prefix": meta_instruction}`` are tokenised turn by turn, cut at
``max_len`` on a turn boundary, and the meta instruction is excluded from the loss (reference
``tools/internlm_sft_on_moss.py``).  Data comes from a local JSON-lines file (no hub access here)."""
from __future__ import annotations

import copy
import json
from typing import Dict, List, Tuple

import torch
from torch.utils.data import Dataset


def process(sample: Dict, tokenizer, max_len: int) -> Dict:
    """→ ``{"input_ids": [...], "no_loss_spans": [(start, end), ...]}``; empty when not even one turn fits."""
    chat = sample["plain_text"].split("<eoa>")[:-1]
    instruction_ids = tokenizer.encode(sample["prefix"])
    assert isinstance(instruction_ids, list) and len(instruction_ids) > 0
    input_ids = copy.deepcopy(instruction_ids)
    spans: List[Tuple[int, int]] = [(0, len(instruction_ids))]     # no loss on the instruction
    for i in range(min(sample["num_turns"], len(chat))):
        turn = tokenizer.encode(chat[i] + "<eoa>", add_special_tokens=False)
        if len(input_ids) + len(turn) > max_len:
            break
        input_ids.extend(turn)
    if len(input_ids) == len(instruction_ids):
        return {"input_ids": [], "no_loss_spans": []}
    return {"input_ids": input_ids, "no_loss_spans": spans}


class SFTDataset(Dataset):
    def __init__(self, samples: List[Dict]):
        self.samples = samples

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, index):
        s = self.samples[index]
        data = torch.tensor(s["input_ids"], dtype=torch.long)
        label = data.clone()
        for a, b in s["no_loss_spans"]:
            label[a:b] = -100
        return data, label


def collate_fn(batch, tokenizer):
    ids, labels = zip(*batch)
    pad = tokenizer.eos_token_id
    input_ids = torch.nn.utils.rnn.pad_sequence(list(ids), batch_first=True, padding_value=pad)
    labels = torch.nn.utils.rnn.pad_sequence(list(labels), batch_first=True, padding_value=-100)
    lengths = torch.tensor([len(x) for x in ids])
    attention_mask = (torch.arange(input_ids.shape[1])[None] < lengths[:, None]).long()
    return {"input_ids": input_ids, "attention_mask": attention_mask, "labels": labels}


def get_dataset(tokenizer, path: str, max_len: int = 1024, num: int = -1, test_size: int = 10):
    """Local JSONL → (train, validation) ``SFTDataset`` pair; the last ``test_size`` usable samples validate."""
    samples = []
    for line in open(path):
        if line.strip():
            p = process(json.loads(line), tokenizer, max_len)
            if p["input_ids"]:
                samples.append(p)
        if 0 < num <= len(samples):
            break
    cut = max(len(samples) - test_size, 1)
    return SFTDataset(samples[:cut]), SFTDataset(samples[cut:])
