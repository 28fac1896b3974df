"""Alpaca-style instruction data (``[{"instruction", "input", "output"}, ...]``) → SFT ``.bin/.bin.meta`` with a
train / valid split.  Prompt tokens are stored NEGATED: the collate functions feed ``|id|`` to the model and mask the label,
so only the answer contributes to the loss (same convention as the reference ``tools/alpaca_tokenizer.py:49-81``).

    python tools/alpaca_tokenizer.py alpaca_data.json out_dir tokenizer.model --split_ratio 0.1
"""
import argparse
import json
import os
import random

from tokenizer import load_sp, write_bin_and_meta  # same folder

USER, BOT = "<|User|>:", "<|Bot|>:"


def chat_sample(item):
    user = f"{USER}{item['instruction']}" + (f"\n{item['input']}" if item.get("input") else "")
    return user, item["output"]


def tokenize_sample(item, sp, eoh_id: int, eoa_id: int, nl_id: int, max_len: int = 2048):
    user, answer = chat_sample(item)
    prompt = sp.encode(user) + [eoh_id, nl_id] + sp.encode(BOT)
    ids = [sp.bos_id()] + [-t for t in prompt] + sp.encode(answer) + [eoa_id, nl_id]
    ids = ids[: max_len - 1] + [sp.eos_id()]
    return ids


def main():
    p = argparse.ArgumentParser()
    p.add_argument("dataset_path")
    p.add_argument("output_path")
    p.add_argument("tokenizer_path")
    p.add_argument("--split_ratio", type=float, default=0.1)
    p.add_argument("--eoh_id", type=int, default=103167)
    p.add_argument("--eoa_id", type=int, default=103166)
    p.add_argument("--nl_id", type=int, default=13)
    p.add_argument("--max_len", type=int, default=2048)
    p.add_argument("--seed", type=int, default=1024)
    a = p.parse_args()
    sp = load_sp(a.tokenizer_path)
    data = json.load(open(a.dataset_path, encoding="utf-8"))
    samples = [tokenize_sample(d, sp, a.eoh_id, a.eoa_id, a.nl_id, a.max_len) for d in data]
    random.Random(a.seed).shuffle(samples)
    n_valid = int(len(samples) * a.split_ratio)
    n_tr = write_bin_and_meta(samples[n_valid:], os.path.join(a.output_path, "train", "en", "dataset.bin"))
    n_va = write_bin_and_meta(samples[:n_valid], os.path.join(a.output_path, "valid", "en", "dataset.bin"))
    print(f"train samples: {n_tr}  valid samples: {n_va}")


if __name__ == "__main__":
    main()
