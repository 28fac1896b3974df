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
    p.add_argument("--shuffle_seed", type=int, default=None,
                   help="shuffle the samples with this seed and cut the first split_ratio off as validation set; default: the "
                        "reference's split (below)")
    a = p.parse_args()
    sp = load_sp(a.tokenizer_path)
    data = json.load(open(a.dataset_path, encoding="utf-8"))
    samples = [tokenize_sample(d, sp, a.eoh_id, a.eoa_id, a.nl_id, a.max_len) for d in data]
    if a.shuffle_seed is not None:
        random.Random(a.shuffle_seed).shuffle(samples)
        n_valid = int(len(samples) * a.split_ratio)
        train, valid = samples[n_valid:], samples[:n_valid]
    else:
        # the reference's split (``tools/alpaca_tokenizer.py:123-139``): file order is kept, the validation samples are the indices
        # numpy draws WITH replacement under seed 0 - same files byte for byte, duplicates in the draw make the set a bit smaller
        import numpy as np

        np.random.seed(0)
        picked = set(np.random.choice(range(len(samples)), int(len(samples) * a.split_ratio)).tolist())
        train = [s for i, s in enumerate(samples) if i not in picked]
        valid = [s for i, s in enumerate(samples) if i in picked]
    n_tr = write_bin_and_meta(train, os.path.join(a.output_path, "train", "en", "dataset.bin"))
    n_va = write_bin_and_meta(valid, os.path.join(a.output_path, "valid", "en", "dataset.bin"))
    print(f"train samples: {n_tr}  valid samples: {n_va}")


if __name__ == "__main__":
    main()
