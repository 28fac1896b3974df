"""Raw text / JSON-lines → the training format: a ``.bin`` file of JSON lines ``{"tokens": [...]}`` plus a
``.bin.meta`` numpy array of ``(byte_offset, n_tokens)`` rows (what ``internevo_b200.data.JsonlDataset`` memory-maps).

    python tools/tokenizer.py --text_input_path corpus.txt --bin_output_path data/train/en/corpus.bin \
        --tokenizer_model tokenizer.model

Input may be ``.txt`` (one sample per line), ``.json`` (a list) or ``.jsonl`` (one object per line, text under ``--key``).
Same role as the reference ``tools/tokenizer.py:16-142`` (which hard-wires its own tokenizer model); the writer is
streaming, so corpora larger than memory are fine.
"""
import argparse
import json
import os
from typing import Iterable, Iterator, List

import numpy as np


def load_sp(model_path: str):
    import sentencepiece as spm

    sp = spm.SentencePieceProcessor()
    sp.Load(model_path)
    return sp


def iter_samples(path: str, key: str = "text") -> Iterator[str]:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".json":
        with open(path, "r", encoding="utf-8") as f:
            for item in json.load(f):
                yield item if isinstance(item, str) else item[key]
    else:
        with open(path, "r", encoding="utf-8") as f:
            for line in f:
                line = line.rstrip("\n")
                if not line.strip():
                    continue
                if ext == ".jsonl":
                    obj = json.loads(line)
                    yield obj if isinstance(obj, str) else obj[key]
                else:
                    yield line


def write_bin_and_meta(token_lists: Iterable[List[int]], bin_path: str) -> int:
    """Stream ``token_lists`` into ``bin_path`` and write ``bin_path + '.meta'``; returns the number of samples."""
    os.makedirs(os.path.dirname(os.path.abspath(bin_path)), exist_ok=True)
    meta = []
    offset = 0
    with open(bin_path, "wb") as out:
        for toks in token_lists:
            if not toks:
                continue
            raw = (json.dumps({"tokens": list(map(int, toks))}) + "\n").encode()
            out.write(raw)
            meta.append((offset, len(toks)))
            offset += len(raw)
    np.save(open(bin_path + ".meta", "wb"), np.asarray(meta, dtype=np.int64).reshape(-1, 2))
    return len(meta)


def text2bin(text_input_path: str, bin_output_path: str, tokenizer_model: str, key: str = "text", add_bos: bool = True,
             add_eos: bool = True) -> int:
    sp = load_sp(tokenizer_model)

    def gen():
        for s in iter_samples(text_input_path, key):
            ids = sp.encode(s)
            if add_bos and sp.bos_id() >= 0:
                ids = [sp.bos_id()] + ids
            if add_eos and sp.eos_id() >= 0:
                ids = ids + [sp.eos_id()]
            yield ids

    return write_bin_and_meta(gen(), bin_output_path)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--text_input_path", required=True, help="*.txt | *.json | *.jsonl")
    p.add_argument("--bin_output_path", required=True, help="output *.bin (a *.bin.meta is written next to it)")
    p.add_argument("--tokenizer_model", default=os.environ.get("TOKENIZER_MODEL", "tokenizer.model"))
    p.add_argument("--key", default="text")
    p.add_argument("--no_bos", action="store_true")
    p.add_argument("--no_eos", action="store_true")
    a = p.parse_args()
    assert a.bin_output_path.endswith(".bin"), "the dataset walker only picks up *.bin files"
    n = text2bin(a.text_input_path, a.bin_output_path, a.tokenizer_model, a.key, not a.no_bos, not a.no_eos)
    print(f"wrote {n} samples to {a.bin_output_path} (+ .meta)")


if __name__ == "__main__":
    main()
