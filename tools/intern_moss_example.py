"""Example: supervised fine-tuning of a converted HF checkpoint on MOSS-002-style dialogue data with the plain
``transformers`` model of ``huggingface/`` (reference ``tools/intern_moss_example.py``).  LoRA is used when ``peft`` is
installed, otherwise all parameters are tuned.  For large-scale SFT use ``train.py`` with ``configs/7B_sft.py``.

    python tools/intern_moss_example.py --model hf_folder --data moss_002_sft.jsonl --epochs 1
"""
import argparse
import os
import sys

import torch
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from moss_002_sft import collate_fn, get_dataset  # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--model", required=True)
    p.add_argument("--data", required=True)
    p.add_argument("--epochs", type=int, default=5)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--lr", type=float, default=9e-6)
    p.add_argument("--max_len", type=int, default=1024)
    p.add_argument("--val_per_steps", type=int, default=1000)
    p.add_argument("--out", default="output")
    a = p.parse_args(argv)
    from transformers import AutoModelForCausalLM, AutoTokenizer, get_linear_schedule_with_warmup

    device = "cuda" if torch.cuda.is_available() else "cpu"
    tok = AutoTokenizer.from_pretrained(a.model, trust_remote_code=True)
    model = AutoModelForCausalLM.from_pretrained(a.model, trust_remote_code=True)
    try:
        from peft import LoraConfig, TaskType, get_peft_model

        model = get_peft_model(model, LoraConfig(task_type=TaskType.CAUSAL_LM, r=32, lora_alpha=32, lora_dropout=0.1,
                                                 target_modules=["wqkv", "wo", "w1", "w2", "w3", "q_proj", "k_proj",
                                                                 "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"]))
    except ImportError:
        print("peft not installed: full-parameter fine-tuning")
    model.to(device)
    train, val = get_dataset(tok, a.data, a.max_len)
    loader = DataLoader(train, batch_size=a.batch_size, shuffle=True, collate_fn=lambda b: collate_fn(b, tok))
    opt = torch.optim.AdamW([q for q in model.parameters() if q.requires_grad], a.lr)
    sched = get_linear_schedule_with_warmup(opt, min(1000, len(loader)), a.epochs * len(loader))
    with open(a.out, "w") as fp:
        for epoch in range(a.epochs):
            model.train()
            for step, batch in enumerate(loader):
                batch = {k: v.to(device) for k, v in batch.items()}
                with torch.autocast(device_type=device, dtype=torch.bfloat16, enabled=device == "cuda"):
                    loss = model(**batch).loss
                loss.backward()
                opt.step()
                sched.step()
                opt.zero_grad()
                if (step + 1) % a.val_per_steps == 0 or step + 1 == len(loader):
                    fp.write(f"Epoch {epoch} Batch {step}: Loss={loss.item()}\n")
                    model.eval()
                    for i in range(len(val)):
                        data, _ = val[i]
                        prefix = tok.decode(data.tolist(), skip_special_tokens=True)
                        gen = model.generate(input_ids=data[None].to(device), do_sample=True, temperature=0.7, top_k=50,
                                             top_p=0.9, repetition_penalty=1.02, max_new_tokens=100)
                        text = tok.decode(gen[0].tolist(), skip_special_tokens=True).replace(prefix, "")
                        fp.write(f"Prefix: {prefix}\nGenerated: {text}\n---------------------------------\n")
                    model.train()
    print("done; log in", a.out)


if __name__ == "__main__":
    main()
