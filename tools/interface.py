"""Streaming text generation for HF-style causal LMs (the converted checkpoints of ``huggingface/`` or any model whose
``forward(input_ids, past_key_values=..., use_cache=True)`` returns ``.logits`` / ``.past_key_values``).

Counterpart of the reference's ``tools/interface.py`` (``GenerationConfig`` + ``generate_interactive``), used by the web
demo and by ``tools/pal_inference.py``.  The decoding loop is self-contained (temperature, top-p, top-k, repetition
penalty, extra EOS ids, KV cache) instead of re-entering ``transformers.generate`` internals, so it is independent of the
installed ``transformers`` version.

    for text in generate_interactive(model, tokenizer, "hello", GenerationConfig(max_length=128)):
        print(text)                       # the decoded continuation so far, grows token by token
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterator, List, Optional, Sequence

import torch


@dataclass
class GenerationConfig:
    max_length: int = 64               # prompt + generated tokens
    max_new_tokens: Optional[int] = None
    top_p: float = 0.8
    top_k: int = 0
    temperature: float = 0.8
    do_sample: bool = True
    repetition_penalty: float = 1.0


def _filter_logits(logits: torch.Tensor, top_k: int, top_p: float) -> torch.Tensor:
    """Mask everything outside the top-k tokens / the smallest nucleus of mass ``top_p`` (1-D logits)."""
    if top_k and top_k < logits.numel():
        kth = torch.topk(logits, top_k).values[-1]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if 0.0 < top_p < 1.0:
        sorted_logits, idx = torch.sort(logits, descending=True)
        cum = torch.softmax(sorted_logits, -1).cumsum(-1)
        drop = cum - torch.softmax(sorted_logits, -1) >= top_p     # keep the token that crosses the threshold
        logits = logits.masked_fill(torch.zeros_like(drop).scatter(0, idx, drop), float("-inf"))
    return logits


def sample_next(logits: torch.Tensor, generated: Sequence[int], cfg: GenerationConfig,
                generator: Optional[torch.Generator] = None) -> int:
    logits = logits.float().clone()
    if cfg.repetition_penalty != 1.0 and len(generated):
        seen = torch.tensor(sorted(set(generated)), device=logits.device)
        vals = logits[seen]
        logits[seen] = torch.where(vals > 0, vals / cfg.repetition_penalty, vals * cfg.repetition_penalty)
    if not cfg.do_sample or cfg.temperature <= 0:
        return int(logits.argmax())
    logits = _filter_logits(logits / cfg.temperature, cfg.top_k, cfg.top_p)
    return int(torch.multinomial(torch.softmax(logits, -1), 1, generator=generator))


@torch.inference_mode()
def generate_interactive(model, tokenizer, prompt: str, generation_config: Optional[GenerationConfig] = None,
                         additional_eos_token_id: Optional[int] = None, stop_fn: Optional[Callable[[str], bool]] = None,
                         generator: Optional[torch.Generator] = None, **overrides) -> Iterator[str]:
    """Yield the decoded continuation after every new token; stops at EOS / ``additional_eos_token_id`` / length / ``stop_fn``."""
    cfg = GenerationConfig(**{**(generation_config.__dict__ if generation_config else {}), **overrides})
    device = next(model.parameters()).device
    enc = tokenizer([prompt], return_tensors="pt")
    input_ids = enc["input_ids"].to(device)
    n_prompt = input_ids.shape[1]
    budget = cfg.max_new_tokens if cfg.max_new_tokens is not None else max(cfg.max_length - n_prompt, 0)
    eos: List[int] = [i for i in [getattr(tokenizer, "eos_token_id", None), additional_eos_token_id] if i is not None]
    ids = input_ids[0].tolist()
    out = model(input_ids=input_ids, use_cache=True)
    past = out.past_key_values
    for _ in range(budget):
        nxt = sample_next(out.logits[0, -1], ids, cfg, generator)
        if nxt in eos:
            break
        ids.append(nxt)
        text = tokenizer.decode(ids[n_prompt:], skip_special_tokens=True)
        yield text
        if stop_fn is not None and stop_fn(text):
            break
        out = model(input_ids=torch.tensor([[nxt]], device=device), past_key_values=past, use_cache=True)
        past = out.past_key_values


def generate(model, tokenizer, prompt: str, generation_config: Optional[GenerationConfig] = None, **kw) -> str:
    """Non-streaming convenience wrapper: the final text."""
    text = ""
    for text in generate_interactive(model, tokenizer, prompt, generation_config, **kw):
        pass
    return text
