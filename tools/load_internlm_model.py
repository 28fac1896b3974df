"""Load a training checkpoint for inference and generate from it (reference ``tools/load_internlm_model.py:22-298``).

    from tools.load_internlm_model import initialize_internlm_model, internlm_interactive_generation
    model = initialize_internlm_model("INTERNLM2_PUBLIC", "llm_ckpts/1000")
    for text in internlm_interactive_generation(model, tokenizer, "hello", max_length=64):
        print(text)

Single process: the checkpoint's tensor-/pipeline-parallel shards are merged on the fly.  Under ``torchrun`` with N
processes the model is built tensor-parallel over N ranks and every rank takes its slice of the merged weights.
"""
import inspect
import os
import sys
from typing import Callable, Dict, Optional

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from ckpt_io import load_full_state, load_model_config  # noqa: E402

from internevo_b200.apis.inference import SequenceGenerator  # noqa: E402
from internevo_b200.core.context import ParallelMode  # noqa: E402
from internevo_b200.core.context import global_context as gpc  # noqa: E402
from internevo_b200.initialize.launch import launch  # noqa: E402
from internevo_b200.models.sharding import shard_state_dict  # noqa: E402
from internevo_b200.utils.registry import MODEL_INITIALIZER  # noqa: E402


def merge_pp_within_tp(folder: str, del_model_prefix: bool = False) -> Dict[str, torch.Tensor]:
    """Full (all pipeline stages, all tensor shards merged) state dict of a checkpoint folder."""
    return load_full_state(folder, load_model_config(folder).get("embed_split_hidden", True))


def match_fn_signature(func: Callable, args_dict: Dict) -> None:
    """Drop the entries of ``args_dict`` that ``func`` does not accept (unless it takes **kwargs)."""
    params = inspect.signature(func).parameters
    if any(p.kind == p.VAR_KEYWORD for p in params.values()):
        return
    for k in [k for k in args_dict if k not in params]:
        args_dict.pop(k)


def get_tp_rank() -> int:
    return gpc.get_local_rank(ParallelMode.TENSOR) if gpc.is_initialized(ParallelMode.TENSOR) else 0


def get_tp_world_size() -> int:
    return gpc.get_world_size(ParallelMode.TENSOR) if gpc.is_initialized(ParallelMode.TENSOR) else 1


def initialize_internlm_model(model_type: str, ckpt_dir: Optional[str], model_config: Optional[dict] = None,
                              del_model_prefix: bool = False, param_dtype: torch.dtype = torch.bfloat16,
                              training: bool = False, seed: int = 1024, port: int = 23574):
    cfg = load_model_config(ckpt_dir) if ckpt_dir else {}
    cfg.update(model_config or {})
    cfg["dtype"] = param_dtype
    cfg["parallel_output"] = False
    cfg.pop("device", None)
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if not gpc.is_initialized(ParallelMode.GLOBAL):
        launch(config=dict(model_type=model_type, model=dict(cfg),
                           parallel=dict(zero1=dict(size=1), tensor=dict(size=world, mode="mtp"), pipeline=dict(size=1),
                                         weight=dict(size=1, overlap=False, memory_pool=False)),
                           data=dict(seq_len=cfg.get("max_position_embeddings", 2048), micro_num=1, micro_bsz=1)),
               rank=rank, world_size=world, host=os.environ.get("MASTER_ADDR", "127.0.0.1"),
               port=int(os.environ.get("MASTER_PORT", port)), local_rank=int(os.environ.get("LOCAL_RANK", 0)), seed=seed,
               backend="nccl" if torch.cuda.is_available() else "gloo")
    builder = MODEL_INITIALIZER.get_module(model_type)
    kwargs = dict(cfg)
    match_fn_signature(builder, kwargs)
    model = builder(**kwargs)
    if ckpt_dir:
        full = merge_pp_within_tp(ckpt_dir, del_model_prefix)
        sd = shard_state_dict(full, get_tp_rank(), get_tp_world_size(), cfg.get("embed_split_hidden", True))
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not [m for m in missing if "inv_freq" not in m], f"missing keys: {missing}"
    model = model.to(param_dtype)
    if torch.cuda.is_available():
        model = model.cuda()
    return model.train(training)


def get_model_device(model):
    return next(model.parameters()).device


@torch.inference_mode()
def internlm_interactive_generation(model, tokenizer, prompt: str, additional_eos_token_list=None, max_length: int = 512,
                                    do_sample: bool = True, temperature: float = 1.0, top_k: int = 50, top_p: float = 1.0,
                                    repetition_penalty: float = 1.0, length_penalty: float = 1.0,
                                    max_new_tokens: int = None):
    """Yields the decoded text after every generated token (streaming).  ``max_length`` counts prompt + generated tokens
    (the reference's convention); ``max_new_tokens`` (OpenAI's ``max_tokens``) counts generated tokens only."""
    ids = tokenizer.encode(prompt)
    bos, eos = tokenizer.bos_id(), tokenizer.eos_id()
    tokens = torch.tensor([[bos] + list(ids)], device=get_model_device(model))
    gen = SequenceGenerator(decoder=model, eos_token_id=eos, pad_token_id=bos, bos_token_id=bos,
                            additional_eos_token_list=additional_eos_token_list)
    n_prompt = tokens.shape[1]
    if max_new_tokens is not None:
        max_length = n_prompt + max(1, int(max_new_tokens))
    for out in gen.streaming_generate(tokens=tokens, max_length=max_length, do_sample=do_sample, temperature=temperature,
                                      top_k=top_k, top_p=top_p, repetition_penalty=repetition_penalty,
                                      length_penalty=length_penalty):
        yield tokenizer.decode(out[0, n_prompt:].tolist())


if __name__ == "__main__":
    import argparse

    import sentencepiece as spm

    p = argparse.ArgumentParser()
    p.add_argument("--model_type", default="INTERNLM2_PUBLIC")
    p.add_argument("--ckpt_dir", required=True)
    p.add_argument("--tokenizer", required=True)
    p.add_argument("--prompt", default="hello")
    a = p.parse_args()
    sp = spm.SentencePieceProcessor()
    sp.Load(a.tokenizer)
    m = initialize_internlm_model(a.model_type, a.ckpt_dir)
    last = ""
    for last in internlm_interactive_generation(m, sp, a.prompt, max_length=128):
        pass
    print(last)
