"""Text generation on top of a trained model: greedy / sampling (top-k, top-p, temperature, repetition penalty), beam
search and token streaming with a per-layer KV cache (reference ``internlm/apis/inference.py:13-966``).

The decode path runs through ``MHA._forward_decode``: prefill uses the training flash-attention kernel, decode steps the
split-KV kernel over the KV cache; greedy / sampling generation replays ONE captured CUDA graph per token (``DecodeGraph``:
the position lives on the device, so nothing in a step depends on the host).  Masked batches and CPU use library SDPA.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F
from torch import nn

__all__ = ["SequenceGenerator", "InferenceParams", "top_k_top_p_filtering", "BeamHypotheses"]

from internevo_b200.utils.logger import get_logger

logger = get_logger(__file__)


class InferenceParams:
    """Generation state handed to the model's attention layers (reference ``:13-46``)."""

    def __init__(self, max_sequence_len, max_batch_size, sequence_len_offset=0, batch_size_offset=0,
                 key_value_memory_dict: dict = None, lengths_per_sample=None, attention_mask=None) -> None:
        self.max_sequence_len: int = max_sequence_len
        self.max_batch_size: int = max_batch_size
        self.sequence_len_offset: int = sequence_len_offset
        self.batch_size_offset: int = batch_size_offset
        self.key_value_memory_dict: dict = key_value_memory_dict if key_value_memory_dict is not None else {}
        self.fused_ft_kernel: bool = False
        self.lengths_per_sample = lengths_per_sample
        self.attention_mask = attention_mask
        self.graph_pos = None   # cuda int32 [1]: current position on the device while a CUDA-graph decode is active

    def reorder_state(self, indices):
        """Beam search: permute the batch dimension of every cached tensor."""
        if self.lengths_per_sample is not None:
            self.lengths_per_sample = self.lengths_per_sample.index_select(index=indices, dim=0)
        for key, value in list(self.key_value_memory_dict.items()):
            if isinstance(value, tuple):
                self.key_value_memory_dict[key] = tuple(v.index_select(0, indices) for v in value)
            else:
                self.key_value_memory_dict[key] = value.index_select(index=indices, dim=0)


def _get_model_device(model):
    assert isinstance(model, nn.Module)
    params = list(model.parameters())
    return params[0].device if params else None


def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, filter_value=-float("Inf"), min_tokens_to_keep=1):
    """Standard top-k / nucleus filtering (reference ``:925-966``)."""
    if top_k > 0:
        top_k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        remove = logits < torch.topk(logits, top_k)[0][..., -1, None]
        logits = logits.masked_fill(remove, filter_value)
    if top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cum = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cum > top_p
        if min_tokens_to_keep > 1:
            remove[..., :min_tokens_to_keep] = 0
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = 0
        remove = remove.scatter(1, sorted_indices, remove)
        logits = logits.masked_fill(remove, filter_value)
    return logits


class BeamHypotheses:
    """n-best list of finished beams (reference ``:883-923``)."""

    def __init__(self, num_beams, max_length, length_penalty, early_stopping):
        self.max_length = max_length - 1
        self.length_penalty, self.early_stopping, self.num_beams = length_penalty, early_stopping, num_beams
        self.hyp: List[Tuple[float, torch.Tensor]] = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.hyp)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / len(hyp) ** self.length_penalty
        if len(self) < self.num_beams or score > self.worst_score:
            self.hyp.append((score, hyp))
            if len(self) > self.num_beams:
                sorted_scores = sorted((s, idx) for idx, (s, _) in enumerate(self.hyp))
                del self.hyp[sorted_scores[0][1]]
                self.worst_score = sorted_scores[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / self.max_length**self.length_penalty


class DecodeGraph:
    """One decode step of the whole model (embedding → L layers → head) captured in a CUDA graph.

    A decode step is ~10 launches per layer of kernels that run for a few microseconds each: launch-bound.  The step is
    made capture-safe by keeping the position on the device (``InferenceParams.graph_pos``: RoPE positions, the KV-cache
    row and the split-KV attention length all read it, see ``MHA._forward_decode_graph``); the graph ends by incrementing
    it, so a replay needs one tiny copy (the new token ids) and nothing else from the host.
    Used by greedy / sampling generation on one tensor-parallel rank without an attention mask; anything else (and any
    capture failure) keeps the eager path."""

    def __init__(self, decoder, params: "InferenceParams", batch: int, device):
        self.params = params
        self.tok = torch.zeros(batch, 1, dtype=torch.long, device=device)
        params.graph_pos = torch.full((1,), params.sequence_len_offset, dtype=torch.int32, device=device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):          # warm-up outside capture: lazy allocations, RoPE tables, workspaces
            for _ in range(2):
                decoder(input_ids=self.tok, inference_params=params)
        torch.cuda.current_stream(device).wait_stream(side)
        params.graph_pos.fill_(params.sequence_len_offset)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits = _logits_of(decoder(input_ids=self.tok, inference_params=params))[:, -1]
            params.graph_pos += 1

    @staticmethod
    def usable(decoder, params: "InferenceParams", device) -> bool:
        import os

        from internevo_b200.core.context import ParallelMode
        from internevo_b200.core.context import global_context as gpc

        if os.environ.get("B200_DECODE_GRAPH", "1") == "0" or device.type != "cuda" or params.attention_mask is not None:
            return False
        if gpc.is_initialized(ParallelMode.TENSOR) and gpc.get_world_size(ParallelMode.TENSOR) > 1:
            return False
        if any(type(m).__name__ == "MoE" for m in decoder.modules()):
            return False   # MoE routing reads slab sizes on the host: not capturable
        p = next(decoder.parameters())
        return p.dtype == torch.bfloat16

    def step(self, tokens: torch.Tensor) -> torch.Tensor:
        self.tok.copy_(tokens)
        self.graph.replay()
        self.params.sequence_len_offset += 1
        return self.logits


def _logits_of(out):
    out = out[0] if isinstance(out, (list, tuple)) else out
    return out.float()


class SequenceGenerator:
    """``SequenceGenerator(decoder, eos_token_id, pad_token_id, bos_token_id).generate(tokens, ...)``."""

    def __init__(self, decoder, eos_token_id, pad_token_id, bos_token_id, additional_eos_token_list=None):
        self.decoder = decoder
        self.eos_token_id, self.pad_token_id, self.bos_token_id = eos_token_id, pad_token_id, bos_token_id
        self.additional_eos_token_list = additional_eos_token_list or []

    # ---------------------------------------------------------------------------------------------------------------
    def _eos_set(self):
        ids = [self.eos_token_id] + list(self.additional_eos_token_list)
        return [i for i in ids if i is not None]

    def _step_logits(self, tokens, inference_params):
        out = self.decoder(input_ids=tokens, inference_params=inference_params)
        return _logits_of(out)[:, -1]

    def _apply_penalty(self, scores, token_ids, repetition_penalty):
        if repetition_penalty != 1.0:
            tok = scores.gather(1, token_ids)
            tok = torch.where(tok < 0, tok * repetition_penalty, tok / repetition_penalty)
            scores = scores.scatter(1, token_ids, tok)
        return scores

    @torch.no_grad()
    def generate(self, tokens: torch.Tensor = None, num_return_sequences: int = 1, max_length: int = 20, num_beams: int = 1,
                 do_sample: bool = True, temperature: float = 1.0, top_k: int = 50, top_p: float = 1.0,
                 repetition_penalty: float = 1, length_penalty: float = 1.0):
        """→ ``[batch, num_return_sequences, length]`` token ids."""
        assert num_return_sequences <= num_beams or num_beams == 1
        if do_sample:
            return sample_generate(self, tokens, max_length, num_beams, temperature, top_k, top_p, repetition_penalty,
                                   length_penalty, num_return_sequences)
        return greedy_generate(self, tokens, max_length, num_beams, num_return_sequences, repetition_penalty,
                               length_penalty)

    @torch.no_grad()
    def streaming_generate(self, tokens: torch.Tensor = None, max_length: int = 20, do_sample: bool = True,
                           temperature: float = 1.0, top_k: int = 50, top_p: float = 1.0, repetition_penalty: float = 1,
                           length_penalty: float = 1.0):
        """Generator yielding the growing ``[batch, 1, length]`` tensor after every new token."""
        if not do_sample:
            temperature, top_k, top_p = 1.0, 1, 1.0
        yield from _no_beam_search(self, tokens, max_length, temperature, top_k, top_p, repetition_penalty, do_sample,
                                   streaming=True)


def greedy_generate(gen, tokens, max_length=20, num_beams=1, num_return_sequences=1, repetition_penalty=1.0,
                    length_penalty=1.0):
    if num_beams == 1:
        out = list(_no_beam_search(gen, tokens, max_length, 1.0, 1, 1.0, repetition_penalty, False))[-1]
    else:
        out = _beam_search(gen, tokens, max_length, num_beams, num_return_sequences, 1.0, 50, 1.0, False,
                           repetition_penalty, length_penalty)
    return out


def sample_generate(gen, tokens, max_length=20, num_beams=1, temperature=1.0, top_k=50, top_p=1.0, repetition_penalty=1.0,
                    length_penalty=1.0, num_return_sequences=1):
    if num_beams == 1:
        return list(_no_beam_search(gen, tokens, max_length, temperature, top_k, top_p, repetition_penalty, True))[-1]
    return _beam_search(gen, tokens, max_length, num_beams, num_return_sequences, temperature, top_k, top_p, True,
                        repetition_penalty, length_penalty)


def _no_beam_search(gen: SequenceGenerator, tokens, max_length, temperature, top_k, top_p, repetition_penalty, do_sample,
                    streaming=False):
    device = _get_model_device(gen.decoder)
    tokens = tokens.to(device)
    B, L0 = tokens.shape
    assert max_length > L0, "max_length must exceed the prompt length"
    params = InferenceParams(max_sequence_len=max_length, max_batch_size=B)
    eos = torch.tensor(gen._eos_set(), device=device)
    done = torch.zeros(B, dtype=torch.bool, device=device)
    seq = tokens
    cur = tokens
    graph = None
    for step in range(L0, max_length):
        if graph is not None:
            scores = graph.step(cur)
        else:
            scores = gen._step_logits(cur, params)
            params.sequence_len_offset += cur.shape[1]
            if step == L0 and max_length - L0 > 2 and DecodeGraph.usable(gen.decoder, params, device):
                try:                           # prefill done (eager); the remaining steps replay one captured graph
                    graph = DecodeGraph(gen.decoder, params, B, device)
                except Exception as e:         # noqa: BLE001 - capture is an optimisation, never a requirement
                    params.graph_pos = None
                    logger.warning(f"CUDA-graph decode unavailable ({type(e).__name__}: {e}); using eager steps")
        scores = gen._apply_penalty(scores, seq, repetition_penalty)
        if do_sample:
            if temperature > 0 and temperature != 1:
                scores = scores / temperature
            probs = F.softmax(top_k_top_p_filtering(scores, top_k, top_p, min_tokens_to_keep=2), dim=-1)
            nxt = torch.multinomial(probs, 1).squeeze(1)
        else:
            nxt = scores.argmax(-1)
        if gen.pad_token_id is not None:
            nxt = nxt.masked_fill(done, gen.pad_token_id)
        seq = torch.cat([seq, nxt[:, None]], 1)
        cur = nxt[:, None]
        if eos.numel():
            done = done | torch.isin(nxt, eos)
        if streaming:
            yield seq[:, None]
        if bool(done.all()):
            break
    if not streaming:
        yield seq[:, None]


def _beam_search(gen: SequenceGenerator, tokens, max_length, num_beams, num_return_sequences, temperature, top_k, top_p,
                 do_sample, repetition_penalty, length_penalty):
    device = _get_model_device(gen.decoder)
    tokens = tokens.to(device)
    B, L0 = tokens.shape
    V = None
    params = InferenceParams(max_sequence_len=max_length, max_batch_size=B * num_beams)
    eos_ids = set(gen._eos_set())
    # first step on the prompt, then expand to beams
    scores = gen._step_logits(tokens, params)
    params.sequence_len_offset += L0
    V = scores.size(-1)
    logp = F.log_softmax(gen._apply_penalty(scores, tokens, repetition_penalty), dim=-1)
    top_lp, top_ix = logp.topk(num_beams, dim=-1)
    params.reorder_state(torch.arange(B, device=device).repeat_interleave(num_beams))
    seq = torch.cat([tokens.repeat_interleave(num_beams, 0), top_ix.reshape(-1, 1)], 1)
    beam_scores = top_lp.reshape(-1)
    hyps = [BeamHypotheses(num_beams, max_length, length_penalty, early_stopping=False) for _ in range(B)]
    finished = [False] * B
    for cur_len in range(L0 + 1, max_length):
        scores = gen._step_logits(seq[:, -1:], params)
        params.sequence_len_offset += 1
        scores = gen._apply_penalty(scores, seq, repetition_penalty)
        if do_sample:
            if temperature > 0 and temperature != 1:
                scores = scores / temperature
            filt = top_k_top_p_filtering(scores, top_k, top_p, min_tokens_to_keep=num_beams + 1)
            lp = F.log_softmax(filt, dim=-1)
            cand = torch.multinomial(F.softmax(filt, dim=-1), 2 * num_beams)
            cand_lp = lp.gather(1, cand) + beam_scores[:, None]
            cand_lp = cand_lp.view(B, -1)
            cand_tok = cand.view(B, -1)
            cand_beam = torch.arange(num_beams, device=device).repeat_interleave(2 * num_beams)[None].expand(B, -1)
        else:
            lp = F.log_softmax(scores, dim=-1) + beam_scores[:, None]
            lp = lp.view(B, num_beams * V)
            cand_lp, idx = lp.topk(2 * num_beams, dim=-1)
            cand_beam, cand_tok = idx // V, idx % V
        order = cand_lp.argsort(dim=-1, descending=True)
        cand_lp, cand_tok, cand_beam = cand_lp.gather(1, order), cand_tok.gather(1, order), cand_beam.gather(1, order)
        new_seq, new_scores, new_src = [], [], []
        for b in range(B):
            kept = 0
            for j in range(cand_lp.size(1)):
                tok, src = int(cand_tok[b, j]), b * num_beams + int(cand_beam[b, j])
                if tok in eos_ids:
                    if j < num_beams:
                        hyps[b].add(seq[src].clone(), float(cand_lp[b, j]))
                    continue
                new_seq.append(torch.cat([seq[src], seq.new_tensor([tok])]))
                new_scores.append(cand_lp[b, j])
                new_src.append(src)
                kept += 1
                if kept == num_beams:
                    break
            while kept < num_beams:  # pad with the last candidate (can only happen when everything hit eos)
                new_seq.append(torch.cat([seq[b * num_beams], seq.new_tensor([gen.pad_token_id or 0])]))
                new_scores.append(cand_lp.new_tensor(-1e9))
                new_src.append(b * num_beams)
                kept += 1
            finished[b] = finished[b] or hyps[b].is_done(float(cand_lp[b].max()))
        seq = torch.stack(new_seq)
        beam_scores = torch.stack(new_scores)
        params.reorder_state(torch.tensor(new_src, device=device))
        if all(finished):
            break
    for b in range(B):
        for j in range(num_beams):
            hyps[b].add(seq[b * num_beams + j], float(beam_scores[b * num_beams + j]))
    out = []
    for b in range(B):
        best = sorted(hyps[b].hyp, key=lambda x: x[0], reverse=True)[:num_return_sequences]
        out.append([h for _, h in best])
    maxlen = max(len(h) for hs in out for h in hs)
    res = tokens.new_full((B, num_return_sequences, maxlen), gen.pad_token_id or 0)
    for b, hs in enumerate(out):
        for j, h in enumerate(hs):
            res[b, j, : len(h)] = h
    return res
