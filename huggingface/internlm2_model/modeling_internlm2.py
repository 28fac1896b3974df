"""HF ``transformers`` model for InternLM2 checkpoints produced by ``tools/convert2hf.py``.

Parameter names and the grouped ``wqkv`` layout ``[kv_head, (q_per_kv q-heads | k | v), head_dim]`` are those of the
published InternLM2 HF checkpoints (reference ``transformers/internlm2_model/modeling_internlm2.py``): a converted
checkpoint loads with ``InternLM2ForCausalLM.from_pretrained`` and ``tools/revert_hf.py`` maps it back.  This file is a
plain-PyTorch inference / fine-tuning model (SDPA attention, HF ``DynamicCache``); the training path of the framework is
``internevo_b200.models`` with the sm_100a kernels.  ``tests/test_hf_models.py`` checks that both produce the same logits
from the same weights.
"""
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn
from transformers.cache_utils import Cache, DynamicCache
from transformers.generation import GenerationMixin
from transformers.modeling_outputs import BaseModelOutputWithPast, CausalLMOutputWithPast, SequenceClassifierOutputWithPast
from transformers.modeling_utils import PreTrainedModel

from .configuration_internlm2 import InternLM2Config


class InternLM2RMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        dt = x.dtype
        x = x.float()
        x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)
        return self.weight * x.to(dt)


class InternLM2RotaryEmbedding(nn.Module):
    """cos / sin for arbitrary position ids; ``linear`` and ``dynamic`` (NTK) scaling as in the reference checkpoints."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, scaling: Optional[dict] = None):
        super().__init__()
        self.dim, self.max_position_embeddings, self.base = dim, max_position_embeddings, float(base)
        # transformers >= 5 normalises `rope_scaling` to {"rope_type": ..., "rope_theta": ...}; older configs use {"type", "factor"}
        kind = (scaling or {}).get("type", (scaling or {}).get("rope_type"))
        self.kind = kind if kind in ("linear", "dynamic") else None
        self.factor = float((scaling or {}).get("factor", 1.0))

    def _inv_freq(self, seq_len, device):
        base = self.base
        if self.kind == "dynamic" and seq_len > self.max_position_embeddings:
            base = base * ((self.factor * seq_len / self.max_position_embeddings) - (self.factor - 1)) ** (
                self.dim / (self.dim - 2))
        return 1.0 / (base ** (torch.arange(0, self.dim, 2, dtype=torch.float32, device=device) / self.dim))

    @torch.no_grad()
    def forward(self, position_ids: torch.Tensor, dtype):
        pos = position_ids.float()
        if self.kind == "linear":
            pos = pos / self.factor
        inv = self._inv_freq(int(position_ids.max()) + 1, position_ids.device)
        freqs = pos[..., None] * inv                       # [B, S, dim/2]
        emb = torch.cat([freqs, freqs], dim=-1)
        return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(q, k, cos, sin):
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)          # broadcast over heads
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


class InternLM2MLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.w1 = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.w3 = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.w2 = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)

    def forward(self, x):
        return self.w2(F.silu(self.w1(x)) * self.w3(x))


class InternLM2Attention(nn.Module):
    def __init__(self, config: InternLM2Config, layer_idx: int):
        super().__init__()
        self.layer_idx = layer_idx
        self.hidden_size, self.num_heads = config.hidden_size, config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.num_kv_heads = config.num_key_value_heads
        self.groups = self.num_heads // self.num_kv_heads
        if self.head_dim * self.num_heads != self.hidden_size:
            raise ValueError("hidden_size must be divisible by num_attention_heads")
        self.wqkv = nn.Linear(self.hidden_size, (self.num_heads + 2 * self.num_kv_heads) * self.head_dim, bias=config.bias)
        self.wo = nn.Linear(self.num_heads * self.head_dim, self.hidden_size, bias=config.bias)
        self.rotary_emb = InternLM2RotaryEmbedding(self.head_dim, config.max_position_embeddings, config.rope_theta,
                                                   config.rope_scaling)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value: Optional[Cache] = None):
        B, S, _ = hidden_states.shape
        qkv = self.wqkv(hidden_states).view(B, S, self.num_kv_heads, self.groups + 2, self.head_dim)
        q = qkv[..., : self.groups, :].reshape(B, S, self.num_heads, self.head_dim).transpose(1, 2)
        k = qkv[..., -2, :].transpose(1, 2)
        v = qkv[..., -1, :].transpose(1, 2)
        cos, sin = self.rotary_emb(position_ids, q.dtype)
        q, k = apply_rotary_pos_emb(q, k, cos, sin)
        if past_key_value is not None:
            k, v = past_key_value.update(k, v, self.layer_idx)
        k = k.repeat_interleave(self.groups, dim=1)
        v = v.repeat_interleave(self.groups, dim=1)
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
        return self.wo(out.transpose(1, 2).reshape(B, S, self.hidden_size))


class InternLM2DecoderLayer(nn.Module):
    def __init__(self, config, layer_idx):
        super().__init__()
        self.attention = InternLM2Attention(config, layer_idx)
        self.feed_forward = InternLM2MLP(config)
        self.attention_norm = InternLM2RMSNorm(config.hidden_size, config.rms_norm_eps)
        self.ffn_norm = InternLM2RMSNorm(config.hidden_size, config.rms_norm_eps)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None):
        hidden_states = hidden_states + self.attention(self.attention_norm(hidden_states), attention_mask, position_ids,
                                                       past_key_value)
        return hidden_states + self.feed_forward(self.ffn_norm(hidden_states))


class InternLM2PreTrainedModel(PreTrainedModel):
    config_class = InternLM2Config
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["InternLM2DecoderLayer"]
    _skip_keys_device_placement = "past_key_values"
    _supports_sdpa = True

    def _init_weights(self, module):
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)


def _causal_mask(attention_mask, S, past, dtype, device):
    """Additive [B, 1, S, past + S] mask: causal, plus padding from the 2-D ``attention_mask`` (1 = keep)."""
    total = past + S
    rows = torch.arange(past, total, device=device)[:, None]
    cols = torch.arange(total, device=device)[None, :]
    mask = torch.zeros(S, total, dtype=dtype, device=device).masked_fill(cols > rows, torch.finfo(dtype).min)[None, None]
    if attention_mask is not None:
        pad = (attention_mask[:, None, None, :total] == 0)
        mask = mask.expand(attention_mask.shape[0], 1, S, total).masked_fill(pad, torch.finfo(dtype).min)
    return mask


class InternLM2Model(InternLM2PreTrainedModel):
    def __init__(self, config: InternLM2Config):
        super().__init__(config)
        self.padding_idx, self.vocab_size = config.pad_token_id, config.vocab_size
        self.tok_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([InternLM2DecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = InternLM2RMSNorm(config.hidden_size, config.rms_norm_eps)
        self.gradient_checkpointing = False
        self.post_init()

    def get_input_embeddings(self):
        return self.tok_embeddings

    def set_input_embeddings(self, value):
        self.tok_embeddings = value

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                use_cache=None, output_hidden_states=None, return_dict=None, **kwargs):
        use_cache = self.config.use_cache if use_cache is None else use_cache
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("pass exactly one of input_ids / inputs_embeds")
        if inputs_embeds is None:
            inputs_embeds = self.tok_embeddings(input_ids)
        B, S, _ = inputs_embeds.shape
        if use_cache and past_key_values is None:
            past_key_values = DynamicCache()
        past = past_key_values.get_seq_length() if past_key_values is not None else 0
        if position_ids is None:
            position_ids = torch.arange(past, past + S, device=inputs_embeds.device)[None].expand(B, S)
        mask = _causal_mask(attention_mask, S, past, inputs_embeds.dtype, inputs_embeds.device)
        hidden_states = inputs_embeds
        all_hidden = () if output_hidden_states else None
        for layer in self.layers:
            if output_hidden_states:
                all_hidden += (hidden_states,)
            if self.gradient_checkpointing and self.training:
                hidden_states = torch.utils.checkpoint.checkpoint(layer, hidden_states, mask, position_ids, None,
                                                                  use_reentrant=False)
            else:
                hidden_states = layer(hidden_states, mask, position_ids, past_key_values)
        hidden_states = self.norm(hidden_states)
        if output_hidden_states:
            all_hidden += (hidden_states,)
        return BaseModelOutputWithPast(last_hidden_state=hidden_states, past_key_values=past_key_values if use_cache else None,
                                       hidden_states=all_hidden)


class InternLM2ForCausalLM(InternLM2PreTrainedModel, GenerationMixin):
    _tied_weights_keys = {}

    def __init__(self, config):
        super().__init__(config)
        self.model = InternLM2Model(config)
        self.vocab_size = config.vocab_size
        self.output = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.tok_embeddings

    def set_input_embeddings(self, value):
        self.model.tok_embeddings = value

    def get_output_embeddings(self):
        return self.output

    def set_output_embeddings(self, new):
        self.output = new

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_hidden_states=None, return_dict=None, **kwargs):
        out = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                         output_hidden_states=output_hidden_states)
        logits = self.output(out.last_hidden_state).float()
        loss = None
        if labels is not None:
            loss = F.cross_entropy(logits[..., :-1, :].reshape(-1, self.vocab_size), labels[..., 1:].reshape(-1).to(logits.device))
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=out.past_key_values,
                                      hidden_states=out.hidden_states)

    # ---- chat helpers of the published checkpoints (`model.chat(tokenizer, query, history)`)
    @staticmethod
    def build_inputs(tokenizer, query: str, history=(), meta_instruction=""):
        prompt = tokenizer.bos_token or ""
        if meta_instruction:
            prompt += f"<|im_start|>system\n{meta_instruction}<|im_end|>\n"
        for q, a in history:
            prompt += f"<|im_start|>user\n{q}<|im_end|>\n<|im_start|>assistant\n{a}<|im_end|>\n"
        prompt += f"<|im_start|>user\n{query}<|im_end|>\n<|im_start|>assistant\n"
        return tokenizer([prompt], return_tensors="pt")

    @torch.no_grad()
    def chat(self, tokenizer, query: str, history=(), max_new_tokens=1024, do_sample=True, temperature=0.8, top_p=0.8,
             meta_instruction="You are an AI assistant whose name is InternLM.", **kwargs):
        inputs = self.build_inputs(tokenizer, query, history, meta_instruction)
        inputs = {k: v.to(self.device) for k, v in inputs.items() if torch.is_tensor(v)}
        eos = [tokenizer.eos_token_id] + [i for i in [tokenizer.convert_tokens_to_ids("<|im_end|>")] if i is not None]
        out = self.generate(**inputs, max_new_tokens=max_new_tokens, do_sample=do_sample, temperature=temperature,
                            top_p=top_p, eos_token_id=eos, **kwargs)
        text = tokenizer.decode(out[0][inputs["input_ids"].shape[1]:], skip_special_tokens=True).split("<|im_end|>")[0]
        return text, list(history) + [(query, text)]


    def stream_chat(self, tokenizer, query: str, history=(), max_new_tokens=1024, do_sample=True, temperature=0.8, top_p=0.8,
                    meta_instruction="You are an AI assistant whose name is InternLM.", **kwargs):
        """Generator form of :meth:`chat`: yields ``(response_so_far, history + [(query, response_so_far)])`` every time a
        new piece of text is available (the contract of the published checkpoints' ``stream_chat``, reference
        ``transformers/internlm2_model/modeling_internlm2.py:1185``).  ``generate`` runs on a worker thread and hands token ids
        to this generator through a queue; decoding is incremental, so a multi-byte character split over two tokens is held
        back until it is complete."""
        import queue
        import threading

        inputs = self.build_inputs(tokenizer, query, history, meta_instruction)
        inputs = {k: v.to(self.device) for k, v in inputs.items() if torch.is_tensor(v)}
        im_end = tokenizer.convert_tokens_to_ids("<|im_end|>")
        eos = [tokenizer.eos_token_id] + [i for i in [im_end] if i is not None]
        q: "queue.Queue" = queue.Queue()
        done = object()

        class _IdStreamer:   # the `streamer` protocol of generate(): put(token ids) ... end()
            def __init__(self):
                self.prompt_seen = False

            def put(self, value):
                if value.dim() > 1:
                    if value.shape[0] > 1:
                        raise ValueError("stream_chat handles one conversation at a time")
                    value = value[0]
                if not self.prompt_seen:      # the first call carries the prompt
                    self.prompt_seen = True
                    return
                q.put(value.tolist())

            def end(self):
                q.put(done)

        failure = []

        def work():
            try:
                self.generate(**inputs, max_new_tokens=max_new_tokens, do_sample=do_sample, temperature=temperature,
                              top_p=top_p, eos_token_id=eos, streamer=_IdStreamer(), **kwargs)
            except BaseException as e:   # surface the error in the consumer instead of dying silently on the thread
                failure.append(e)
                q.put(done)

        threading.Thread(target=work, daemon=True).start()
        ids, text = [], ""
        hist = list(history)
        yield text, hist + [(query, text)]
        while True:
            item = q.get()
            if item is done:
                break
            ids.extend(t for t in item if t not in eos)
            new = tokenizer.decode(ids, skip_special_tokens=True)
            if new.endswith("\ufffd") or new == text:      # incomplete UTF-8 sequence, or nothing visible yet
                continue
            text = new
            yield text, hist + [(query, text)]
        if failure:
            raise failure[0]


class InternLM2LinearScalingRotaryEmbedding(InternLM2RotaryEmbedding):
    """Position-interpolation RoPE (positions divided by ``scaling_factor``); name of the reference / published code."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, scaling_factor=1.0):
        super().__init__(dim, max_position_embeddings, base, {"type": "linear", "factor": scaling_factor})


class InternLM2DynamicNTKScalingRotaryEmbedding(InternLM2RotaryEmbedding):
    """Dynamic NTK RoPE (the base grows once the sequence exceeds ``max_position_embeddings``)."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000, scaling_factor=1.0):
        super().__init__(dim, max_position_embeddings, base, {"type": "dynamic", "factor": scaling_factor})


class InternLM2ForSequenceClassification(InternLM2PreTrainedModel):
    """Reward-model style head: score of the last non-pad token (reference ``modeling_internlm2.py`` classification head)."""

    def __init__(self, config):
        super().__init__(config)
        self.num_labels = config.num_labels
        self.model = InternLM2Model(config)
        self.score = nn.Linear(config.hidden_size, self.num_labels, bias=False)
        self.post_init()

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, labels=None, **kwargs):
        h = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids, use_cache=False)
        logits = self.score(h.last_hidden_state)
        B = logits.shape[0]
        if attention_mask is not None:
            last = attention_mask.long().cumsum(1).argmax(1)
        elif self.config.pad_token_id is not None and input_ids is not None:
            last = (input_ids != self.config.pad_token_id).long().cumsum(1).argmax(1)
        else:
            last = torch.full((B,), logits.shape[1] - 1, device=logits.device)
        pooled = logits[torch.arange(B, device=logits.device), last]
        loss = None
        if labels is not None:
            loss = F.mse_loss(pooled.squeeze(-1), labels.float()) if self.num_labels == 1 else F.cross_entropy(pooled, labels)
        return SequenceClassifierOutputWithPast(loss=loss, logits=pooled)
