"""HF ``transformers`` model for InternLM (v1) checkpoints produced by ``tools/convert2hf.py --family internlm``:
LLaMA-shaped blocks whose attention projections carry a bias (reference ``transformers/internlm_model/modeling_internlm.py``).
Shares the numerics helpers (RMSNorm, rotary, masks) with ``huggingface/internlm2_model``."""
import torch
import torch.nn.functional as F
from torch import nn
from transformers.cache_utils import DynamicCache
from transformers.generation import GenerationMixin
from transformers.modeling_outputs import BaseModelOutputWithPast, CausalLMOutputWithPast
from transformers.modeling_utils import PreTrainedModel

from ..internlm2_model.modeling_internlm2 import (InternLM2RMSNorm, InternLM2RotaryEmbedding, _causal_mask,
                                                  apply_rotary_pos_emb)
from .configuration_internlm import InternLMConfig


class InternLMMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class InternLMAttention(nn.Module):
    def __init__(self, config: InternLMConfig, layer_idx: int):
        super().__init__()
        self.layer_idx = layer_idx
        self.hidden_size, self.num_heads = config.hidden_size, config.num_attention_heads
        self.head_dim = self.hidden_size // self.num_heads
        self.q_proj = nn.Linear(self.hidden_size, self.hidden_size, bias=config.bias)
        self.k_proj = nn.Linear(self.hidden_size, self.hidden_size, bias=config.bias)
        self.v_proj = nn.Linear(self.hidden_size, self.hidden_size, bias=config.bias)
        self.o_proj = nn.Linear(self.hidden_size, self.hidden_size, bias=config.bias)
        scaling = {"type": "dynamic", "factor": config.rotary.get("scaling_factor", 1.0)} \
            if config.rotary.get("type") == "dynamic" else None
        self.rotary_emb = InternLM2RotaryEmbedding(self.head_dim, config.max_position_embeddings,
                                                   config.rotary.get("base", config.rope_theta), scaling)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None):
        B, S, _ = hidden_states.shape
        q = self.q_proj(hidden_states).view(B, S, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.k_proj(hidden_states).view(B, S, self.num_heads, self.head_dim).transpose(1, 2)
        v = self.v_proj(hidden_states).view(B, S, self.num_heads, self.head_dim).transpose(1, 2)
        cos, sin = self.rotary_emb(position_ids, q.dtype)
        q, k = apply_rotary_pos_emb(q, k, cos, sin)
        if past_key_value is not None:
            k, v = past_key_value.update(k, v, self.layer_idx)
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
        return self.o_proj(out.transpose(1, 2).reshape(B, S, self.hidden_size))


class InternLMDecoderLayer(nn.Module):
    def __init__(self, config, layer_idx):
        super().__init__()
        self.self_attn = InternLMAttention(config, layer_idx)
        self.mlp = InternLMMLP(config)
        self.input_layernorm = InternLM2RMSNorm(config.hidden_size, config.rms_norm_eps)
        self.post_attention_layernorm = InternLM2RMSNorm(config.hidden_size, config.rms_norm_eps)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None):
        hidden_states = hidden_states + self.self_attn(self.input_layernorm(hidden_states), attention_mask, position_ids,
                                                       past_key_value)
        return hidden_states + self.mlp(self.post_attention_layernorm(hidden_states))


class InternLMPreTrainedModel(PreTrainedModel):
    config_class = InternLMConfig
    base_model_prefix = "model"
    supports_gradient_checkpointing = True
    _no_split_modules = ["InternLMDecoderLayer"]
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


class InternLMModel(InternLMPreTrainedModel):
    def __init__(self, config: InternLMConfig):
        super().__init__(config)
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, config.pad_token_id)
        self.layers = nn.ModuleList([InternLMDecoderLayer(config, i) for i in range(config.num_hidden_layers)])
        self.norm = InternLM2RMSNorm(config.hidden_size, config.rms_norm_eps)
        self.post_init()

    def get_input_embeddings(self):
        return self.embed_tokens

    def set_input_embeddings(self, value):
        self.embed_tokens = value

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                use_cache=None, output_hidden_states=None, **kwargs):
        use_cache = self.config.use_cache if use_cache is None else use_cache
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        B, S, _ = inputs_embeds.shape
        if use_cache and past_key_values is None:
            past_key_values = DynamicCache()
        past = past_key_values.get_seq_length() if past_key_values is not None else 0
        if position_ids is None:
            position_ids = torch.arange(past, past + S, device=inputs_embeds.device)[None].expand(B, S)
        mask = _causal_mask(attention_mask, S, past, inputs_embeds.dtype, inputs_embeds.device)
        h = inputs_embeds
        all_hidden = () if output_hidden_states else None
        for layer in self.layers:
            if output_hidden_states:
                all_hidden += (h,)
            h = layer(h, mask, position_ids, past_key_values)
        h = self.norm(h)
        if output_hidden_states:
            all_hidden += (h,)
        return BaseModelOutputWithPast(last_hidden_state=h, past_key_values=past_key_values if use_cache else None,
                                       hidden_states=all_hidden)


class InternLMForCausalLM(InternLMPreTrainedModel, GenerationMixin):
    _tied_weights_keys = {}

    def __init__(self, config):
        super().__init__(config)
        self.model = InternLMModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.post_init()

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new):
        self.lm_head = new

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_hidden_states=None, **kwargs):
        out = self.model(input_ids=input_ids, attention_mask=attention_mask, position_ids=position_ids,
                         past_key_values=past_key_values, inputs_embeds=inputs_embeds, use_cache=use_cache,
                         output_hidden_states=output_hidden_states)
        logits = self.lm_head(out.last_hidden_state).float()
        loss = None
        if labels is not None:
            loss = F.cross_entropy(logits[..., :-1, :].reshape(-1, self.vocab_size), labels[..., 1:].reshape(-1).to(logits.device))
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=out.past_key_values,
                                      hidden_states=out.hidden_states)

    @torch.no_grad()
    def chat(self, tokenizer, query: str, history=(), max_new_tokens=1024, do_sample=True, temperature=0.8, top_p=0.8,
             **kwargs):
        prompt = "".join(f"<|User|>:{q}<eoh>\n<|Bot|>:{a}<eoa>\n" for q, a in history) + f"<|User|>:{query}<eoh>\n<|Bot|>:"
        inputs = tokenizer([prompt], return_tensors="pt")
        inputs = {k: v.to(self.device) for k, v in inputs.items() if torch.is_tensor(v)}
        out = self.generate(**inputs, max_new_tokens=max_new_tokens, do_sample=do_sample, temperature=temperature,
                            top_p=top_p, **kwargs)
        text = tokenizer.decode(out[0][inputs["input_ids"].shape[1]:], skip_special_tokens=True).split("<eoa>")[0]
        return text, list(history) + [(query, text)]
