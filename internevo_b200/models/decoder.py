"""Packed-sequence decoder-only transformer shared by the InternLM / InternLM2 / LLaMA-2 / InternLM-MoE families.

Each family is a ``FamilySpec`` (sub-module names = checkpoint key names, attention weight layout, MLP kind).  The
block structure and numerics follow the reference ``PackedFlashBaseLayer1D`` / ``PackedFlashLlamaLayer1D``
(``internlm/model/modeling_internlm.py:35-237``, ``modeling_internlm2.py:481-763``, ``modeling_llama.py``,
``modeling_moe.py:36-258``): pre-norm residual stream, ``norm(residual)`` feeding attention / MLP, optional post-norm,
depth-scaled init, ``embed_grad_scale``, vocab-parallel head.

B200-first: residual-add and RMSNorm are one kernel that carries ``(hidden, residual)`` through the stack; the MLP is
one GEMM with SwiGLU epilogue + one GEMM; activations are ``[tokens, hidden]``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from internevo_b200 import ops
from internevo_b200.core.context import (
    IS_REPLICA_ZERO_PARALLEL,
    IS_TENSOR_DATA_PARALLEL,
    IS_TENSOR_EXPERT_DATA_PARALLEL,
    IS_TENSOR_ZERO_PARALLEL,
    IS_WEIGHT_ZERO_PARALLEL,
    ParallelMode,
)
from internevo_b200.core.context import global_context as gpc
from internevo_b200.initialize.initialize_tensor import normal_, scaled_init_method_normal, scaled_init_method_uniform, uniform_
from internevo_b200.parallel.linear import RewardModelLinear, ScaleColumnParallelLinear
from internevo_b200.solver.activation_checkpoint import activation_checkpoint
from internevo_b200.solver.pipeline_utils import partition_uniform
from internevo_b200.utils.logger import get_logger
from internevo_b200.utils.nvtx import nvtx_range

from .modules import Embedding1D, FeedForward, GeluMLP, MHA, VocabParallelEmbedding, _is_isp

logger = get_logger(__file__)


@dataclass
class FamilySpec:
    name: str
    attn_layout: str        # "internlm" | "internlm2" | "llama"
    attn_name: str          # sub-module name of the attention block
    mlp_name: str
    norm1_name: str
    norm2_name: str
    layers_name: str
    embed_name: str
    final_norm_name: str
    head_name: str
    attn_bias: bool


INTERNLM_SPEC = FamilySpec("INTERNLM", "internlm", "mixer", "mlp", "norm1", "norm2", "blocks", "embedding", "norm",
                           "head", True)
INTERNLM2_SPEC = FamilySpec("INTERNLM2_PUBLIC", "internlm2", "attention", "feed_forward", "attention_norm", "ffn_norm",
                            "layers", "tok_embeddings", "norm", "output", False)
LLAMA_SPEC = FamilySpec("LLAMA2", "llama", "attention", "feed_forward", "attention_norm", "ffn_norm", "layers",
                        "tok_embeddings", "norm", "output", False)


def _tp_mode() -> str:
    if gpc.config is None:
        return "mtp"
    t = gpc.config.get("parallel", {}).get("tensor", None)   # a partial config (tools, unit tests) means plain mtp
    return t.get("mode", "mtp") if isinstance(t, dict) else "mtp"


def _linear_group():
    """Group over which linear weights are sharded: WEIGHT for isp, TENSOR otherwise."""
    return gpc.get_group(ParallelMode.WEIGHT if _is_isp() else ParallelMode.TENSOR)


def _make_norm(norm_type, hidden_size, eps, device, dtype):
    if norm_type == "rmsnorm":
        return ops.RMSNorm(hidden_size, eps=eps, device=device, dtype=dtype)
    # fused (dropout +) residual-add + LayerNorm kernel (csrc/layernorm.cu); plain PyTorch on CPU / fp32
    return ops.LayerNorm(hidden_size, eps=eps, device=device, dtype=dtype)


class DecoderLayer(nn.Module):
    """One transformer block. ``forward(hidden, residual, ...) -> (hidden, residual)`` in pre-norm mode."""

    def __init__(self, spec: FamilySpec, hidden_size, num_attention_heads, num_kv_attention_heads=None, mlp_ratio=4.0,
                 attn_drop_rate=0.0, drop_rate=0.0, max_position_embeddings=2048, dtype=torch.float,
                 layer_norm_epsilon=1e-6, checkpoint=False, layer_idx=0, use_dynamic_ntk_rope=False,
                 residual_in_fp32=False, device=None, apply_post_layer_norm=False, no_bias=True, norm_type="rmsnorm",
                 use_scaled_init=True, use_swiglu=True, attn_wqkv_init_std=0.02, attn_other_init_std=0.02,
                 ffn_uplayer_init_std=0.02, ffn_other_init_std=0.02, init_type="normal", rope_base=10000,
                 tp_mode="mtp", moe_cfg: Optional[dict] = None, adapt_hf=True, dropout_selective_checkpoint=True):
        super().__init__()
        self.spec, self.checkpoint, self.layer_idx = spec, checkpoint, layer_idx
        self._nvtx_name = f"layer{layer_idx}"
        self.prenorm = not apply_post_layer_norm
        self.drop_rate = drop_rate
        self.residual_in_fp32 = residual_in_fp32
        self.use_scaled_init, self.use_swiglu = use_scaled_init, use_swiglu
        self.std = dict(wqkv=attn_wqkv_init_std, attn_other=attn_other_init_std, up=ffn_uplayer_init_std,
                        ffn_other=ffn_other_init_std)
        self.init_type = init_type
        bias = spec.attn_bias if spec.attn_layout == "internlm" else (not no_bias)
        attn = MHA(hidden_size, num_attention_heads, num_kv_attention_heads, process_group=_linear_group(),
                   sequence_process_group=gpc.get_group(ParallelMode.TENSOR), bias=bias, rope_base=rope_base,
                   max_position_embeddings=max_position_embeddings, use_dynamic_ntk_rope=use_dynamic_ntk_rope,
                   layout=spec.attn_layout, tp_mode=tp_mode, interleaved_rope=not adapt_hf, layer_idx=layer_idx,
                   device=device, dtype=dtype, dropout=attn_drop_rate)
        setattr(self, spec.attn_name, attn)
        setattr(self, spec.norm1_name, _make_norm(norm_type, hidden_size, layer_norm_epsilon, device, dtype))
        setattr(self, spec.norm2_name, _make_norm(norm_type, hidden_size, layer_norm_epsilon, device, dtype))
        self.is_moe = moe_cfg is not None and moe_cfg.get("num_experts", 1) > 1
        if self.is_moe:
            from .moe import MoE

            mlp = MoE(hidden_size=hidden_size, mlp_ratio=mlp_ratio, device=device, dtype=dtype, **moe_cfg)
        elif use_swiglu:
            mlp = FeedForward(hidden_size, int(hidden_size * mlp_ratio), out_features=hidden_size,
                              process_group=_linear_group(), bias=False, device=device, dtype=dtype, tp_mode=tp_mode)
        else:
            mlp = GeluMLP(hidden_size, int(hidden_size * mlp_ratio), out_features=hidden_size,
                          process_group=_linear_group(), bias=not no_bias, device=device, dtype=dtype, tp_mode=tp_mode)
        setattr(self, spec.mlp_name, mlp)
        self.reset_parameters()

    # ------------------------------------------------------------------------------------------------------------
    def _init_fns(self):
        if self.init_type == "normal":
            return normal_, scaled_init_method_normal
        return uniform_, scaled_init_method_uniform

    def reset_parameters(self):
        """Depth-scaled init: output projections use ``sigma / sqrt(2 * (layer_idx + 1))`` (reference
        ``modeling_internlm2.py:645-675``)."""
        init, scaled = self._init_fns()
        attn = getattr(self, self.spec.attn_name)
        mlp = getattr(self, self.spec.mlp_name)
        with torch.no_grad():
            for name, p in attn.named_parameters():
                if p.ndim == 1:
                    p.zero_()
                elif any(k in name for k in ("wq", "wk", "wv", "Wqkv")):
                    init(std=self.std["wqkv"])(p)
                elif self.use_scaled_init:
                    scaled(sigma=self.std["attn_other"], num_layers=self.layer_idx + 1)(p)
                else:
                    init(std=self.std["attn_other"])(p)
            if self.is_moe:
                return
            for name, p in mlp.named_parameters():
                if p.ndim == 1:
                    p.zero_()
                elif ("w2" in name or "fc2" in name) and self.use_scaled_init:
                    scaled(sigma=self.std["ffn_other"], num_layers=self.layer_idx + 1)(p)
                elif any(k in name for k in ("w13", "w1", "w3", "fc1")):
                    init(std=self.std["up"])(p)
                else:
                    init(std=self.std["ffn_other"])(p)

    def forward(self, hidden_states, residual=None, cu_seqlens=None, indexes=None, inference_params=None,
                max_seqlen=None):
        with nvtx_range(self._nvtx_name):
            if self.checkpoint and self.training:
                return activation_checkpoint(self._forward, False, hidden_states, residual, cu_seqlens, indexes,
                                             inference_params, max_seqlen)
            return self._forward(hidden_states, residual, cu_seqlens, indexes, inference_params, max_seqlen)

    def _drop(self, x):
        return F.dropout(x, self.drop_rate, self.training) if self.drop_rate > 0 else x

    def _forward(self, hidden_states, residual, cu_seqlens, indexes, inference_params, max_seqlen):
        s = self.spec
        attn, mlp = getattr(self, s.attn_name), getattr(self, s.mlp_name)
        norm1, norm2 = getattr(self, s.norm1_name), getattr(self, s.norm2_name)
        moe_loss = None
        if self.prenorm:
            # residual_{l} = dropout(hidden) + residual ; hidden = norm1(residual)   -- one fused kernel
            if residual is None:
                residual = self._drop(hidden_states)
                if self.residual_in_fp32 and residual.dtype != torch.float32:
                    hidden_states, residual = norm1(torch.zeros_like(residual), residual.float())
                else:
                    hidden_states = norm1(residual)
            else:
                hidden_states, residual = norm1(self._drop(hidden_states), residual)
            hidden_states = attn(hidden_states, cu_seqlens=cu_seqlens, indexes=indexes, max_seqlen=max_seqlen,
                                 inference_params=inference_params)
            hidden_states, residual = norm2(self._drop(hidden_states), residual)
            if self.is_moe:
                hidden_states, moe_loss, _ = mlp(hidden_states)
            else:
                hidden_states = mlp(hidden_states)
            if self.is_moe:
                return hidden_states + residual, moe_loss
            return hidden_states, residual
        # post-norm (reference modeling_internlm2.py:741-762)
        assert residual is None
        mixer_out = attn(hidden_states, cu_seqlens=cu_seqlens, indexes=indexes, max_seqlen=max_seqlen,
                         inference_params=inference_params)
        hidden_states = norm1(self._drop(mixer_out) + hidden_states)
        mlp_out = mlp(hidden_states)
        if self.is_moe:
            mlp_out, moe_loss, _ = mlp_out
        hidden_states = norm2(self._drop(mlp_out) + hidden_states)
        if self.is_moe:
            return hidden_states, moe_loss
        return hidden_states, None


class PackedDecoder(nn.Module):
    """A pipeline chunk: optional embedding (``first``), ``num_layers`` blocks, optional final norm + head (``last``)."""

    def __init__(self, spec: FamilySpec, num_layers=12, hidden_size=768, num_attention_heads=12,
                 num_kv_attention_heads=None, vocab_size=50304, mlp_ratio=4.0, attn_drop_rate=0.0, drop_rate=0.0,
                 max_position_embeddings=2048, dtype=torch.float, checkpoint=0.0, layer_norm_epsilon=1e-5, first=False,
                 last=False, embed_split_hidden=False, embed_grad_scale=0.1, parallel_output=True, start_layer_idx=0,
                 use_dynamic_ntk_rope=False, device=None, residual_in_fp32=False, norm_type="rmsnorm", is_reward=False,
                 dropout_selective_checkpoint=True, use_scaled_init=True, use_swiglu=True, use_flash_attn=True,
                 apply_post_layer_norm=False, no_bias=True, embedding_init_std=0.02, attn_wqkv_init_std=0.02,
                 attn_other_init_std=0.02, ffn_uplayer_init_std=0.02, ffn_other_init_std=0.02, out_head_init_std=0.02,
                 init_type="normal", rope_base=10000, norm_head=False, adapt_hf=True, moe_cfg=None, **unused):
        super().__init__()
        self.spec = spec
        self.first, self.last = first, last
        self.start_layer_idx = start_layer_idx   # global index of this chunk's first block (per-expert checkpoint names)
        self.embed_grad_scale = embed_grad_scale
        self.parallel_output = parallel_output
        self.tp_mode = _tp_mode()
        self.is_moe = moe_cfg is not None and moe_cfg.get("num_experts", 1) > 1
        self.hidden_size = hidden_size
        self.use_flash_attn = use_flash_attn
        checkpoint_layer_num = int(num_layers * checkpoint)
        sp = gpc.config.parallel.get("sequence_parallel", False) if gpc.config is not None else False
        init = normal_ if init_type == "normal" else uniform_
        if first:
            if embed_split_hidden or _is_isp():
                emb = Embedding1D(vocab_size, hidden_size, dtype=dtype, device=device)
            else:
                emb = VocabParallelEmbedding(vocab_size, hidden_size, process_group=gpc.get_group(ParallelMode.TENSOR),
                                             sequence_parallel=sp, dtype=dtype, device=device)
            with torch.no_grad():
                init(std=embedding_init_std)(emb.weight)
            setattr(self, spec.embed_name, emb)
        layers = nn.ModuleList([
            DecoderLayer(spec, hidden_size, num_attention_heads, num_kv_attention_heads, mlp_ratio, attn_drop_rate,
                         drop_rate, max_position_embeddings, dtype, layer_norm_epsilon, lid < checkpoint_layer_num,
                         lid + start_layer_idx, use_dynamic_ntk_rope, residual_in_fp32, device, apply_post_layer_norm,
                         no_bias, norm_type, use_scaled_init, use_swiglu, attn_wqkv_init_std, attn_other_init_std,
                         ffn_uplayer_init_std, ffn_other_init_std, init_type, rope_base, self.tp_mode, moe_cfg,
                         adapt_hf, dropout_selective_checkpoint)
            for lid in range(num_layers)
        ])
        setattr(self, spec.layers_name, layers)
        if last:
            if not apply_post_layer_norm:
                setattr(self, spec.final_norm_name, _make_norm(norm_type, hidden_size, layer_norm_epsilon, device, dtype))
            if is_reward:
                head = RewardModelLinear(hidden_size, 1, process_group=gpc.get_group(ParallelMode.TENSOR), bias=False,
                                         device=device, dtype=dtype, weight_scale=embed_grad_scale)
            else:
                head = ScaleColumnParallelLinear(hidden_size, vocab_size, process_group=_linear_group(), bias=False,
                                                 device=device, dtype=dtype, weight_scale=embed_grad_scale,
                                                 norm_head=norm_head)
            with torch.no_grad():
                init(std=out_head_init_std)(head.weight)
            setattr(self, spec.head_name, head)
        self.apply_post_layer_norm = apply_post_layer_norm
        self._set_param_attrs()

    # ------------------------------------------------------------------------------------------------------------
    def _set_param_attrs(self):
        """Tag every parameter with its reduction class (reference ``train/pipeline.py:98-154``)."""
        isp = _is_isp()
        for name, p in self.named_parameters():
            if getattr(p, "is_expert", False):
                setattr(p, IS_TENSOR_EXPERT_DATA_PARALLEL, True)
            elif ("norm" in name.split(".")[-2] or ".gate." in name or name.endswith("gate.weight")
                  or name.endswith(".wg.weight") or name.endswith("coefficient.weight") or name.endswith("coefficient.bias")):
                setattr(p, IS_REPLICA_ZERO_PARALLEL, True)
            elif isp and (self.spec.embed_name in name):
                setattr(p, IS_TENSOR_DATA_PARALLEL, True)
            elif isp:
                setattr(p, IS_WEIGHT_ZERO_PARALLEL, True)
            else:
                setattr(p, IS_TENSOR_ZERO_PARALLEL, True)

    @property
    def layer_list(self):
        return getattr(self, self.spec.layers_name)

    def forward(self, hidden_states=None, cu_seqlens=None, input_ids=None, indexes=None, inference_params=None,
                **kwargs):
        s = self.spec
        if inference_params is not None:
            return self._forward_generate(hidden_states, input_ids, inference_params)
        packed = cu_seqlens is not None
        if self.first and input_ids is not None:
            if not packed:  # [B, S] un-packed batch → flatten with uniform segments
                B, S = input_ids.shape
                cu_seqlens = torch.arange(0, (B + 1) * S, S, device=input_ids.device, dtype=torch.int32)
                indexes = torch.arange(S, device=input_ids.device).repeat(B)
                self._unpacked_shape = (B, S)
            ids = input_ids.reshape(-1)
            hidden_states = getattr(self, s.embed_name)(ids)
            if self.embed_grad_scale != 1:
                hidden_states = self.embed_grad_scale * hidden_states + (1 - self.embed_grad_scale) * hidden_states.detach()
        elif hidden_states is not None and hidden_states.dim() == 3:
            B, S, _ = hidden_states.shape
            if not packed:
                cu_seqlens = torch.arange(0, (B + 1) * S, S, device=hidden_states.device, dtype=torch.int32)
                indexes = torch.arange(S, device=hidden_states.device).repeat(B)
            hidden_states = hidden_states.reshape(-1, hidden_states.shape[-1])
        if isinstance(cu_seqlens, (list, tuple)):  # ragged batch field sliced to one micro-batch
            cu_seqlens = cu_seqlens[0]
        if cu_seqlens is not None and cu_seqlens.dim() > 1:
            cu_seqlens = cu_seqlens.reshape(-1) if cu_seqlens.shape[0] == 1 else cu_seqlens[0]
        if indexes is not None:
            indexes = indexes.reshape(-1)
            max_position = None
        # host-side max_seqlen: computed from the CPU copy kept by the scheduler when available; no .item() sync here
        max_seqlen = kwargs.get("max_seqlen", None)
        if max_seqlen is not None:
            max_seqlen = int(max_seqlen.reshape(-1)[0]) if torch.is_tensor(max_seqlen) else int(max_seqlen)
        if max_seqlen is None and cu_seqlens is not None:
            max_seqlen = int((cu_seqlens[1:] - cu_seqlens[:-1]).max().item())
        if _is_isp() and indexes is not None and gpc.get_world_size(ParallelMode.TENSOR) > 1:
            # activations are sequence-sharded: keep the matching slice of the position ids (cu_seqlens stay global)
            n = gpc.get_world_size(ParallelMode.TENSOR)
            indexes = indexes.chunk(n)[gpc.get_local_rank(ParallelMode.TENSOR)].contiguous()
        residual = None
        moe_losses = []
        for layer in self.layer_list:
            out = layer(hidden_states, residual, cu_seqlens=cu_seqlens, indexes=indexes, max_seqlen=max_seqlen)
            if self.is_moe:
                hidden_states, moe_loss = out
                moe_losses.append(moe_loss)
                residual = None
            else:
                hidden_states, residual = out
        if self.last:
            if not self.apply_post_layer_norm:
                norm = getattr(self, s.final_norm_name)
                if residual is not None:
                    hidden_states, _ = norm(hidden_states, residual)
                else:
                    hidden_states = norm(hidden_states)
            head = getattr(self, s.head_name)
            if isinstance(head, RewardModelLinear):
                hidden_states = head(hidden_states)
            else:
                hidden_states = head(hidden_states, gather_dim=0, tp_mode=self.tp_mode)
            if not self.parallel_output and gpc.get_world_size(ParallelMode.TENSOR) > 1 and not _is_isp():
                from internevo_b200.parallel.functional import gather_forward_split_backward

                hidden_states = gather_forward_split_backward(hidden_states, gpc.get_group(ParallelMode.TENSOR), dim=-1)
        elif residual is not None:
            # pipeline boundary: ship one tensor (hidden + residual folded), the next stage restarts the pair
            hidden_states = (hidden_states + residual).to(hidden_states.dtype)   # residual may be fp32 (residual_in_fp32)
        if self.is_moe:
            return hidden_states, moe_losses
        return hidden_states

    @torch.no_grad()
    def _forward_generate(self, hidden_states, input_ids, inference_params):
        s = self.spec
        if gpc.get_world_size(ParallelMode.TENSOR) > 1 and gpc.config.parallel.get("sequence_parallel", False):
            # a decode step is ONE token per sequence: there is nothing to shard along the sequence, and the msp / fsp / isp
            # linears expect sequence-sharded activations.  Generation runs with tensor mode "mtp" (or tp = 1), as in the
            # reference, whose inference path has no sequence-parallel branch either.
            raise NotImplementedError("generation with a KV cache needs parallel.tensor.mode='mtp' (or tensor size 1); "
                                      "load the checkpoint with an mtp layout for inference")
        if self.first and input_ids is not None:
            B, S = input_ids.shape
            hidden_states = getattr(self, s.embed_name)(input_ids.reshape(-1)).reshape(B, S, -1)
        B, S, Hd = hidden_states.shape
        residual = None
        for layer in self.layer_list:
            attn, mlp = getattr(layer, s.attn_name), getattr(layer, s.mlp_name)
            norm1, norm2 = getattr(layer, s.norm1_name), getattr(layer, s.norm2_name)
            x2 = hidden_states.reshape(B * S, Hd)
            if residual is None:
                residual = x2
                h = norm1(residual)
            else:
                h, residual = norm1(x2, residual)
            h = attn(h.reshape(B, S, Hd), inference_params=inference_params).reshape(B * S, Hd)
            h, residual = norm2(h, residual)
            out = mlp(h)
            if layer.is_moe:
                out = out[0] + residual
                residual = None
            hidden_states = out.reshape(B, S, Hd)
        if self.last:
            x2 = hidden_states.reshape(B * S, Hd)
            norm = getattr(self, s.final_norm_name)
            x2 = norm(x2, residual)[0] if residual is not None else norm(x2)
            logits = getattr(self, s.head_name)(x2)
            if gpc.get_world_size(ParallelMode.TENSOR) > 1 and not _is_isp():
                from internevo_b200.parallel.functional import gather_forward_split_backward

                logits = gather_forward_split_backward(logits, gpc.get_group(ParallelMode.TENSOR), dim=-1)
            return logits.reshape(B, S, -1)
        return hidden_states if residual is None else (hidden_states.reshape(B * S, Hd) + residual).reshape(B, S, Hd)


def build_generic_model_1d(spec: FamilySpec, num_layers, num_chunks, device=None, **kwargs):
    """Uniformly partition ``num_layers`` over pipeline stages × chunks and build this rank's chunk(s)
    (reference ``modeling_internlm2.py:1012-1053``)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    pp_size = gpc.get_world_size(ParallelMode.PIPELINE)
    pp_rank = gpc.get_local_rank(ParallelMode.PIPELINE)
    all_parts = partition_uniform(num_layers, pp_size, num_chunks)
    parts = all_parts[pp_rank]
    if gpc.is_rank_for_log():
        logger.info(f"The layer sharding is {all_parts}.")
    models = []
    kwargs.pop("num_chunks", None)
    for start, end in parts:
        kw = dict(kwargs)
        kw["num_layers"] = end - start
        kw["first"] = start == 0
        kw["last"] = end == num_layers and len(all_parts[-1]) != 0
        kw["device"] = device
        kw["start_layer_idx"] = start
        models.append(PackedDecoder(spec, **kw).to(device))
    model = models[0] if len(models) == 1 else nn.ModuleList(models)
    setattr(model, "first_layer", 0)
    setattr(model, "last_layer", num_layers)
    return model
