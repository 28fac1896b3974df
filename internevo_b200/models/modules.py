"""Building blocks shared by the model families: embeddings, rotary tables, attention (three weight layouts), SwiGLU
MLP with a fused gate/up weight.

Parity targets: ``internlm/model/modules/{embedding,mlp,multi_head_attention}.py`` and the ``MHA`` classes of
``modeling_internlm2.py:54-478`` / ``modeling_llama.py``.  B200-first differences:

* activations are 2-D ``[tokens, hidden]`` (the packed layout *is* the native layout, no fake batch dim);
* ``w1`` and ``w3`` live in ONE interleaved parameter so gate/up are a single GEMM whose epilogue applies SwiGLU
  (``state_dict`` still exposes ``w1``/``w3`` so checkpoints keep the InternLM2 layout);
* RoPE rotates q and k in place inside the packed ``wqkv`` output in one launch, and attention reads q/k/v as strided
  views of that same buffer (no split / concat copies).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from internevo_b200 import ops
from internevo_b200.core.context import ParallelMode
from internevo_b200.core.context import global_context as gpc
from internevo_b200.ops.attention import flash_attention_packed, flash_attention_varlen
from internevo_b200.ops.rope import (  # noqa: F401  (reference names, embedding.py:89-283)
    ApplyRotaryEmb,
    ApplyRotaryEmbQKV_,
    apply_rotary_emb,
    apply_rotary_emb_qkv_,
)
from internevo_b200.ops.swiglu import swiglu_interleaved_bwd
from internevo_b200.parallel.functional import (
    gather_forward_split_backward,
    reduce_from_group,
    reduce_scatter_seq,
    seq_all_to_all,
    split_forward_gather_backward,
)
from internevo_b200.parallel.linear import get_linear_cls


def _ws(group):
    return 1 if group is None else dist.get_world_size(group)


# ----------------------------------------------------------------------------------------------------------------
# embeddings
# ----------------------------------------------------------------------------------------------------------------
class _ScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s):
        ctx.s = s
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.s, None


def _scale_grad(x, s):
    return _ScaleGrad.apply(x, s)


class Embedding1D(nn.Module):
    """Token embedding split along the *hidden* dimension; forward gathers the hidden shards and (under sequence
    parallel) keeps only the local sequence slice (reference ``modules/embedding.py:17-60``)."""

    def __init__(self, num_embeddings, embedding_dim, *args, padding_idx=None, dtype=None, device=None, **kwargs):
        super().__init__()
        tp = gpc.tensor_parallel_size if gpc.config is not None and not _is_isp() else 1
        self.num_embeddings, self.embed_dim = num_embeddings, embedding_dim
        self.embed_dim_per_partition = embedding_dim // tp
        self.padding_idx = padding_idx
        self.weight = nn.Parameter(torch.empty(num_embeddings, self.embed_dim_per_partition, dtype=dtype, device=device))
        nn.init.normal_(self.weight, std=0.02)

    def forward(self, input_: torch.Tensor) -> torch.Tensor:
        out = F.embedding(input_, self.weight, self.padding_idx)
        group = gpc.get_group(ParallelMode.TENSOR)
        if _is_isp():
            # ISP: weights are replicated over the sequence group, activations sequence-sharded.  Every rank's loss is ITS
            # shard's share of the micro-batch loss (``losses.py``) and the objective is the mean of the shares, so the
            # gathered gradient (one block per shard) carries a factor 1 / tp: without it the embedding would be updated
            # with tp x the true gradient.
            if _ws(group) <= 1:
                return out
            return _scale_grad(split_forward_gather_backward(out, group, dim=0), 1.0 / _ws(group))
        out = gather_forward_split_backward(out, group, dim=-1)
        if gpc.config.parallel.get("sequence_parallel", False) and _ws(group) > 1:
            out = split_forward_gather_backward(out, group, dim=0)
        return out


class VocabParallelEmbedding(nn.Module):
    """Vocabulary-sharded embedding (``embed_split_hidden=False``; replaces flash-attn's ``ParallelGPT2Embeddings``,
    reference ``modeling_internlm2.py:875-889``): masked lookup + all-reduce (reduce-scatter under SP)."""

    def __init__(self, num_embeddings, embedding_dim, process_group=None, sequence_parallel=False, dtype=None,
                 device=None):
        super().__init__()
        ws = _ws(process_group)
        assert num_embeddings % ws == 0
        self.process_group, self.sequence_parallel = process_group, sequence_parallel
        self.vocab_per_rank = num_embeddings // ws
        self.vocab_start = (dist.get_rank(process_group) if ws > 1 else 0) * self.vocab_per_rank
        self.weight = nn.Parameter(torch.empty(self.vocab_per_rank, embedding_dim, dtype=dtype, device=device))
        nn.init.normal_(self.weight, std=0.02)

    def forward(self, input_):
        ws = _ws(self.process_group)
        if ws <= 1:
            return F.embedding(input_, self.weight)
        local = input_ - self.vocab_start
        mask = (local < 0) | (local >= self.vocab_per_rank)
        out = F.embedding(local.masked_fill(mask, 0), self.weight)
        out = out.masked_fill(mask.unsqueeze(-1), 0.0)
        if self.sequence_parallel:
            return reduce_scatter_seq(out, self.process_group)
        return reduce_from_group(out, self.process_group)


class RotaryEmbedding(nn.Module):
    """cos/sin cache indexed by position ids (reference ``modules/embedding.py:263-478``; linear-scaling and
    dynamic-NTK variants are options of the same table object)."""

    def __init__(self, dim: int, base=10000, scale_base=0, device=None, scaling_factor=1.0, max_position_embeddings=0,
                 dynamic_ntk=False):
        super().__init__()
        self.dim = dim
        self.tables = ops.RotaryTables(dim, base, device=device, scaling_factor=scaling_factor,
                                       ntk_max_position=max_position_embeddings if dynamic_ntk else 0)

    def get(self, max_pos: int, device):
        return self.tables.get(max_pos, device)


LinearRotaryEmbedding = RotaryEmbedding


def DynamicNTKScalingRotaryEmbedding(dim, base=10000, scale_base=0, device=None, max_position_embeddings=2048,
                                     scaling_factor=1.0):
    return RotaryEmbedding(dim, base, scale_base, device, scaling_factor, max_position_embeddings, dynamic_ntk=True)


def _is_isp() -> bool:
    try:
        t = gpc.config.parallel["tensor"]
        return isinstance(t, dict) and t.get("mode", "mtp") == "isp"
    except Exception:
        return False


# ----------------------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------------------
def _sp_peer_attention_enabled(max_seqlen, t_local) -> bool:
    """ISP attention over the sequence group: ``B200_SP_ATTN=peer`` forces the in-kernel peer-K/V kernel, ``a2a`` the
    Ulysses all-to-all form (reference behaviour); by default the peer kernel runs when the peer-memory back-ends are on
    (``fused_comm``) and the longest sequence fits one rank's window."""
    mode = os.environ.get("B200_SP_ATTN", "")
    if mode == "a2a" or not torch.cuda.is_available():
        return False
    if mode == "peer":
        return True
    if not (gpc.config is not None and gpc.config.get("fused_comm", False)):
        return False
    return max_seqlen is not None and max_seqlen <= t_local


class MHA(nn.Module):
    """Causal self-attention over packed sequences with rotary embeddings.

    ``layout`` selects the projection weights (and therefore checkpoint key names):

    * ``"internlm2"``: fused ``wqkv`` in the interleaved ``(kv_head, q_per_kv + 2, head_dim)`` order + ``wo``
    * ``"llama"``:     separate ``wq``, ``wk``, ``wv`` + ``wo``
    * ``"internlm"``:  fused ``Wqkv`` in ``(3, head, head_dim)`` order with bias + ``out_proj`` with bias (MHA only)
    """

    def __init__(self, embed_dim: int, num_heads: int, num_kv_heads: Optional[int] = None, process_group=None,
                 sequence_process_group=None, bias: bool = False, rope_base: int = 10000, max_position_embeddings=2048,
                 use_dynamic_ntk_rope=False, rotary_emb_scale_base=0, layout: str = "internlm2", tp_mode="mtp",
                 interleaved_rope: bool = False, layer_idx=None, device=None, dtype=None, causal=True,
                 softmax_scale=None, dropout=0.0):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.num_kv_heads = num_kv_heads or num_heads
        assert embed_dim % num_heads == 0 and num_heads % self.num_kv_heads == 0
        self.head_dim = embed_dim // num_heads
        self.q_per_kv = num_heads // self.num_kv_heads
        self.kv_dim = self.head_dim * self.num_kv_heads
        self.layout, self.tp_mode, self.causal = layout, tp_mode, causal
        self.softmax_scale = softmax_scale
        self.interleaved_rope = interleaved_rope
        self.layer_idx = layer_idx
        self.process_group = process_group
        self.sequence_process_group = sequence_process_group  # ISP: Ulysses group
        self.dropout = dropout
        tp = _ws(process_group) if tp_mode != "isp" else 1
        assert self.num_kv_heads % tp == 0, "kv heads must be divisible by the tensor parallel size"
        self.local_heads, self.local_kv_heads = num_heads // tp, self.num_kv_heads // tp
        col, row = get_linear_cls(tp_mode, "column"), get_linear_cls(tp_mode, "row")
        sp = gpc.config.parallel.get("sequence_parallel", False) if gpc.config is not None else False
        kw = dict(process_group=process_group, sequence_parallel=sp, device=device, dtype=dtype)
        if layout == "internlm2":
            self.wqkv = col(embed_dim, embed_dim + 2 * self.kv_dim, bias=bias, **kw)
            self.wo = row(embed_dim, embed_dim, bias=bias, **kw)
        elif layout == "llama":
            self.wq = col(embed_dim, embed_dim, bias=bias, **kw)
            self.wk = col(embed_dim, self.kv_dim, bias=bias, **kw)
            self.wv = col(embed_dim, self.kv_dim, bias=bias, **kw)
            self.wo = row(embed_dim, embed_dim, bias=bias, **kw)
        elif layout == "internlm":
            assert self.num_kv_heads == num_heads, "InternLM-v1 layout is MHA"
            self.Wqkv = col(embed_dim, 3 * embed_dim, bias=bias, **kw)
            self.out_proj = row(embed_dim, embed_dim, bias=bias, **kw)
        else:
            raise ValueError(layout)
        if use_dynamic_ntk_rope:
            self.rotary_emb = DynamicNTKScalingRotaryEmbedding(self.head_dim, rope_base, rotary_emb_scale_base, device,
                                                               max_position_embeddings, 1.0)
        else:
            self.rotary_emb = RotaryEmbedding(self.head_dim, rope_base, rotary_emb_scale_base, device)

    # -- projections -------------------------------------------------------------------------------------------
    def _qkv(self, x, indexes, max_pos):
        """→ q ``[T, H, D]``, k, v ``[T, Hkv, D]`` (views of one buffer whenever the layout allows), RoPE applied."""
        D = self.head_dim
        cos, sin = self.rotary_emb.get(max_pos, x.device)
        if self.layout == "internlm2":
            qkv = self.wqkv(x)  # [T, Hkv_l * (qpk + 2) * D]
            T = qkv.shape[0]
            gs = self.q_per_kv + 2
            qkv = ops.apply_rotary_packed(qkv, indexes, cos, sin, gs, gs - 1, self.interleaved_rope, head_dim=D)
            g = qkv.view(T, -1, gs, D)
            self._packed = g  # consumed by forward(): the native attention kernel reads q/k/v out of this one buffer
            q = g[:, :, : self.q_per_kv]  # [T, Hkv, qpk, D]  (strided view)
            return q, g[:, :, -2], g[:, :, -1]
        if self.layout == "llama":
            q, k, v = self.wq(x), self.wk(x), self.wv(x)
            T = q.shape[0]
            q = ops.apply_rotary_packed(q, indexes, cos, sin, 1, 1, self.interleaved_rope, head_dim=D).view(T, -1, D)
            k = ops.apply_rotary_packed(k, indexes, cos, sin, 1, 1, self.interleaved_rope, head_dim=D).view(T, -1, D)
            return q, k, v.view(T, -1, D)
        qkv = self.Wqkv(x)  # [T, 3 * H_l * D] ordered (three, head, dim)
        T = qkv.shape[0]
        H = qkv.shape[1] // (3 * D)
        # rotate q and k heads only: group = 3H heads, first 2H rotate
        qkv = ops.apply_rotary_packed(qkv, indexes, cos, sin, 3 * H, 2 * H, self.interleaved_rope, head_dim=D)
        qkv = qkv.view(T, 3, H, D)
        return qkv[:, 0], qkv[:, 1], qkv[:, 2]

    def forward(self, x, cu_seqlens=None, indexes=None, max_seqlen=None, inference_params=None, **kwargs):
        """x ``[T_local, hidden]`` → ``[T_local, hidden]``."""
        if inference_params is not None:
            return self._forward_decode(x, inference_params, indexes)
        max_pos = int(kwargs.get("max_position", 0)) or (int(max_seqlen) if max_seqlen is not None else x.shape[0])
        self._packed = None
        q, k, v = self._qkv(x, indexes, max_pos)
        T = q.shape[0]
        D = self.head_dim
        sp_group = self.sequence_process_group if self.tp_mode == "isp" else None
        packed, self._packed = self._packed, None
        drop_p = float(self.dropout) if self.training else 0.0
        if packed is not None and drop_p == 0.0 and not (sp_group is not None and _ws(sp_group) > 1):
            if cu_seqlens is None:
                cu_seqlens = torch.tensor([0, T], device=x.device, dtype=torch.int32)
                max_seqlen = T
            ctx = flash_attention_packed(packed, cu_seqlens, max_seqlen, causal=self.causal, scale=self.softmax_scale)
            return self.wo(ctx.reshape(T, -1))
        if q.dim() == 4:  # internlm2 grouped view → [T, H, D] (copy only when a library kernel needs it)
            q = q.reshape(T, -1, D)
        # peer kernel when no sequence is longer than one rank's window (packed SFT batches: 2.2x faster than the all-to-all
        # form at sp = 2, profiles/sp_attn_check_n2_r2_v1.json); one long sequence re-reads the earlier ranks' K / V over
        # NVLink once per q block and is better served by the head-scattered form
        if sp_group is not None and _ws(sp_group) > 1 and drop_p == 0.0 and _sp_peer_attention_enabled(
                None if max_seqlen is None else int(max_seqlen), T):
            # sequence-parallel attention with in-kernel peer K / V (parallel/sp_attention.py): no all-to-all at all
            from internevo_b200.parallel.sp_attention import sp_flash_attention

            cu = cu_seqlens
            if cu is None:
                cu = torch.tensor([0, T * _ws(sp_group)], device=x.device, dtype=torch.int32)
                max_seqlen = T * _ws(sp_group)
            ctx = sp_flash_attention(q, k, v, cu, max_seqlen, sp_group, causal=self.causal, scale=self.softmax_scale)
            if ctx is not None:
                proj = self.out_proj if self.layout == "internlm" else self.wo
                return proj(ctx.reshape(T, -1))
        if sp_group is not None and _ws(sp_group) > 1:
            # Ulysses: heads scattered, sequence gathered (reference multi_head_attention.py:56-135)
            q = seq_all_to_all(q.contiguous(), sp_group, scatter_dim=1, gather_dim=0)
            k = seq_all_to_all(k.contiguous(), sp_group, scatter_dim=1, gather_dim=0)
            v = seq_all_to_all(v.contiguous(), sp_group, scatter_dim=1, gather_dim=0)
        if cu_seqlens is None:
            cu_seqlens = torch.tensor([0, q.shape[0]], device=q.device, dtype=torch.int32)
            max_seqlen = q.shape[0]
        ctx = flash_attention_varlen(q, k, v, cu_seqlens, max_seqlen, causal=self.causal, scale=self.softmax_scale,
                                     dropout_p=drop_p)
        if sp_group is not None and _ws(sp_group) > 1:
            ctx = seq_all_to_all(ctx, sp_group, scatter_dim=0, gather_dim=1)
        ctx = ctx.reshape(ctx.shape[0], -1)
        proj = self.out_proj if self.layout == "internlm" else self.wo
        return proj(ctx)

    # -- generation path: KV cache, one sequence per row -------------------------------------------------------------
    def _forward_decode(self, x, inference_params, indexes=None):
        """x ``[B, S, hidden]`` with a per-layer KV cache in ``inference_params`` (reference ``MHA._forward``
        inference branch, ``modeling_internlm2.py:245-402``).  Prefill runs the training attention kernel, decode steps the
        split-KV kernel; masked (left-padded) batches and CPU use library SDPA."""
        B, S, _ = x.shape
        D = self.head_dim
        gpos = getattr(inference_params, "graph_pos", None)
        if gpos is not None and S == 1:
            return self._forward_decode_graph(x, inference_params, gpos)
        off = inference_params.sequence_len_offset
        pos = torch.arange(off, off + S, device=x.device, dtype=torch.int32).repeat(B)
        q, k, v = self._qkv(x.reshape(B * S, -1), pos, off + S)
        q = q.reshape(B, S, -1, D)
        k = k.reshape(B, S, -1, D)
        v = v.reshape(B, S, -1, D)
        cache = inference_params.key_value_memory_dict
        if self.layer_idx not in cache:
            cache[self.layer_idx] = (
                torch.empty(inference_params.max_batch_size, inference_params.max_sequence_len, k.shape[2], D,
                            dtype=k.dtype, device=k.device),
                torch.empty(inference_params.max_batch_size, inference_params.max_sequence_len, k.shape[2], D,
                            dtype=k.dtype, device=k.device),
            )
        kc, vc = cache[self.layer_idx]
        b0 = inference_params.batch_size_offset
        kc[b0:b0 + B, off:off + S] = k
        vc[b0:b0 + B, off:off + S] = v
        proj = self.out_proj if self.layout == "internlm" else self.wo
        if getattr(inference_params, "attention_mask", None) is None and x.is_cuda:
            if S == 1:      # decode step: split-KV kernel over the cache
                from internevo_b200.ops.attention import decode_attention

                o = decode_attention(q[:, 0], kc[b0:b0 + B], vc[b0:b0 + B], off + 1, self.softmax_scale)
                return proj(o.reshape(B, 1, -1))
            if off == 0:    # prefill: the training attention kernel on B packed sequences of length S
                cu = torch.arange(0, (B + 1) * S, S, device=x.device, dtype=torch.int32)
                o = flash_attention_varlen(q.reshape(B * S, -1, D), k.reshape(B * S, -1, D), v.reshape(B * S, -1, D), cu, S,
                                           causal=True, scale=self.softmax_scale)
                return proj(o.reshape(B, S, -1))
        kk, vv = kc[b0:b0 + B, : off + S], vc[b0:b0 + B, : off + S]
        mask = None
        if S > 1:
            mask = torch.ones(S, off + S, dtype=torch.bool, device=x.device).tril(diagonal=off)
        if getattr(inference_params, "attention_mask", None) is not None:
            am = inference_params.attention_mask[:, None, -S:, : off + S] if inference_params.attention_mask.dim() == 3 \
                else inference_params.attention_mask[:, None, None, : off + S]
            mask = am if mask is None else (mask[None, None] & am)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), kk.transpose(1, 2), vv.transpose(1, 2), attn_mask=mask,
                                           scale=self.softmax_scale, enable_gqa=q.shape[2] != kk.shape[2])
        o = o.transpose(1, 2).reshape(B, S, -1)
        return proj(o)


    def _forward_decode_graph(self, x, inference_params, gpos):
        """Decode step whose position lives on the device (``gpos``: cuda int32 ``[1]``): nothing in it depends on a host
        value that changes from step to step, so the whole model step can be captured once in a CUDA graph and replayed
        (``apis/inference.py::DecodeGraph``).  RoPE positions, the KV-cache row and the attention length all read ``gpos``."""
        from internevo_b200.ops.attention import decode_attention

        B = x.shape[0]
        D = self.head_dim
        q, k, v = self._qkv(x.reshape(B, -1), gpos.expand(B).contiguous(), inference_params.max_sequence_len)
        kc, vc = inference_params.key_value_memory_dict[self.layer_idx]
        b0 = inference_params.batch_size_offset
        row = gpos.to(torch.int64)
        kc[b0:b0 + B].index_copy_(1, row, k.reshape(B, 1, -1, D))
        vc[b0:b0 + B].index_copy_(1, row, v.reshape(B, 1, -1, D))
        o = decode_attention(q.reshape(B, -1, D), kc[b0:b0 + B], vc[b0:b0 + B], 1, self.softmax_scale, seqlen_dev=gpos)
        proj = self.out_proj if self.layout == "internlm" else self.wo
        return proj(o.reshape(B, 1, -1))


# ----------------------------------------------------------------------------------------------------------------
# MLP
# ----------------------------------------------------------------------------------------------------------------
class _FusedSwiGLUMLPFn(torch.autograd.Function):
    """h = SwiGLU(x @ W13^T) computed by ONE GEMM with the activation in its epilogue; backward recomputes nothing:
    the stored interleaved gate/up tile feeds the fused dSwiGLU kernel."""

    @staticmethod
    def forward(ctx, x, w13):
        gu, h = ops.matmul_swiglu(x, w13)
        ctx.save_for_backward(x, w13, gu)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, w13, gu = ctx.saved_tensors
        dgu = swiglu_interleaved_bwd(dh.contiguous(), gu)
        dx = ops.matmul(dgu, w13, b_mn=True) if ctx.needs_input_grad[0] else None
        from internevo_b200.ops.gemm import wgrad

        dw = wgrad(dgu, x, w13) if ctx.needs_input_grad[1] else None
        return dx, dw


class FeedForward(nn.Module):
    """SwiGLU MLP ``w2(silu(w1 x) * w3 x)`` (reference ``modules/mlp.py:13-106``).

    ``w1`` and ``w3`` are stored as one parameter ``w13`` of shape ``[2 * F_local, hidden]`` with rows interleaved
    (row ``2j`` = ``w1[j]``, row ``2j+1`` = ``w3[j]``).  ``state_dict`` / ``load_state_dict`` translate to and from the
    separate ``w1.weight`` / ``w3.weight`` keys of the InternLM2 checkpoint layout.
    """

    def __init__(self, in_features, hidden_features, out_features=None, process_group=None, bias=False, device=None,
                 dtype=None, multiple_of: int = 256, tp_mode: str = "mtp", **kwargs):
        super().__init__()
        assert not bias, "SwiGLU MLP has no bias in any supported model"
        out_features = out_features or in_features
        hidden_features = multiple_of * ((hidden_features + multiple_of - 1) // multiple_of)
        self.hidden_features, self.tp_mode, self.process_group = hidden_features, tp_mode, process_group
        ws = _ws(process_group)
        sp = gpc.config.parallel.get("sequence_parallel", False) if gpc.config is not None else False
        row = get_linear_cls(tp_mode, "row")
        if tp_mode == "isp":
            col = get_linear_cls(tp_mode, "column")
            self.w13 = col(in_features, 2 * hidden_features, process_group=process_group, bias=False, device=device,
                           dtype=dtype, multiple_of=2)
        else:
            assert hidden_features % ws == 0
            col = get_linear_cls(tp_mode, "column")
            self.w13 = col(in_features, 2 * hidden_features, process_group=process_group, bias=False,
                           sequence_parallel=sp, device=device, dtype=dtype, multiple_of=2)
        self.w2 = row(hidden_features, out_features, process_group=process_group, bias=False, sequence_parallel=sp,
                      device=device, dtype=dtype)
        self._register_state_dict_hook(self._split_w13)
        self._register_load_state_dict_pre_hook(self._merge_w13)

    @staticmethod
    def _split_w13(module, state_dict, prefix, local_metadata):
        key = prefix + "w13.weight"
        if key in state_dict:
            w = state_dict.pop(key)
            state_dict[prefix + "w1.weight"] = w[0::2].contiguous()
            state_dict[prefix + "w3.weight"] = w[1::2].contiguous()
        return state_dict

    def _merge_w13(self, state_dict, prefix, *args):
        k1, k3 = prefix + "w1.weight", prefix + "w3.weight"
        if k1 in state_dict and k3 in state_dict:
            w1, w3 = state_dict.pop(k1), state_dict.pop(k3)
            state_dict[prefix + "w13.weight"] = torch.stack([w1, w3], dim=1).reshape(-1, w1.shape[-1])

    def forward(self, x):
        w13 = self.w13.weight
        ws = _ws(self.process_group)
        plain = (ws <= 1 or self.tp_mode == "mtp") and self.tp_mode != "isp"
        if plain and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2:
            if ws > 1:
                from internevo_b200.parallel.functional import copy_to_group

                x = copy_to_group(x, self.process_group)
            h = _FusedSwiGLUMLPFn.apply(x, w13)
        else:
            gu = self.w13(x)
            h = ops.swiglu_interleaved(gu)
        return self.w2(h)


class GeluMLP(nn.Module):
    """Non-SwiGLU MLP ``fc2(gelu(fc1 x))`` (replaces flash-attn ``ParallelFusedMLP``, reference
    ``modeling_internlm2.py:613-629``)."""

    def __init__(self, in_features, hidden_features, out_features=None, process_group=None, bias=True, device=None,
                 dtype=None, tp_mode="mtp", **kwargs):
        super().__init__()
        col, row = get_linear_cls(tp_mode, "column"), get_linear_cls(tp_mode, "row")
        sp = gpc.config.parallel.get("sequence_parallel", False) if gpc.config is not None else False
        self.fc1 = col(in_features, hidden_features, process_group=process_group, bias=bias, sequence_parallel=sp,
                       device=device, dtype=dtype)
        self.fc2 = row(hidden_features, out_features or in_features, process_group=process_group, bias=bias,
                       sequence_parallel=sp, device=device, dtype=dtype)

    def forward(self, x):
        fc1 = self.fc1
        if _ws(getattr(fc1, "process_group", None)) <= 1 and x.is_cuda and x.dtype == torch.bfloat16 and \
                type(fc1).__name__.startswith("ColumnParallelLinear"):
            return self.fc2(ops.linear_gelu(x, fc1.weight, fc1.bias))  # GELU in the GEMM epilogue
        return self.fc2(F.gelu(fc1(x), approximate="tanh"))


# ---- reference module names that are one class here ----------------------------------------------------------------
# The reference keeps a FeedForward per tensor-parallel mode (``modules/mlp.py``: FeedForward / MegatronFeedForward /
# ISPFeedForward); here the linear classes returned by ``get_linear_cls(tp_mode)`` carry the mode and the block is shared.
BaseFeedForward = FeedForward
MegatronFeedForward = FeedForward
ISPFeedForward = FeedForward


def get_mlp_cls(tp_mode: str):
    assert tp_mode in ("mtp", "msp", "fsp", "isp"), tp_mode
    return FeedForward


class SelfAttention(nn.Module):
    """Unfused softmax attention on packed-projection input ``qkv [B, S, 3, H, D]`` (reference
    ``multi_head_attention.py`` / flash-attn's module of the same name): fp32 softmax, optional causal mask and
    ``key_padding_mask [B, S]`` (True = keep).  CPU path and numerical oracle for the tcgen05 flash kernel."""

    def __init__(self, causal: bool = False, softmax_scale=None, attention_dropout: float = 0.0):
        super().__init__()
        self.causal, self.softmax_scale = causal, softmax_scale
        self.dropout = nn.Dropout(attention_dropout)

    def forward(self, qkv, causal=None, key_padding_mask=None):
        q, k, v = qkv.unbind(dim=2)
        return _reference_attention(q, k, v, self.causal if causal is None else causal, self.softmax_scale,
                                    key_padding_mask, self.dropout)


class CrossAttention(nn.Module):
    """Unfused attention with separate ``q [B, Sq, H, D]`` and ``kv [B, Sk, 2, Hkv, D]`` (grouped-query when
    ``Hkv < H``); causal masking is aligned to the END of the key sequence, which is what a KV cache needs."""

    def __init__(self, causal: bool = False, softmax_scale=None, attention_dropout: float = 0.0):
        super().__init__()
        self.causal, self.softmax_scale = causal, softmax_scale
        self.dropout = nn.Dropout(attention_dropout)

    def forward(self, q, kv, causal=None, key_padding_mask=None):
        k, v = kv.unbind(dim=2)
        if k.shape[2] != q.shape[2]:
            rep = q.shape[2] // k.shape[2]
            k, v = k.repeat_interleave(rep, dim=2), v.repeat_interleave(rep, dim=2)
        return _reference_attention(q, k, v, self.causal if causal is None else causal, self.softmax_scale,
                                    key_padding_mask, self.dropout)


def _reference_attention(q, k, v, causal, scale, key_padding_mask, dropout):
    Sq, Sk, D = q.shape[1], k.shape[1], q.shape[-1]
    scale = scale if scale is not None else D ** -0.5
    scores = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * scale
    if key_padding_mask is not None:
        scores = scores.masked_fill(~key_padding_mask[:, None, None, :].bool(), float("-inf"))
    if causal:
        row = torch.arange(Sq, device=q.device)[:, None] + (Sk - Sq)
        scores = scores.masked_fill(torch.arange(Sk, device=q.device)[None, :] > row, float("-inf"))
    probs = dropout(torch.softmax(scores, dim=-1).to(v.dtype))
    return torch.einsum("bhqk,bkhd->bqhd", probs, v)


class DistributedAttention(nn.Module):
    """Ulysses wrapper: scatter heads / gather sequence over ``sequence_process_group`` before ``local_attention``,
    inverse afterwards (reference ``multi_head_attention.py:56-135``).  ``MHA`` inlines the same exchange around the
    varlen flash kernel; this module exists for custom attention cores.  Tensors are ``[B, S_local, ..., H, D]``."""

    def __init__(self, local_attention: nn.Module, sequence_process_group, scatter_idx: int = 2, gather_idx: int = 1):
        super().__init__()
        self.local_attn, self.spg = local_attention, sequence_process_group
        self.scatter_idx, self.gather_idx = scatter_idx, gather_idx

    def _a2a(self, x, head_dim, seq_dim, inverse=False):
        if self.spg is None or _ws(self.spg) <= 1:
            return x
        sd, gd = (seq_dim, head_dim) if inverse else (head_dim, seq_dim)
        return seq_all_to_all(x.contiguous(), self.spg, scatter_dim=sd, gather_dim=gd)

    def forward(self, qkv=None, kv=None, q=None, **kwargs):
        if qkv is not None:                                        # [B, S, 3, H, D]
            out = self.local_attn(self._a2a(qkv, 3, 1), **kwargs)
        else:                                                      # q [B, S, H, D], kv [B, S, 2, Hkv, D]
            out = self.local_attn(self._a2a(q, 2, 1), self._a2a(kv, 3, 1), **kwargs)
        return self._a2a(out, 2, 1, inverse=True)                  # [B, S_full, H/sp, D] -> [B, S_local, H, D]

