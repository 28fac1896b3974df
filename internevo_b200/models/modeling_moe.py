"""InternLM-MoE by the reference's class names (``internlm/model/modeling_moe.py:44-487``): the v1 block whose MLP is a
mixture of experts (pass ``moe_cfg=dict(num_experts=..., moe_type=..., top_k=...)``)."""
from ._compat import SPECS, layer_class, model_class
from .families import build_model_with_moe_cfg  # noqa: F401

PackedFlashBaseLayer1D = layer_class("PackedFlashBaseLayer1D", SPECS["internlm"], "One InternLM-MoE block.")
PackedFlashInternLm1D = model_class("PackedFlashInternLm1D", SPECS["internlm"], "A pipeline chunk of InternLM-MoE blocks.")
__all__ = ["PackedFlashBaseLayer1D", "PackedFlashInternLm1D", "build_model_with_moe_cfg"]
