"""InternLM2 by the reference's class names (``internlm/model/modeling_internlm2.py:484-1055``)."""
from ._compat import SPECS, layer_class, model_class
from .families import build_model_with_cfg_internlm2 as build_model_with_cfg  # noqa: F401
from .modules import MHA, FeedForward  # noqa: F401

PackedFlashLlamaLayer1D = layer_class("PackedFlashLlamaLayer1D", SPECS["internlm2"],
                                      "One InternLM2 block: GQA attention with the grouped wqkv layout + SwiGLU MLP.")
PackedFlashLlama1D = model_class("PackedFlashLlama1D", SPECS["internlm2"], "A pipeline chunk of InternLM2 blocks.")
__all__ = ["PackedFlashLlamaLayer1D", "PackedFlashLlama1D", "build_model_with_cfg", "MHA", "FeedForward"]
