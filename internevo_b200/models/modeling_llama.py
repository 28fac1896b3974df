"""LLaMA-2 by the reference's class names (``internlm/model/modeling_llama.py:476-1023``)."""
from ._compat import SPECS, layer_class, model_class
from .families import build_model_with_cfg_llama as build_model_with_cfg  # noqa: F401
from .modules import MHA as _MHA

PackedFlashLlamaLayer1D = layer_class("PackedFlashLlamaLayer1D", SPECS["llama"],
                                      "One LLaMA-2 block: separate wq / wk / wv projections + SwiGLU MLP.")
PackedFlashLlama1D = model_class("PackedFlashLlama1D", SPECS["llama"], "A pipeline chunk of LLaMA-2 blocks.")


class MHA(_MHA):
    """Attention with separate ``wq`` / ``wk`` / ``wv`` projections (the reference defines it in this file, :46-473)."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("layout", "llama")
        super().__init__(*args, **kwargs)


__all__ = ["PackedFlashLlamaLayer1D", "PackedFlashLlama1D", "MHA", "build_model_with_cfg"]
