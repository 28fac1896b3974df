"""InternLM (v1) by the reference's class names (``internlm/model/modeling_internlm.py:38-463``)."""
from ._compat import SPECS, layer_class, model_class
from .families import build_model_with_cfg  # noqa: F401

PackedFlashBaseLayer1D = layer_class("PackedFlashBaseLayer1D", SPECS["internlm"],
                                     "One InternLM v1 block: MHA with a fused, biased Wqkv + SwiGLU MLP (pre-norm).")
PackedFlashInternLm1D = model_class("PackedFlashInternLm1D", SPECS["internlm"],
                                    "A pipeline chunk of InternLM v1 blocks (+ embedding / final norm / head).")
__all__ = ["PackedFlashBaseLayer1D", "PackedFlashInternLm1D", "build_model_with_cfg"]
