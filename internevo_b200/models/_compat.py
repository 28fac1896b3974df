"""Class names of the reference's per-family model files (``internlm/model/modeling_{internlm,internlm2,llama,moe}.py``) on
top of the one generic decoder: ``PackedFlashBaseLayer1D`` / ``PackedFlashInternLm1D`` etc. are thin subclasses that pin
the ``FamilySpec`` and accept the reference's constructor keywords, so user code that builds a block or a model by class
name keeps working.  Calling convention of a stand-alone block is the reference's: ``block(hidden, cu_seqlens=...,
indexes=..., max_seqlen=...) -> hidden`` (inside ``PackedDecoder`` the blocks thread ``(hidden, residual)`` instead so that
the residual add fuses into the next norm kernel)."""
from __future__ import annotations

import inspect

from .decoder import INTERNLM2_SPEC, INTERNLM_SPEC, LLAMA_SPEC, DecoderLayer, PackedDecoder

_LAYER_KEYS = set(inspect.signature(DecoderLayer.__init__).parameters) - {"self", "spec"}
_MODEL_KEYS = set(inspect.signature(PackedDecoder.__init__).parameters) - {"self", "spec", "unused"}


def _layer_class(name, spec, doc):
    class _Layer(DecoderLayer):
        def __init__(self, *args, **kwargs):
            names = [k for k in inspect.signature(DecoderLayer.__init__).parameters if k not in ("self", "spec")]
            kwargs.update(dict(zip(names, args)))
            if spec.attn_layout == "internlm":
                kwargs.setdefault("num_kv_attention_heads", kwargs.get("num_attention_heads"))
            super().__init__(spec, **{k: v for k, v in kwargs.items() if k in _LAYER_KEYS})

        def forward(self, hidden_states, residual=None, **kwargs):
            out = super().forward(hidden_states, residual, **kwargs)
            if residual is not None or not self.prenorm or self.is_moe:
                return out
            hidden, res = out                       # stand-alone call: finish the block's residual add here
            return hidden + res

    _Layer.__name__ = _Layer.__qualname__ = name
    _Layer.__doc__ = doc
    return _Layer


def _model_class(name, spec, doc):
    class _Model(PackedDecoder):
        def __init__(self, *args, **kwargs):
            names = [k for k in inspect.signature(PackedDecoder.__init__).parameters if k not in ("self", "spec", "unused")]
            kwargs.update(dict(zip(names, args)))
            if spec.attn_layout == "internlm":
                kwargs.setdefault("num_kv_attention_heads", kwargs.get("num_attention_heads", 12))
            super().__init__(spec, **{k: v for k, v in kwargs.items() if k in _MODEL_KEYS})

    _Model.__name__ = _Model.__qualname__ = name
    _Model.__doc__ = doc
    return _Model


SPECS = {"internlm": INTERNLM_SPEC, "internlm2": INTERNLM2_SPEC, "llama": LLAMA_SPEC}
layer_class, model_class = _layer_class, _model_class
