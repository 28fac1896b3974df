"""HF configuration of InternLM2 checkpoints written by ``tools/convert2hf.py`` (``model_type = "internlm2"``).

Field names follow the published InternLM2 HF checkpoints (reference ``transformers/internlm2_model/configuration_internlm2.py``)
so that ``config.json`` files are interchangeable."""
from transformers.configuration_utils import PretrainedConfig


class InternLM2Config(PretrainedConfig):
    model_type = "internlm2"
    keys_to_ignore_at_inference = ["past_key_values"]

    def __init__(self, vocab_size=103168, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=None, hidden_act="silu", max_position_embeddings=2048,
                 initializer_range=0.02, rms_norm_eps=1e-6, use_cache=True, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                 tie_word_embeddings=False, bias=False, rope_theta=10000, rope_scaling=None, attn_implementation="eager",
                 **kwargs):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_key_value_heads = num_key_value_heads or num_attention_heads
        self.hidden_act = hidden_act
        self.max_position_embeddings = max_position_embeddings
        self.initializer_range = initializer_range
        self.rms_norm_eps = rms_norm_eps
        self.use_cache = use_cache
        self.bias = bias
        self.rope_theta = rope_theta
        self.rope_scaling = rope_scaling
        self._validate_rope_scaling()
        self.attn_implementation = attn_implementation or "eager"
        super().__init__(pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id,
                         tie_word_embeddings=tie_word_embeddings, **kwargs)

    def _validate_rope_scaling(self):
        if self.rope_scaling is None:
            return
        if not isinstance(self.rope_scaling, dict):
            raise ValueError(f"`rope_scaling` must be a dict with `type` and `factor`, got {self.rope_scaling}")
        kind = self.rope_scaling.get("type", self.rope_scaling.get("rope_type"))
        factor = self.rope_scaling.get("factor", 1.0)
        if kind not in ("linear", "dynamic", "default", None):
            raise ValueError(f"`rope_scaling` type must be 'linear' or 'dynamic', got {kind}")
        if not isinstance(factor, (int, float)) or factor < 1.0:
            raise ValueError(f"`rope_scaling` factor must be a number >= 1, got {factor}")
