from .configuration_internlm2 import InternLM2Config
from .modeling_internlm2 import InternLM2ForCausalLM, InternLM2ForSequenceClassification, InternLM2Model

__all__ = ["InternLM2Config", "InternLM2Model", "InternLM2ForCausalLM", "InternLM2ForSequenceClassification"]
