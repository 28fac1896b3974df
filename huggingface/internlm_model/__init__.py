from .configuration_internlm import InternLMConfig
from .modeling_internlm import InternLMForCausalLM, InternLMModel

__all__ = ["InternLMConfig", "InternLMModel", "InternLMForCausalLM"]
