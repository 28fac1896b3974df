from internevo_b200.parallel.functional import gather_forward_split_backward
from internevo_b200.parallel.linear import BaseScaleColumnParallelLinear, RewardModelLinear, ScaleColumnParallelLinear

from . import families  # noqa: F401  (registers the model builders)
from .decoder import DecoderLayer, PackedDecoder, build_generic_model_1d
from .families import (
    build_model_with_cfg,
    build_model_with_cfg_internlm2,
    build_model_with_cfg_llama,
    build_model_with_moe_cfg,
)
from .metrics import AccPerplex
from .modules import (
    MHA,
    DynamicNTKScalingRotaryEmbedding,
    Embedding1D,
    FeedForward,
    LinearRotaryEmbedding,
    RotaryEmbedding,
    VocabParallelEmbedding,
)
from .moe import MoE

__all__ = ["DecoderLayer", "PackedDecoder", "build_generic_model_1d", "MHA", "Embedding1D", "FeedForward",
           "RotaryEmbedding", "LinearRotaryEmbedding", "DynamicNTKScalingRotaryEmbedding", "VocabParallelEmbedding",
           "MoE", "AccPerplex", "ScaleColumnParallelLinear", "BaseScaleColumnParallelLinear", "RewardModelLinear", "gather_forward_split_backward",
           "build_model_with_cfg", "build_model_with_cfg_internlm2", "build_model_with_cfg_llama",
           "build_model_with_moe_cfg"]
