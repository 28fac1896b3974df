from . import families  # noqa: F401  (registers the model builders)
from .decoder import DecoderLayer, PackedDecoder, build_generic_model_1d
from .modules import (
    MHA,
    DynamicNTKScalingRotaryEmbedding,
    Embedding1D,
    FeedForward,
    LinearRotaryEmbedding,
    RotaryEmbedding,
    VocabParallelEmbedding,
)

__all__ = ["DecoderLayer", "PackedDecoder", "build_generic_model_1d", "MHA", "Embedding1D", "FeedForward",
           "RotaryEmbedding", "LinearRotaryEmbedding", "DynamicNTKScalingRotaryEmbedding", "VocabParallelEmbedding"]
