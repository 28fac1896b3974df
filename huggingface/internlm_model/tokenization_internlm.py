"""InternLM (v1) tokenizer: the same SentencePiece wrapper as InternLM2 (reference ``tokenization_internlm.py``)."""
from ..internlm2_model.tokenization_internlm2 import InternLM2Tokenizer


class InternLMTokenizer(InternLM2Tokenizer):
    pass
