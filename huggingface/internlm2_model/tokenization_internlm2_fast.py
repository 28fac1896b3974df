"""Fast (``tokenizers``-backed) tokenizer for InternLM2 checkpoints: the SentencePiece BPE model is converted to a
``tokenizers`` pipeline once (``SpmConverter``), BOS / EOS are added by a template post-processor that follows
``add_bos_token`` / ``add_eos_token`` (reference ``transformers/internlm2_model/tokenization_internlm2_fast.py``).  Falls back
to (and shares the vocabulary file with) the slow :class:`InternLM2Tokenizer`."""
import os
from shutil import copyfile
from typing import Optional, Tuple

from tokenizers import decoders, normalizers, processors
from transformers.convert_slow_tokenizer import SLOW_TO_FAST_CONVERTERS, SpmConverter
from transformers.tokenization_utils_fast import PreTrainedTokenizerFast

from .tokenization_internlm2 import InternLM2Tokenizer

VOCAB_FILES_NAMES = {"vocab_file": "./tokenizer.model"}


class InternLM2Converter(SpmConverter):
    """SentencePiece BPE with byte fallback: pieces ``<0xNN>`` are merged back into bytes by the decoder chain."""

    handle_byte_fallback = True

    def vocab(self, proto):
        head = [("<unk>", 0.0), ("<s>", 0.0), ("</s>", 0.0)]
        return head + [(p.piece, p.score) for p in proto.pieces[3:]]

    def unk_id(self, proto):
        return 0

    def decoder(self, replacement, add_prefix_space):
        chain = [decoders.Replace("▁", " "), decoders.ByteFallback(), decoders.Fuse()]
        if self.proto.normalizer_spec.add_dummy_prefix:
            chain.append(decoders.Strip(content=" ", left=1))
        return decoders.Sequence(chain)

    def normalizer(self, proto):
        steps = []
        if proto.normalizer_spec.add_dummy_prefix:
            steps.append(normalizers.Prepend(prepend="▁"))
        steps.append(normalizers.Replace(pattern=" ", content="▁"))
        return normalizers.Sequence(steps)

    def pre_tokenizer(self, replacement, add_prefix_space):
        return None


SLOW_TO_FAST_CONVERTERS["InternLM2Tokenizer"] = InternLM2Converter


class InternLM2TokenizerFast(PreTrainedTokenizerFast):
    vocab_files_names = VOCAB_FILES_NAMES
    slow_tokenizer_class = InternLM2Tokenizer
    padding_side = "left"
    model_input_names = ["input_ids", "attention_mask"]
    _auto_class = "AutoTokenizer"

    def __init__(self, vocab_file, unk_token="<unk>", bos_token="<s>", eos_token="</s>", pad_token="</s>",
                 sp_model_kwargs=None, add_bos_token=True, add_eos_token=False, decode_with_prefix_space=False,
                 clean_up_tokenization_spaces=False, **kwargs):
        super().__init__(vocab_file=vocab_file, unk_token=unk_token, bos_token=bos_token, eos_token=eos_token,
                         pad_token=pad_token, sp_model_kwargs=sp_model_kwargs, add_bos_token=add_bos_token,
                         add_eos_token=add_eos_token, decode_with_prefix_space=decode_with_prefix_space,
                         clean_up_tokenization_spaces=clean_up_tokenization_spaces, **kwargs)
        self._add_bos_token, self._add_eos_token = add_bos_token, add_eos_token
        self.update_post_processor()
        self.vocab_file = vocab_file

    @property
    def can_save_slow_tokenizer(self) -> bool:
        return bool(self.vocab_file) and os.path.isfile(self.vocab_file)

    def update_post_processor(self):
        """``[BOS] A [EOS]`` / ``[BOS] A [EOS] [BOS] B [EOS]`` with the two specials switched by the flags."""
        bos, eos = self.bos_token, self.eos_token
        if self._add_bos_token and self.bos_token_id is None:
            raise ValueError("add_bos_token = True but bos_token = None")
        if self._add_eos_token and self.eos_token_id is None:
            raise ValueError("add_eos_token = True but eos_token = None")
        head = f"{bos}:0 " if self._add_bos_token else ""
        tail = f" {eos}:0" if self._add_eos_token else ""
        head2 = f" {bos}:1" if self._add_bos_token else ""
        tail2 = f" {eos}:1" if self._add_eos_token else ""
        specials = []
        if self._add_bos_token:
            specials.append((bos, self.bos_token_id))
        if self._add_eos_token:
            specials.append((eos, self.eos_token_id))
        self._tokenizer.post_processor = processors.TemplateProcessing(
            single=f"{head}$A:0{tail}", pair=f"{head}$A:0{tail}{head2} $B:1{tail2}", special_tokens=specials)

    @property
    def add_bos_token(self):
        return self._add_bos_token

    @add_bos_token.setter
    def add_bos_token(self, value):
        self._add_bos_token = value
        self.update_post_processor()

    @property
    def add_eos_token(self):
        return self._add_eos_token

    @add_eos_token.setter
    def add_eos_token(self, value):
        self._add_eos_token = value
        self.update_post_processor()

    def save_vocabulary(self, save_directory: str, filename_prefix: Optional[str] = None) -> Tuple[str]:
        if not self.can_save_slow_tokenizer:
            raise ValueError("the SentencePiece model file is needed to save a slow tokenizer's vocabulary")
        os.makedirs(save_directory, exist_ok=True)
        out = os.path.join(save_directory, (filename_prefix + "-" if filename_prefix else "") + "tokenizer.model")
        if os.path.abspath(self.vocab_file) != os.path.abspath(out):
            copyfile(self.vocab_file, out)
        return (out,)
