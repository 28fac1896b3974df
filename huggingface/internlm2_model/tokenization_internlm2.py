"""SentencePiece tokenizer class for converted InternLM / InternLM2 checkpoints (reference
``transformers/internlm2_model/tokenization_internlm2.py`` and ``internlm_model/tokenization_internlm.py``): BOS is
prepended (``add_bos_token``), EOS optionally appended, byte-fallback pieces are decoded by SentencePiece itself."""
import os
from shutil import copyfile
from typing import Any, Dict, List, Optional, Tuple

import sentencepiece as spm
from transformers.tokenization_utils import PreTrainedTokenizer

VOCAB_FILES_NAMES = {"vocab_file": "./tokenizer.model"}


class InternLM2Tokenizer(PreTrainedTokenizer):
    vocab_files_names = VOCAB_FILES_NAMES
    model_input_names = ["input_ids", "attention_mask"]
    _auto_class = "AutoTokenizer"

    def __init__(self, vocab_file, unk_token="<unk>", bos_token="<s>", eos_token="</s>", pad_token="</s>",
                 sp_model_kwargs: Optional[Dict[str, Any]] = None, add_bos_token=True, add_eos_token=False,
                 decode_with_prefix_space=False, clean_up_tokenization_spaces=False, **kwargs):
        self.sp_model_kwargs = {} if sp_model_kwargs is None else sp_model_kwargs
        self.vocab_file = vocab_file
        self.add_bos_token, self.add_eos_token = add_bos_token, add_eos_token
        self.decode_with_prefix_space = decode_with_prefix_space
        self.sp_model = spm.SentencePieceProcessor(**self.sp_model_kwargs)
        self.sp_model.Load(vocab_file)
        self._no_prefix_space_tokens = None
        super().__init__(bos_token=bos_token, eos_token=eos_token, unk_token=unk_token, pad_token=pad_token,
                         clean_up_tokenization_spaces=clean_up_tokenization_spaces, **kwargs)

    # ---- vocabulary
    @property
    def vocab_size(self):
        return self.sp_model.get_piece_size()

    @property
    def bos_token_id(self) -> Optional[int]:
        return self.sp_model.bos_id()

    @property
    def eos_token_id(self) -> Optional[int]:
        return self.sp_model.eos_id()

    def get_vocab(self):
        vocab = {self.convert_ids_to_tokens(i): i for i in range(self.vocab_size)}
        vocab.update(self.added_tokens_encoder)
        return vocab

    def _tokenize(self, text):
        return self.sp_model.encode(text, out_type=str)

    def _convert_token_to_id(self, token):
        return self.sp_model.piece_to_id(token)

    def _convert_id_to_token(self, index):
        return self.sp_model.IdToPiece(index)

    @property
    def no_prefix_space_tokens(self):
        if self._no_prefix_space_tokens is None:
            vocab = self.convert_ids_to_tokens(list(range(self.vocab_size)))
            self._no_prefix_space_tokens = {i for i, tok in enumerate(vocab) if not tok.startswith("▁")}
        return self._no_prefix_space_tokens

    def _maybe_add_prefix_space(self, tokens, decoded):
        if tokens and tokens[0] not in self.no_prefix_space_tokens:
            return " " + decoded
        return decoded

    def convert_tokens_to_string(self, tokens):
        current, out, prev_special = [], "", False
        for token in tokens:
            if token in self.all_special_tokens:      # special tokens are never fed to SentencePiece
                if not prev_special:
                    out += " "
                out += self.sp_model.decode(current) + token
                prev_special, current = True, []
            else:
                current.append(token)
                prev_special = False
        out += self.sp_model.decode(current)
        out = self.clean_up_tokenization(out) if hasattr(self, "clean_up_tokenization") and self.clean_up_tokenization_spaces else out
        out = self._maybe_add_prefix_space(tokens=tokens, decoded=out) if self.decode_with_prefix_space else out
        return out[1:] if out.startswith(" ") and prev_special is False and tokens and tokens[0] in self.all_special_tokens else out

    # ---- files
    def save_vocabulary(self, save_directory, filename_prefix: Optional[str] = None) -> Tuple[str]:
        if not os.path.isdir(save_directory):
            raise ValueError(f"Vocabulary path ({save_directory}) should be a directory")
        out = os.path.join(save_directory, (filename_prefix + "-" if filename_prefix else "") + "tokenizer.model")
        if os.path.abspath(self.vocab_file) != os.path.abspath(out) and os.path.isfile(self.vocab_file):
            copyfile(self.vocab_file, out)
        elif not os.path.isfile(self.vocab_file):
            with open(out, "wb") as f:
                f.write(self.sp_model.serialized_model_proto())
        return (out,)

    # ---- special tokens
    def build_inputs_with_special_tokens(self, token_ids_0, token_ids_1=None):
        bos = [self.bos_token_id] if self.add_bos_token else []
        out = bos + token_ids_0
        if token_ids_1 is not None:
            out = out + token_ids_1
        if self.add_eos_token:
            out = out + [self.eos_token_id]
        return out

    def get_special_tokens_mask(self, token_ids_0: List[int], token_ids_1: Optional[List[int]] = None,
                                already_has_special_tokens: bool = False) -> List[int]:
        if already_has_special_tokens:
            return super().get_special_tokens_mask(token_ids_0=token_ids_0, token_ids_1=token_ids_1,
                                                   already_has_special_tokens=True)
        bos = [1] if self.add_bos_token else []
        eos = [1] if self.add_eos_token else []
        if token_ids_1 is None:
            return bos + [0] * len(token_ids_0) + eos
        return bos + [0] * len(token_ids_0) + [0] * len(token_ids_1) + eos

    def create_token_type_ids_from_sequences(self, token_ids_0, token_ids_1=None):
        n = len(token_ids_0) + (len(token_ids_1) if token_ids_1 else 0) + int(self.add_bos_token) + int(self.add_eos_token)
        return [0] * n
