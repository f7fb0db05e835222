"""Caption tokenisation for the text branch (host side; reference models/transformer.py:59,129).

The reference builds `RobertaTokenizerFast.from_pretrained("roberta-base")`, which needs vocabulary files from the
HuggingFace hub.  `build_tokenizer` returns that tokenizer when the files are available locally and otherwise the
deterministic `CharTokenizer` below (one id per character between <s> and </s>), which is what the synthetic
benchmark, the oracle and the golden fixtures use: a caption of n characters yields exactly n + 2 ids and
`char_to_token(i, c) == c + 1`, as SURVEY.md §8(d) prescribes.
"""
from __future__ import annotations

from collections import UserDict
from typing import Dict, List, Optional

import torch

BOS, PAD, EOS = 0, 1, 2
_CHAR_BASE = 4


class TokenBatch(UserDict):
    """Minimal BatchEncoding work-alike: mapping of tensors + `.to()`, attribute access and `char_to_token`.
    A UserDict like transformers.BatchEncoding, NOT a dict subclass: DistributedDataParallel rebuilds every dict it
    finds among the forward arguments (`type(obj)(items)`), which would drop the caption lengths that `char_to_token`
    needs when `memory_cache["tokenized"]` travels through the wrapped model (engine.py:66)."""

    def __init__(self, data: Dict[str, torch.Tensor], lengths: List[int]):
        super().__init__(data)
        self._lengths = list(lengths)

    def __getattr__(self, item):
        if item == "data":  # UserDict storage, not yet set during construction / unpickling
            raise AttributeError(item)
        try:
            return self.data[item]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(item) from e

    def to(self, device) -> "TokenBatch":
        from .util.misc import h2d

        return TokenBatch({k: h2d(v, device) for k, v in self.items()}, self._lengths)

    def char_to_token(self, batch_or_char_index: int, char_index: Optional[int] = None) -> Optional[int]:
        if char_index is None:
            batch_index, char_index = 0, batch_or_char_index
        else:
            batch_index = batch_or_char_index
        if 0 <= char_index < self._lengths[batch_index]:
            return char_index + 1
        return None

    def __deepcopy__(self, memo):
        return TokenBatch({k: v.clone() for k, v in self.items()}, self._lengths)


class CharTokenizer:
    """One token per character, RoBERTa special ids (<s>=0, <pad>=1, </s>=2), right padding to the longest."""

    vocab_size = 50265
    pad_token_id = PAD

    def __call__(self, text: List[str], padding="longest", return_tensors="pt") -> TokenBatch:
        assert return_tensors == "pt"
        if isinstance(text, str):
            text = [text]
        ids = [[BOS] + [_CHAR_BASE + (ord(ch) % (self.vocab_size - _CHAR_BASE)) for ch in t] + [EOS] for t in text]
        longest = max(len(s) for s in ids)
        input_ids = torch.full((len(ids), longest), PAD, dtype=torch.long)
        attention = torch.zeros((len(ids), longest), dtype=torch.long)
        for i, s in enumerate(ids):
            input_ids[i, : len(s)] = torch.tensor(s, dtype=torch.long)
            attention[i, : len(s)] = 1
        return TokenBatch({"input_ids": input_ids, "attention_mask": attention}, [len(t) for t in text])

    batch_encode_plus = __call__


class _HFTokenizer:
    """Adapter giving a real RobertaTokenizerFast the call surface used by the text branch."""

    def __init__(self, tok):
        self._tok = tok

    def __call__(self, text, padding="longest", return_tensors="pt"):
        return self._tok(text, padding=padding, return_tensors=return_tensors)

    batch_encode_plus = __call__


def build_tokenizer(text_encoder_type: str = "roberta-base", synthetic: Optional[bool] = None):
    """`synthetic=None` tries the local HuggingFace files first and falls back to CharTokenizer (no network here)."""
    if synthetic:
        return CharTokenizer()
    try:
        from transformers import RobertaTokenizerFast

        return _HFTokenizer(RobertaTokenizerFast.from_pretrained(text_encoder_type, local_files_only=True))
    except Exception:
        if synthetic is False:
            raise
        return CharTokenizer()
