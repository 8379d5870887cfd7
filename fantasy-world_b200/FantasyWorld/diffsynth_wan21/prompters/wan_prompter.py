"""Mirror of FantasyWorld/diffsynth_wan21/prompters/wan_prompter.py: text cleaning, the HuggingFace tokenizer wrapper and
`WanPrompter.encode_prompt` (tokenise -> umT5 encoder -> zero the padded positions), SURVEY §8f N3.

The tokenizer itself (`AutoTokenizer.from_pretrained(<google/umt5-xxl dir>)`) and `ftfy` are third-party, as in the reference; both
are imported where they are first needed so that the module imports without them (there is no tokenizer checkpoint and no ftfy
in the build environment: tests drive `encode_prompt` with a stand-in tokenizer object).
"""
from __future__ import annotations

import html
import re
import string

import torch

from .base_prompter import BasePrompter

_WS = re.compile(r"\s+")
_NO_PUNCT = str.maketrans("", "", string.punctuation)


def basic_clean(text):
    import ftfy          # same dependency as the reference (wan_prompter.py:5,12): mojibake repair has no stand-in
    return html.unescape(html.unescape(ftfy.fix_text(text))).strip()


def whitespace_clean(text):
    return _WS.sub(" ", text).strip()


def canonicalize(text, keep_punctuation_exact_string=None):
    """Underscores to spaces, punctuation stripped (except an exact marker string), lower-cased, whitespace collapsed."""
    text = text.replace("_", " ")
    pieces = text.split(keep_punctuation_exact_string) if keep_punctuation_exact_string else [text]
    text = (keep_punctuation_exact_string or "").join(piece.translate(_NO_PUNCT) for piece in pieces)
    return whitespace_clean(text.lower())


_CLEANERS = {
    None: lambda t: t,
    "whitespace": lambda t: whitespace_clean(basic_clean(t)),
    "lower": lambda t: whitespace_clean(basic_clean(t)).lower(),
    "canonicalize": lambda t: canonicalize(basic_clean(t)),
}


class HuggingfaceTokenizer:
    """`AutoTokenizer` with a cleaning mode and (optionally) fixed-length padding / truncation (wan_prompter.py:36-80)."""

    def __init__(self, name, seq_len=None, clean=None, **kwargs):
        if clean not in _CLEANERS:
            raise AssertionError(f"clean must be one of {list(_CLEANERS)}")
        from transformers import AutoTokenizer
        self.name, self.seq_len, self.clean = name, seq_len, clean
        self.tokenizer = AutoTokenizer.from_pretrained(name, **kwargs)
        self.vocab_size = self.tokenizer.vocab_size

    def _clean(self, text):
        return _CLEANERS[self.clean](text)

    def __call__(self, sequence, **kwargs):
        want_mask = kwargs.pop("return_mask", False)
        texts = [sequence] if isinstance(sequence, str) else list(sequence)
        options = dict(return_tensors="pt")
        if self.seq_len is not None:
            options.update(padding="max_length", truncation=True, max_length=self.seq_len)
        options.update(kwargs)
        enc = self.tokenizer([self._clean(t) for t in texts], **options)
        return (enc.input_ids, enc.attention_mask) if want_mask else enc.input_ids


class WanPrompter(BasePrompter):
    def __init__(self, tokenizer_path=None, text_len=512):
        super().__init__()
        self.text_len, self.text_encoder, self.tokenizer = text_len, None, None
        self.fetch_tokenizer(tokenizer_path)

    def fetch_tokenizer(self, tokenizer_path=None):
        if tokenizer_path is not None:
            self.tokenizer = HuggingfaceTokenizer(name=tokenizer_path, seq_len=self.text_len, clean="whitespace")

    def fetch_models(self, text_encoder=None):
        self.text_encoder = text_encoder

    @torch.no_grad()
    def encode_prompt(self, prompt, positive=True, device="cuda"):
        if self.tokenizer is None or self.text_encoder is None:
            raise RuntimeError("WanPrompter.encode_prompt: fetch_tokenizer(path) and fetch_models(text_encoder) first")
        ids, mask = self.tokenizer(self.process_prompt(prompt, positive=positive), return_mask=True, add_special_tokens=True)
        mask = mask.to(device)
        emb = self.text_encoder(ids.to(device), mask)
        # the reference zeroes [:, v:] for EVERY sequence's length v in turn (wan_prompter.py:106-108), i.e. from the shortest
        # length on, for the whole batch; with the single prompt the sampler passes this is "zero the padding"
        shortest = int((mask > 0).sum(dim=1).min())
        emb[:, shortest:] = 0
        return emb
