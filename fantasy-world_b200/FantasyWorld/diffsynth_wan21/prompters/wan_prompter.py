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


def basic_clean(text):
    import ftfy          # same dependency as the reference (wan_prompter.py:5,12): mojibake repair has no stand-in
    text = ftfy.fix_text(text)
    return html.unescape(html.unescape(text)).strip()


def whitespace_clean(text):
    return re.sub(r"\s+", " ", text).strip()


def canonicalize(text, keep_punctuation_exact_string=None):
    """Underscores to spaces, punctuation stripped (except an exact marker string), lower-cased, whitespace collapsed."""
    strip = str.maketrans("", "", string.punctuation)
    text = text.replace("_", " ")
    if keep_punctuation_exact_string:
        text = keep_punctuation_exact_string.join(part.translate(strip) for part in text.split(keep_punctuation_exact_string))
    else:
        text = text.translate(strip)
    return re.sub(r"\s+", " ", text.lower()).strip()


class HuggingfaceTokenizer:
    def __init__(self, name, seq_len=None, clean=None, **kwargs):
        assert clean in (None, "whitespace", "lower", "canonicalize")
        self.name, self.seq_len, self.clean = name, seq_len, clean
        from transformers import AutoTokenizer
        self.tokenizer = AutoTokenizer.from_pretrained(name, **kwargs)
        self.vocab_size = self.tokenizer.vocab_size

    def __call__(self, sequence, **kwargs):
        return_mask = kwargs.pop("return_mask", False)
        call = {"return_tensors": "pt"}
        if self.seq_len is not None:
            call.update(padding="max_length", truncation=True, max_length=self.seq_len)
        call.update(**kwargs)
        if isinstance(sequence, str):
            sequence = [sequence]
        if self.clean:
            sequence = [self._clean(u) for u in sequence]
        enc = self.tokenizer(sequence, **call)
        return (enc.input_ids, enc.attention_mask) if return_mask else enc.input_ids

    def _clean(self, text):
        if self.clean == "whitespace":
            return whitespace_clean(basic_clean(text))
        if self.clean == "lower":
            return whitespace_clean(basic_clean(text)).lower()
        if self.clean == "canonicalize":
            return canonicalize(basic_clean(text))
        return text


class WanPrompter(BasePrompter):
    def __init__(self, tokenizer_path=None, text_len=512):
        super().__init__()
        self.text_len = text_len
        self.text_encoder = None
        self.tokenizer = None
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
        prompt = self.process_prompt(prompt, positive=positive)
        ids, mask = self.tokenizer(prompt, return_mask=True, add_special_tokens=True)
        ids, mask = ids.to(device), mask.to(device)
        emb = self.text_encoder(ids, mask)
        # the reference zeroes [:, v:] for EVERY sequence's length v in turn (wan_prompter.py:106-108), i.e. from the shortest
        # length on, for the whole batch; with the single prompt the sampler passes this is "zero the padding"
        emb[:, int(mask.gt(0).sum(dim=1).min()):] = 0
        return emb
