"""Mirror of FantasyWorld/diffsynth_wan21/prompters/base_prompter.py: prompt refiner / extender plumbing (host logic only)."""
from __future__ import annotations

import contextlib

import torch


@contextlib.contextmanager
def _unbounded_window(tokenizer):
    """Lift `model_max_length` for a probing pass (silences the tokenizer's over-length warning), then restore it."""
    saved = tokenizer.model_max_length
    tokenizer.model_max_length = 99999999
    try:
        yield saved
    finally:
        tokenizer.model_max_length = saved


def tokenize_long_prompt(tokenizer, prompt, max_length=None):
    """Ids of a prompt that may exceed the tokenizer's window, as [num_windows, window]: padded up to the next multiple of the
    window (base_prompter.py:6-36)."""
    with _unbounded_window(tokenizer) as native:
        window = native if max_length is None else max_length
        n_tokens = tokenizer(prompt, return_tensors="pt").input_ids.shape[1]
    n_windows = max(1, -(-n_tokens // window))
    ids = tokenizer(prompt, return_tensors="pt", padding="max_length", max_length=n_windows * window, truncation=True).input_ids
    return ids.reshape(n_windows, window)


class BasePrompter:
    """Holds optional prompt refiners (text -> text) and extenders (dict -> dict) built from a model manager."""

    def __init__(self):
        self.refiners, self.extenders = [], []

    def _load(self, bucket, model_manager, classes):
        bucket.extend(cls.from_model_manager(model_manager) for cls in classes)

    def load_prompt_refiners(self, model_manager, refiner_classes=[]):
        self._load(self.refiners, model_manager, refiner_classes)

    def load_prompt_extenders(self, model_manager, extender_classes=[]):
        self._load(self.extenders, model_manager, extender_classes)

    @torch.no_grad()
    def process_prompt(self, prompt, positive=True):
        if isinstance(prompt, (list, tuple)):
            return [self.process_prompt(one, positive=positive) for one in prompt]
        for refine in self.refiners:
            prompt = refine(prompt, positive=positive)
        return prompt

    @torch.no_grad()
    def extend_prompt(self, prompt: str, positive=True):
        state = {"prompt": prompt}
        for extend in self.extenders:
            state = extend(state)
        return state
