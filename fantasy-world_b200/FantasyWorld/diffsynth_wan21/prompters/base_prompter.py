"""Mirror of FantasyWorld/diffsynth_wan21/prompters/base_prompter.py: prompt refiner / extender plumbing (host logic only)."""
from __future__ import annotations

import torch


def tokenize_long_prompt(tokenizer, prompt, max_length=None):
    """Tokenise a prompt longer than the tokenizer's window into [num_sentences, window] ids (base_prompter.py:6-36): the padded
    length is the next multiple of the window."""
    window = tokenizer.model_max_length if max_length is None else max_length
    tokenizer.model_max_length = 99999999            # silence the "longer than the maximum" warning for the probe pass
    try:
        n_tokens = tokenizer(prompt, return_tensors="pt").input_ids.shape[1]
    finally:
        tokenizer.model_max_length = window
    padded = (n_tokens + window - 1) // window * window
    ids = tokenizer(prompt, return_tensors="pt", padding="max_length", max_length=padded, truncation=True).input_ids
    return ids.reshape(ids.shape[1] // window, window)


class BasePrompter:
    def __init__(self):
        self.refiners = []
        self.extenders = []

    def load_prompt_refiners(self, model_manager, refiner_classes=[]):
        self.refiners.extend(cls.from_model_manager(model_manager) for cls in refiner_classes)

    def load_prompt_extenders(self, model_manager, extender_classes=[]):
        self.extenders.extend(cls.from_model_manager(model_manager) for cls in extender_classes)

    @torch.no_grad()
    def process_prompt(self, prompt, positive=True):
        if isinstance(prompt, list):
            return [self.process_prompt(p, positive=positive) for p in prompt]
        for refiner in self.refiners:
            prompt = refiner(prompt, positive=positive)
        return prompt

    @torch.no_grad()
    def extend_prompt(self, prompt: str, positive=True):
        extended = dict(prompt=prompt)
        for extender in self.extenders:
            extended = extender(extended)
        return extended
