"""Minimal stand-in for FantasyWorld/diffsynth_wan21/pipelines/wan_video.py: the attribute surface
FantasyWorldFusionModel and inference_wan21.py touch (`dit`, `scheduler`, `device`, `torch_dtype`, `generate_noise`,
`prepare_extra_input`, `load_models_to_device`, `vae`).  The VAE (`pipe.vae.decode(..., tiled=True)`) is the
mirror in ..models.wan_video_vae (attach with `enable_vae()` or by loading a checkpoint through the model manager).  The text /
image encoders (`encode_prompt`, `encode_image`: T5, CLIP) run once per sample outside the denoising loop and stay with the
reference: they raise with a clear message instead of silently doing something else.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..schedulers.flow_match import FlowMatchScheduler


class WanVideoPipeline(nn.Module):
    def __init__(self, device="cuda", torch_dtype=torch.bfloat16, tokenizer_path=None):
        super().__init__()
        self.device, self.torch_dtype = device, torch_dtype
        self.scheduler = FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        self.text_encoder = None
        self.image_encoder = None
        self.dit = None
        self.vae = None
        self.model_names = ['text_encoder', 'dit', 'vae', 'image_encoder']
        self.cpu_offload = False

    @staticmethod
    def from_model_manager(model_manager, torch_dtype=None, device=None):
        pipe = WanVideoPipeline(device=device or model_manager.device, torch_dtype=torch_dtype or model_manager.torch_dtype)
        pipe.dit = model_manager.fetch_model("wan_video_dit")
        return pipe

    def enable_vae(self, z_dim: int = 16, state_dict=None, device=None, dtype=None):
        """Attach the Wan VAE mirror (random-init unless a state_dict with the reference's `model.*` keys is given)."""
        from ..models.wan_video_vae import WanVideoVAE
        self.vae = WanVideoVAE(z_dim=z_dim)
        if state_dict is not None:
            self.vae.load_state_dict(state_dict, strict=True)
        self.vae.to(device=device or self.device, dtype=dtype or self.torch_dtype)
        return self.vae

    def generate_noise(self, shape, seed=None, device="cpu", dtype=torch.float16):
        gen = None if seed is None else torch.Generator(device).manual_seed(seed)
        return torch.randn(shape, generator=gen, device=device, dtype=dtype)

    def prepare_extra_input(self, latents=None):
        return {}

    def load_models_to_device(self, loadmodel_names=[]):
        return None  # no CPU offload: 37 GB of weights stay resident in 180 GB of HBM

    def _out_of_scope(self, what):
        raise NotImplementedError(f"{what} is outside the B200 hot-path build (runs once per sample; see DESIGN.md §scope)")

    def encode_prompt(self, *a, **k):
        self._out_of_scope("T5 prompt encoding")

    def encode_image(self, *a, **k):
        self._out_of_scope("CLIP / VAE image encoding")
