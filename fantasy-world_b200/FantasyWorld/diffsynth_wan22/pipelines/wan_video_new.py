"""Minimal stand-in for FantasyWorld/diffsynth_wan22/pipelines/wan_video_new.py: what FantasyWorldFusionModel (Wan2.2) and
inference_wan22.py touch on the denoising path — `ModelConfig`, `WanVideoPipeline.from_pretrained(...)` returning an
object with `dit`, `scheduler`, `device`, `torch_dtype`.  Model download / hash detection / T5 / VAE units of the reference
pipeline (~2800 lines) are I/O and once-per-sample work, out of scope.  With no checkpoint on disk the DiT is
random-initialised at the Wan2.2-Fun-A14B-Control-Camera configuration (what the benchmarks use).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from ..models.wan_video_dit import WAN22_FUN_A14B_CONTROL_CAMERA, WanModel
from ..schedulers.flow_match import FlowMatchScheduler


@dataclass
class ModelConfig:
    path: Optional[str] = None
    model_id: Optional[str] = None
    origin_file_pattern: Optional[str] = None
    local_model_path: Optional[str] = None
    offload_device: Optional[str] = None
    offload_dtype: Optional[torch.dtype] = None


class WanVideoPipeline(nn.Module):
    def __init__(self, device="cuda", torch_dtype=torch.bfloat16, tokenizer_path=None):
        super().__init__()
        self.device, self.torch_dtype = device, torch_dtype
        self.scheduler = FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        self.dit = None
        self.text_encoder = self.vae = self.image_encoder = None

    @staticmethod
    def from_pretrained(torch_dtype=torch.bfloat16, device="cuda", model_configs=(), tokenizer_config=None, dit_config=None, **kw):
        pipe = WanVideoPipeline(device=device, torch_dtype=torch_dtype)
        cfg = dict(dit_config or WAN22_FUN_A14B_CONTROL_CAMERA)
        pipe.dit = WanModel(**cfg).to(torch_dtype)
        import glob
        import os
        for mc in model_configs:
            root = mc.local_model_path or mc.path
            if root and mc.origin_file_pattern and "diffusion_pytorch_model" in mc.origin_file_pattern:
                files = sorted(glob.glob(os.path.join(root, mc.origin_file_pattern)))
                if files:
                    from safetensors.torch import load_file
                    sd = {}
                    for f in files:
                        sd.update(load_file(f, device="cpu"))
                    keys = set(pipe.dit.state_dict().keys())
                    pipe.dit.load_state_dict({k: v for k, v in sd.items() if k in keys}, strict=False)
        return pipe

    def generate_noise(self, shape, seed=None, device="cpu", dtype=torch.float16):
        gen = None if seed is None else torch.Generator(device).manual_seed(seed)
        return torch.randn(shape, generator=gen, device=device, dtype=dtype)

    def load_models_to_device(self, names=()):
        return None
