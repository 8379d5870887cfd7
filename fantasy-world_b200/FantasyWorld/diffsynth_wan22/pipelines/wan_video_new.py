"""Stand-in for FantasyWorld/diffsynth_wan22/pipelines/wan_video_new.py: what FantasyWorldFusionModel (Wan2.2) and
inference_wan22.py touch — `ModelConfig`, `WanVideoPipeline.from_pretrained(...)` returning an object with `dit`, `scheduler`,
`device`, `torch_dtype`, `vae`, `text_encoder`, `prompter`, `generate_noise`, and the CONDITIONING call
`pipe(prompt=..., negative_prompt=..., input_image=..., end_image=..., seed=..., tiled=True, height=..., width=..., return_condition=True)
-> (inputs_shared, inputs_posi, inputs_nega)` (inference_wan22.py:345-353) that yields `context` for both prompts and `y`.

The reference pipeline is a list of 22 "units" run over three dicts (wan_video_new.py:51-74, 497-533); with the inputs the FantasyWorld CLI
passes, the ones that do anything are the shape check, the noise initialiser, the prompt embedder (umT5) and the VAE image embedder;
the others return {} (no audio, VACE, reference / control video, CLIP — Wan2.2 has no image encoder —, TeaCache, ...).  Those four are
restated here as methods, with the reference's arithmetic (image scaling in the pipeline dtype, mask folding, tiled VAE encode).  The
denoising half of `__call__` is not mirrored: FantasyWorld drives its own dual-expert loop (inference_wan22.py:163-300,
fusion/model_wan22.py).  Model download / hash detection are I/O and out of scope: checkpoints are resolved on local disk
only.  With no DiT checkpoint the DiT is random-initialised at the Wan2.2-Fun-A14B-Control-Camera configuration (what the benchmarks use).
"""
from __future__ import annotations

import glob
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from ...diffsynth_wan21.prompters import WanPrompter
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

    def local_files(self):
        """Files / directories this config names that exist on local disk (no download): `path`, else the pattern under
        `local_model_path[/model_id]`, else under ./models/<model_id> (where the reference's downloader puts them)."""
        if self.path:
            return sorted(self.path) if isinstance(self.path, (list, tuple)) else [self.path]
        if not self.origin_file_pattern:
            return []
        roots = []
        if self.local_model_path:
            roots += [self.local_model_path] + ([os.path.join(self.local_model_path, self.model_id)] if self.model_id else [])
        if self.model_id:
            roots.append(os.path.join("models", self.model_id))
        for root in roots:
            hits = sorted(glob.glob(os.path.join(root, self.origin_file_pattern)))
            if hits:
                return hits
        return []


_UNSUPPORTED_INPUTS = ("input_video", "input_audio", "audio_embeds", "s2v_pose_video", "s2v_pose_latents", "motion_video", "control_video",
                       "reference_image", "camera_control_direction", "cameras_interp", "vace_video", "vace_video_mask",
                       "vace_reference_image", "animate_pose_video", "animate_face_video", "animate_inpaint_video", "animate_mask_video",
                       "vap_video", "longcat_video", "motion_bucket_id", "sliding_window_size", "sliding_window_stride",
                       "tea_cache_l1_thresh")


class WanVideoPipeline(nn.Module):
    def __init__(self, device="cuda", torch_dtype=torch.bfloat16, tokenizer_path=None):
        super().__init__()
        self.device, self.torch_dtype = device, torch_dtype
        self.height_division_factor = self.width_division_factor = 16
        self.time_division_factor, self.time_division_remainder = 4, 1
        self.scheduler = FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        self.prompter = WanPrompter(tokenizer_path=tokenizer_path)
        self.dit = self.dit2 = None
        self.text_encoder = self.vae = self.image_encoder = None

    @staticmethod
    def from_pretrained(torch_dtype=torch.bfloat16, device="cuda", model_configs=(),
                        tokenizer_config: Optional[ModelConfig] = ModelConfig(model_id="Wan-AI/Wan2.1-T2V-1.3B", origin_file_pattern="google/*"),
                        dit_config=None, side_configs=None, **kw):
        """wan_video_new.py:344-418 on local files: DiT safetensors shards, `models_t5_umt5-xxl-enc-bf16.pth`, `Wan2.1_VAE.pth`, and the
        tokenizer directory.  `dit_config` / `side_configs` (extensions): reduced configurations for tests."""
        from ...diffsynth_wan21.models.model_manager import ModelManager
        pipe = WanVideoPipeline(device=device, torch_dtype=torch_dtype)
        cfg = dict(dit_config or WAN22_FUN_A14B_CONTROL_CAMERA)
        pipe.dit = WanModel(**cfg).to(torch_dtype)
        sides = ModelManager(torch_dtype=torch_dtype, device="cpu", side_configs=side_configs)
        for mc in model_configs:
            files = mc.local_files()
            shards = [f for f in files if f.endswith(".safetensors")]
            if shards:
                from safetensors.torch import load_file
                sd = {}
                for f in shards:
                    sd.update(load_file(f, device="cpu"))
                keys = set(pipe.dit.state_dict().keys())
                pipe.dit.load_state_dict({k: v for k, v in sd.items() if k in keys}, strict=False)
            for f in files:
                if f.endswith((".pth", ".pt", ".ckpt")):
                    sides.load_state_dict_model(torch.load(f, map_location="cpu", weights_only=True), path=f,
                                                torch_dtype=mc.offload_dtype or torch_dtype, device=mc.offload_device or "cpu")
        pipe.text_encoder = sides.fetch_model("wan_video_text_encoder")
        pipe.vae = sides.fetch_model("wan_video_vae")
        pipe.image_encoder = sides.fetch_model("wan_video_image_encoder")
        if pipe.vae is not None:
            pipe.height_division_factor = pipe.width_division_factor = pipe.vae.upsampling_factor * 2
        if tokenizer_config:
            pipe.prompter.fetch_models(pipe.text_encoder)
            found = [p for p in tokenizer_config.local_files() if os.path.isdir(p)]
            if found:
                pipe.prompter.fetch_tokenizer(found[0])
        return pipe

    # -- helpers of the reference's BasePipeline (diffsynth_wan22/utils/__init__.py:44-67, 118-123) --------------------------------
    def check_resize_height_width(self, height, width, num_frames=None):
        def up(v, q):
            return (v + q - 1) // q * q
        height, width = up(height, self.height_division_factor), up(width, self.width_division_factor)
        if num_frames is None:
            return height, width
        if num_frames % self.time_division_factor != self.time_division_remainder:
            num_frames = up(num_frames, self.time_division_factor) + self.time_division_remainder
        return height, width, num_frames

    def preprocess_image(self, image, torch_dtype=None, device=None, pattern="B C H W", min_value=-1, max_value=1):
        """PIL image -> [1, C, H, W] in [min, max]; the scaling runs in the pipeline dtype (bf16), as in the reference."""
        if pattern != "B C H W":
            raise NotImplementedError("preprocess_image: only the 'B C H W' layout is used on this path")
        x = torch.from_numpy(np.array(image, dtype=np.float32)).to(dtype=torch_dtype or self.torch_dtype, device=device or self.device)
        x = x * ((max_value - min_value) / 255) + min_value
        return x.permute(2, 0, 1).unsqueeze(0)

    def generate_noise(self, shape, seed=None, rand_device="cpu", rand_torch_dtype=torch.float32, device=None, torch_dtype=None):
        gen = None if seed is None else torch.Generator(rand_device).manual_seed(seed)
        noise = torch.randn(shape, generator=gen, device=rand_device, dtype=rand_torch_dtype)
        return noise.to(dtype=torch_dtype or self.torch_dtype, device=device or self.device)

    def load_models_to_device(self, names=()):
        return None      # no CPU offload: everything stays resident in 180 GB of HBM

    # -- the conditioning units (wan_video_new.py:723-745, 777-790, 856-893) ---------------------------------------------------------
    def embed_prompt(self, prompt, positive=True):
        if self.text_encoder is None or self.prompter.tokenizer is None:
            raise RuntimeError("WanVideoPipeline: no umT5 text encoder / tokenizer loaded (load_text_encoder=True and a local "
                               "google/umt5-xxl tokenizer directory are needed for prompt conditioning)")
        return self.prompter.encode_prompt(prompt, positive=positive, device=self.device)

    @torch.no_grad()
    def embed_image_vae(self, input_image, end_image, num_frames, height, width, tiled, tile_size, tile_stride):
        """`y`: 4 known-frame mask channels (4 pixel frames folded per latent frame, the first one repeated) stacked on the VAE
        encoding of the clip whose unknown frames are zero."""
        if self.vae is None:
            raise RuntimeError("WanVideoPipeline: no VAE loaded (load_vae=True) for the image conditioning")
        dev = self.device
        first = self.preprocess_image(input_image.resize((width, height))).to(dev)
        known = torch.zeros(1, num_frames, height // 8, width // 8, device=dev)
        known[:, 0] = 1
        if end_image is not None:
            last = self.preprocess_image(end_image.resize((width, height))).to(dev)
            clip = torch.cat([first.transpose(0, 1), torch.zeros(3, num_frames - 2, height, width, device=dev, dtype=first.dtype),
                              last.transpose(0, 1)], dim=1)
            known[:, -1] = 1
        else:
            clip = torch.cat([first.transpose(0, 1), torch.zeros(3, num_frames - 1, height, width, device=dev, dtype=first.dtype)], dim=1)
        known = torch.cat([known[:, :1].repeat_interleave(4, dim=1), known[:, 1:]], dim=1)
        known = known.view(1, known.shape[1] // 4, 4, height // 8, width // 8).transpose(1, 2)[0]
        z = self.vae.encode([clip.to(dtype=self.torch_dtype, device=dev)], device=dev, tiled=tiled, tile_size=tile_size,
                            tile_stride=tile_stride)[0]
        return torch.cat([known.to(self.torch_dtype), z.to(dtype=self.torch_dtype, device=dev)]).unsqueeze(0)

    @torch.no_grad()
    def __call__(self, prompt, negative_prompt="", input_image=None, end_image=None, denoising_strength=1.0, seed=None, rand_device="cpu",
                 height=480, width=832, num_frames=81, cfg_scale=5.0, num_inference_steps=50, sigma_shift=5.0, tiled=True,
                 tile_size=(30, 52), tile_stride=(15, 26), return_condition=False, **other):
        used = [k for k in _UNSUPPORTED_INPUTS if other.get(k) is not None]
        if used:
            raise NotImplementedError(f"WanVideoPipeline.__call__: inputs outside the FantasyWorld path: {used}")
        if not return_condition:
            raise NotImplementedError("WanVideoPipeline.__call__ computes the conditioning only (return_condition=True); the sampler loop "
                                      "is FantasyWorld's own (inference_wan22.py generate_video_with_dual_models)")
        self.scheduler.set_timesteps(num_inference_steps, denoising_strength=denoising_strength, shift=sigma_shift)
        height, width, num_frames = self.check_resize_height_width(height, width, num_frames)
        shared = {"input_image": input_image, "end_image": end_image, "seed": seed, "rand_device": rand_device, "height": height,
                  "width": width, "num_frames": num_frames, "cfg_scale": cfg_scale, "sigma_shift": sigma_shift, "tiled": tiled,
                  "tile_size": tile_size, "tile_stride": tile_stride, "denoising_strength": denoising_strength}
        posi = {"prompt": prompt, "num_inference_steps": num_inference_steps}
        nega = {"negative_prompt": negative_prompt, "num_inference_steps": num_inference_steps}
        z_dim = self.vae.model.z_dim if self.vae is not None else 16
        f = self.vae.upsampling_factor if self.vae is not None else 8
        noise = self.generate_noise((1, z_dim, (num_frames - 1) // 4 + 1, height // f, width // f), seed=seed, rand_device=rand_device)
        shared["noise"] = shared["latents"] = noise
        posi["context"] = self.embed_prompt(prompt, positive=True)
        nega["context"] = self.embed_prompt(negative_prompt, positive=False)
        if input_image is not None and getattr(self.dit, "require_vae_embedding", True):
            shared["y"] = self.embed_image_vae(input_image, end_image, num_frames, height, width, tiled, tile_size, tile_stride)
        return shared, posi, nega
