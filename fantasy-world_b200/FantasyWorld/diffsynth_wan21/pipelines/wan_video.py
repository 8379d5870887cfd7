"""Stand-in for FantasyWorld/diffsynth_wan21/pipelines/wan_video.py: the attribute and method surface FantasyWorldFusionModel and
inference_wan21.py touch (`dit`, `scheduler`, `device`, `torch_dtype`, `generate_noise`, `prepare_extra_input`,
`load_models_to_device`, `vae`, `prompter`, `encode_prompt`, `encode_image`, `preprocess_image(s)`).

The sub-models are the mirrors next to this file: the VAE (`pipe.vae.decode(..., tiled=True)`), the umT5 text encoder
and the CLIP ViT-H image encoder (N3).  Attach them by loading checkpoints through the model manager (`fetch_models`) or random-init
with `enable_vae()` / `enable_text_encoder()` / `enable_image_encoder()`; `encode_prompt` / `encode_image` raise with a clear message
while their model (or the tokenizer) is missing — never a silent substitute.  What stays out: CPU offload / VRAM
management, TeaCache, the VACE / motion-controller branches, the pipeline's own `__call__` (FantasyWorld drives the loop itself:
fusion/model_wan21.py:196-330).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from ..prompters import WanPrompter
from ..schedulers.flow_match import FlowMatchScheduler


class WanVideoPipeline(nn.Module):
    def __init__(self, device="cuda", torch_dtype=torch.bfloat16, tokenizer_path=None):
        super().__init__()
        self.device, self.torch_dtype = device, torch_dtype
        self.scheduler = FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
        self.prompter = WanPrompter(tokenizer_path=tokenizer_path)
        self.text_encoder = None
        self.image_encoder = None
        self.dit = None
        self.vae = None
        self.model_names = ['text_encoder', 'dit', 'vae', 'image_encoder']
        self.cpu_offload = False

    def fetch_models(self, model_manager):
        """wan_video.py:168-186: pick the loaded sub-models; the tokenizer directory sits next to the T5 checkpoint."""
        found = model_manager.fetch_model("wan_video_text_encoder", require_model_path=True)
        if found is not None:
            self.text_encoder, path = found
            self.prompter.fetch_models(self.text_encoder)
            tok = os.path.join(os.path.dirname(path), "google/umt5-xxl") if path else None
            if tok and os.path.isdir(tok):
                self.prompter.fetch_tokenizer(tok)
        self.dit = model_manager.fetch_model("wan_video_dit")
        self.vae = model_manager.fetch_model("wan_video_vae")
        self.image_encoder = model_manager.fetch_model("wan_video_image_encoder")

    @staticmethod
    def from_model_manager(model_manager, torch_dtype=None, device=None):
        pipe = WanVideoPipeline(device=device or model_manager.device, torch_dtype=torch_dtype or model_manager.torch_dtype)
        pipe.fetch_models(model_manager)
        return pipe

    # -- random-init attachments (benchmarks / tests: there are no checkpoints in this environment) ----------------------------
    def _attach(self, model, state_dict, device, dtype):
        if state_dict is not None:
            model.load_state_dict(state_dict, strict=True)
        return model.to(device=device or self.device, dtype=dtype or self.torch_dtype).eval()

    def enable_vae(self, z_dim: int = 16, state_dict=None, device=None, dtype=None):
        """Attach the Wan VAE mirror (random-init unless a state_dict with the reference's `model.*` keys is given)."""
        from ..models.wan_video_vae import WanVideoVAE
        self.vae = self._attach(WanVideoVAE(z_dim=z_dim), state_dict, device, dtype)
        return self.vae

    def _materialize(self, model, state_dict, device, dtype, init):
        """A model built under torch.device("meta") gets its storage directly on the target device (umT5-XXL is 5.7 B parameters:
        never on the host in fp32): from `state_dict` when given, else random-initialised there by `init(model)`."""
        from fwb_synth import materialize
        device, dtype = device or self.device, dtype or self.torch_dtype
        if state_dict is not None:
            model.load_state_dict(state_dict, strict=True, assign=True)
            return model.to(device=device, dtype=dtype).eval()
        materialize(model, device, dtype)
        with torch.no_grad():
            init(model)
        return model.eval()

    def enable_text_encoder(self, state_dict=None, device=None, dtype=None, **config):
        """Attach the umT5 encoder mirror: `state_dict` in the released checkpoint's layout, else the encoder's own random init."""
        from ..models.wan_video_text_encoder import WanTextEncoder, init_weights

        def init(m):
            m.apply(init_weights)
            nn.init.normal_(m.token_embedding.weight)

        with torch.device("meta"):
            model = WanTextEncoder(**config)
        self.text_encoder = self._materialize(model, state_dict, device, dtype, init)
        self.prompter.fetch_models(self.text_encoder)
        return self.text_encoder

    def enable_image_encoder(self, state_dict=None, device=None, dtype=None, seed=0, **config):
        """Attach the CLIP image tower mirror: `state_dict` with the `model.*` keys (after the converter), else per-key synthetic
        weights (fwb_synth) generated on the device."""
        from fwb_synth import synth_init
        from ..models.wan_video_image_encoder import WanImageEncoder
        model = WanImageEncoder(device="meta", **config)
        self.image_encoder = self._materialize(model, state_dict, device, dtype, lambda m: synth_init(m, seed))
        return self.image_encoder

    # -- sampler helpers -----------------------------------------------------------------------------------------------------
    def generate_noise(self, shape, seed=None, device="cpu", dtype=torch.float16):
        gen = None if seed is None else torch.Generator(device).manual_seed(seed)
        return torch.randn(shape, generator=gen, device=device, dtype=dtype)

    def prepare_extra_input(self, latents=None):
        return {}

    def load_models_to_device(self, loadmodel_names=[]):
        return None  # no CPU offload: 37 GB of weights stay resident in 180 GB of HBM

    def preprocess_image(self, image):
        """PIL image -> [1, 3, H, W] fp32 in [-1, 1] (pipelines/base.py:30-32)."""
        return torch.from_numpy(np.array(image, dtype=np.float32) * (2 / 255) - 1).permute(2, 0, 1).unsqueeze(0)

    def preprocess_images(self, images):
        return [self.preprocess_image(im) for im in images]

    # -- conditioning (once per sample) ----------------------------------------------------------------------------------------
    def encode_prompt(self, prompt, positive=True):
        """wan_video.py:213-216."""
        if self.text_encoder is None or self.prompter.tokenizer is None:
            raise RuntimeError("encode_prompt: no T5 text encoder / tokenizer attached (fetch_models, enable_text_encoder, "
                               "prompter.fetch_tokenizer)")
        return {"context": self.prompter.encode_prompt(prompt, positive=positive, device=self.device)}

    @torch.no_grad()
    def encode_image(self, image, end_image, num_frames, height, width, tiled=False, tile_size=(34, 34), tile_stride=(18, 16)):
        """wan_video.py:218-276: CLIP tokens of the first (and last) frame; `y` = 4 first/last-frame mask channels, folded 4 pixel
        frames per latent frame, stacked on the VAE encoding of the clip with the unknown frames zeroed."""
        if self.image_encoder is None or self.vae is None:
            raise RuntimeError("encode_image: no CLIP image encoder / VAE attached (fetch_models, enable_image_encoder, enable_vae)")
        dev = self.device
        first = self.preprocess_image(image.resize((width, height))).to(dev)
        clip_context = self.image_encoder.encode_image([first])
        known = torch.zeros(1, num_frames, height // 8, width // 8, device=dev)
        known[:, 0] = 1
        frames = [first.transpose(0, 1)]
        if end_image is not None:
            last = self.preprocess_image(end_image.resize((width, height))).to(dev)
            frames += [torch.zeros(3, num_frames - 2, height, width, device=dev), last.transpose(0, 1)]
            if getattr(self.dit, "has_image_pos_emb", False):
                clip_context = torch.cat([clip_context, self.image_encoder.encode_image([last])], dim=1)
            known[:, -1] = 1
        else:
            frames.append(torch.zeros(3, num_frames - 1, height, width, device=dev))
        clip = torch.cat(frames, dim=1)
        # the first frame's flag is repeated 4x so that (1 + (T-1)) pixel frames fold into (T+3)/4 groups of 4
        known = torch.cat([known[:, :1].repeat_interleave(4, dim=1), known[:, 1:]], dim=1)
        known = known.view(1, known.shape[1] // 4, 4, height // 8, width // 8).transpose(1, 2)[0]
        y = self.vae.encode([clip.to(dtype=self.torch_dtype, device=dev)], device=dev, tiled=tiled, tile_size=tile_size,
                            tile_stride=tile_stride)[0]
        y = torch.cat([known.to(dtype=self.torch_dtype), y.to(dtype=self.torch_dtype, device=dev)]).unsqueeze(0)
        return {"clip_feature": clip_context.to(dtype=self.torch_dtype, device=dev), "y": y}
