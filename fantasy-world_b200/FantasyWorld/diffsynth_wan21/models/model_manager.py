"""Minimal stand-in for FantasyWorld/diffsynth_wan21/models/model_manager.py.

The reference ModelManager (hash-of-state-dict-keys model detection over a registry of ~100 architectures, safetensors /
ModelScope download, model_manager.py + configs/model_config.py) is I/O and out of scope.  This version keeps the
calls FantasyWorldFusionModel and WanVideoPipeline make — `load_models(paths)`, `fetch_model(name, require_model_path=...)` — for
the four checkpoints inference_wan21.py:183-188 lists: the Wan DiT safetensors shards (keys already in DiffSynth naming),
`Wan2.1_VAE.pth`, `models_clip_open-clip-xlm-roberta-large-vit-huge-14.pth` and `models_t5_umt5-xxl-enc-bf16.pth`.  The three
`.pth` files are recognised by a characteristic key instead of the reference's key-set hash and go through the same
`state_dict_converter().from_civitai` as in the reference (configs/model_config.py:26-29).  With no path it builds a random-init
WanModel of the requested configuration (what the benchmarks use — there are no checkpoints here).
"""
from __future__ import annotations

import torch

WAN21_I2V_14B = dict(dim=5120, in_dim=36, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                     patch_size=(1, 2, 2), num_heads=40, num_layers=40, has_image_input=True)


def detect_pth(keys) -> str | None:
    """Which Wan side model a checkpoint's key set belongs to (None: not one of ours)."""
    keys = set(keys)
    if "token_embedding.weight" in keys and "blocks.0.attn.q.weight" in keys:
        return "wan_video_text_encoder"
    if "visual.patch_embedding.weight" in keys or "model.visual.patch_embedding.weight" in keys:
        return "wan_video_image_encoder"
    if "model_state" in keys or "encoder.conv1.weight" in keys or "decoder.conv1.weight" in keys:
        return "wan_video_vae"
    return None


def _build_side_model(name):
    if name == "wan_video_text_encoder":
        from .wan_video_text_encoder import WanTextEncoder
        return WanTextEncoder
    if name == "wan_video_image_encoder":
        from .wan_video_image_encoder import WanImageEncoder
        return WanImageEncoder
    from .wan_video_vae import WanVideoVAE
    return WanVideoVAE


class ModelManager:
    def __init__(self, torch_dtype=torch.bfloat16, device="cpu", dit_config: dict | None = None, side_configs: dict | None = None):
        """`side_configs` (extension, tests): constructor overrides per side model name, e.g. {"wan_video_text_encoder": {...}}."""
        self.torch_dtype, self.device = torch_dtype, device
        self.dit_config = dict(dit_config or WAN21_I2V_14B)
        self.side_configs = dict(side_configs or {})
        self.models = {}
        self.model_paths = {}

    def load_state_dict_model(self, state_dict, path=None, torch_dtype=None, device=None):
        """One `.pth` state dict -> the matching side model (strict load after the reference's key conversion)."""
        name = detect_pth(state_dict.keys())
        if name is None:
            return None
        cls = _build_side_model(name)
        cfg = self.side_configs.get(name, {})
        if name == "wan_video_vae":
            model = cls(**cfg)                   # small; holds plain (non-parameter) mean / std tensors
        elif name == "wan_video_image_encoder":
            model = cls(device="meta", **cfg)
        else:
            with torch.device("meta"):
                model = cls(**cfg)
        sd = cls.state_dict_converter().from_civitai(state_dict)
        model.load_state_dict(sd, strict=True, assign=True)
        model = model.to(device=device or self.device, dtype=torch_dtype or self.torch_dtype).eval()
        self.models[name], self.model_paths[name] = model, path
        return name

    def load_models(self, file_paths=None, torch_dtype=None, device=None):
        from .wan_video_dit import WanModel
        dtype = torch_dtype or self.torch_dtype
        dit = WanModel(**self.dit_config)   # honours an enclosing torch.device(...) context (e.g. "meta")
        dit = dit.to(dtype)
        if file_paths:
            from safetensors.torch import load_file
            paths = [p for group in file_paths for p in (group if isinstance(group, (list, tuple)) else [group])]
            sd = {}
            for p in paths:
                if str(p).endswith(".safetensors"):
                    sd.update(load_file(str(p), device="cpu"))
                elif str(p).endswith((".pth", ".pt", ".ckpt")):
                    self.load_state_dict_model(torch.load(str(p), map_location="cpu", weights_only=True), path=str(p),
                                               torch_dtype=dtype, device=device)
            dit_keys = set(dit.state_dict().keys())
            sd = {k: v for k, v in sd.items() if k in dit_keys}
            if sd:
                missing, unexpected = dit.load_state_dict(sd, strict=False)
                assert not unexpected
        self.models["wan_video_dit"] = dit

    def fetch_model(self, name, file_path=None, require_model_path=False, **kw):
        model = self.models.get(name)
        if model is None:
            return None
        return (model, self.model_paths.get(name)) if require_model_path else model
