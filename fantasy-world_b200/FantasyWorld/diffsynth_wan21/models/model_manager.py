"""Minimal stand-in for FantasyWorld/diffsynth_wan21/models/model_manager.py.

The reference ModelManager (hash-of-state-dict-keys model detection, safetensors / ModelScope download) is I/O and out
of scope.  This version keeps the two calls FantasyWorldFusionModel makes: `load_models(paths)` and
`fetch_model(name)`.  It understands Wan DiT safetensors shards (keys already in DiffSynth naming); with no path it
builds a random-init WanModel of the requested configuration (what the benchmarks use — there are no checkpoints here).
"""
from __future__ import annotations

import torch

WAN21_I2V_14B = dict(dim=5120, in_dim=36, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                     patch_size=(1, 2, 2), num_heads=40, num_layers=40, has_image_input=True)


class ModelManager:
    def __init__(self, torch_dtype=torch.bfloat16, device="cpu", dit_config: dict | None = None):
        self.torch_dtype, self.device = torch_dtype, device
        self.dit_config = dict(dit_config or WAN21_I2V_14B)
        self.models = {}

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
            dit_keys = set(dit.state_dict().keys())
            sd = {k: v for k, v in sd.items() if k in dit_keys}
            if sd:
                missing, unexpected = dit.load_state_dict(sd, strict=False)
                assert not unexpected
        self.models["wan_video_dit"] = dit

    def fetch_model(self, name, **kw):
        return self.models.get(name)
