"""B200-native mirror of FantasyWorld/diffsynth_wan22/models/wan_video_dit.py (reference).  The Wan2.2 DiT block is the
Wan2.1 block (same parameters, same math; the reference copy only inlines the cross-attention processor and adds a few
constructor flags), so the fwb200-backed classes are shared with the Wan2.1 mirror; this module adds the Wan2.2 WanModel
constructor surface (`require_vae_embedding`, `require_clip_embedding`, `seperated_timestep`,
`fuse_vae_embedding_in_latents`) and its `patchify(x, control_camera_latents_input)` convention (5-D output).
"""
from typing import Optional, Tuple

import torch

from ...diffsynth_wan21.models import wan_video_dit as _w21
from ...diffsynth_wan21.models.wan_video_dit import (AttentionModule, CrossAttention, DiTBlock, GateModule, Head, MLP,  # noqa: F401
                                                     RMSNorm, SelfAttention, build_freqs_3d_with_extra_cis, flash_attention,
                                                     modulate, precompute_freqs_cis, precompute_freqs_cis_3d, rope_apply,
                                                     sinusoidal_embedding_1d)
from .wan_video_camera_controller import SimpleAdapter  # noqa: F401

# Wan2.2-Fun-A14B-Control-Camera (ref: diffsynth_wan22/models/wan_video_dit.py:841-859)
WAN22_FUN_A14B_CONTROL_CAMERA = dict(has_image_input=False, patch_size=(1, 2, 2), in_dim=36, dim=5120, ffn_dim=13824, freq_dim=256,
                                     text_dim=4096, out_dim=16, num_heads=40, num_layers=40, eps=1e-6, has_ref_conv=False,
                                     add_control_adapter=True, in_dim_control_adapter=24, require_clip_embedding=False)


class WanModel(_w21.WanModel):
    def __init__(self, dim: int, in_dim: int, ffn_dim: int, out_dim: int, text_dim: int, freq_dim: int, eps: float,
                 patch_size: Tuple[int, int, int], num_heads: int, num_layers: int, has_image_input: bool,
                 has_image_pos_emb: bool = False, has_ref_conv: bool = False, add_control_adapter: bool = False,
                 in_dim_control_adapter: int = 24, seperated_timestep: bool = False, require_vae_embedding: bool = True,
                 require_clip_embedding: bool = True, fuse_vae_embedding_in_latents: bool = False):
        super().__init__(dim=dim, in_dim=in_dim, ffn_dim=ffn_dim, out_dim=out_dim, text_dim=text_dim, freq_dim=freq_dim, eps=eps,
                         patch_size=patch_size, num_heads=num_heads, num_layers=num_layers, has_image_input=has_image_input,
                         has_image_pos_emb=has_image_pos_emb, has_ref_conv=has_ref_conv, add_control_adapter=add_control_adapter,
                         in_dim_control_adapter=in_dim_control_adapter)
        self.seperated_timestep = seperated_timestep
        self.require_vae_embedding = require_vae_embedding
        self.require_clip_embedding = require_clip_embedding
        self.fuse_vae_embedding_in_latents = fuse_vae_embedding_in_latents

    def patchify(self, x: torch.Tensor, control_camera_latents_input: Optional[torch.Tensor] = None):
        """Returns [b, dim, f, h, w] (the Wan2.2 reference leaves the token flattening to the caller).
        ref: diffsynth_wan22/models/wan_video_dit.py:390-396."""
        tok, (f, h, w) = super().patchify(x, control_camera_latents_input)
        return tok.view(tok.shape[0], f, h, w, -1).permute(0, 4, 1, 2, 3)
