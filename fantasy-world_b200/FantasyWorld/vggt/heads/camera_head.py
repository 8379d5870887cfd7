"""Mirror of FantasyWorld/vggt/heads/camera_head.py (reference): iterative camera-pose regression from the camera
tokens of the last aggregator layer (last denoising step only).  Same state_dict keys.

81 tokens x 2048 channels: the trunk Blocks run on the fwb200 kernels (GEMM / attention D=128 / LayerNorm) through the
generic Block path; the remaining O(81 x 2048) element-wise work is plain torch.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ...wan.modules.vae_modified import ChannelExpandAndReshape
from ..heads.head_act import activate_pose
from ..layers import Mlp
from ..layers.block import Block


def modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    return x * (1 + scale) + shift


class CameraHead(nn.Module):
    def __init__(self, dim_in: int = 2048, trunk_depth: int = 4, pose_encoding_type: str = "absT_quaR_FoV",
                 num_heads: int = 16, mlp_ratio: int = 4, init_values: float = 0.01, trans_act: str = "linear",
                 quat_act: str = "linear", fl_act: str = "relu"):
        super().__init__()
        if pose_encoding_type != "absT_quaR_FoV":
            raise ValueError(f"Unsupported camera encoding type: {pose_encoding_type}")
        self.target_dim = 9
        self.trans_act, self.quat_act, self.fl_act, self.trunk_depth = trans_act, quat_act, fl_act, trunk_depth
        self.trunk = nn.Sequential(*[Block(dim=dim_in, num_heads=num_heads, mlp_ratio=mlp_ratio, init_values=init_values)
                                     for _ in range(trunk_depth)])
        self.token_norm = nn.LayerNorm(dim_in)
        self.trunk_norm = nn.LayerNorm(dim_in)
        self.empty_pose_tokens = nn.Parameter(torch.zeros(1, 1, self.target_dim))
        self.embed_pose = nn.Linear(self.target_dim, dim_in)
        self.poseLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(dim_in, 3 * dim_in, bias=True))
        self.camera_time_upsample = ChannelExpandAndReshape(input_channels=dim_in)
        self.adaln_norm = nn.LayerNorm(dim_in, elementwise_affine=False, eps=1e-6)
        self.pose_branch = Mlp(in_features=dim_in, hidden_features=dim_in // 2, out_features=self.target_dim, drop=0)

    def forward(self, aggregated_tokens_list: list, num_iterations: int = 4) -> list:
        """Token 0 of every latent frame of the last layer; frames 1.. are expanded 4x in time (un-normalised, as in the
        reference), frame 0 is LayerNorm-ed -> 1 + 4 (S-1) pose tokens.  ref: camera_head.py:76-97."""
        cam = aggregated_tokens_list[-1][:, :, 0]
        rest = self.camera_time_upsample(cam[:, 1:])
        first = self.token_norm(cam)[:, 0:1]
        return self.trunk_fn(torch.cat([first, rest.to(first.dtype)], dim=1), num_iterations)

    def trunk_fn(self, pose_tokens: torch.Tensor, num_iterations: int) -> list:
        """ref: camera_head.py:99-145."""
        B, S, _ = pose_tokens.shape
        enc = None
        outs = []
        for _ in range(num_iterations):
            src = self.empty_pose_tokens.expand(B, S, -1) if enc is None else enc.detach()
            shift, scale, gate = self.poseLN_modulation(self.embed_pose(src)).chunk(3, dim=-1)
            h = gate * modulate(self.adaln_norm(pose_tokens), shift, scale) + pose_tokens
            h = self.trunk(h)
            delta = self.pose_branch(self.trunk_norm(h))
            enc = delta if enc is None else enc + delta
            outs.append(activate_pose(enc, trans_act=self.trans_act, quat_act=self.quat_act, fl_act=self.fl_act))
        return outs
