"""B200-native mirror of FantasyWorld/diffsynth_wan21/models/pose_adaptor_ac3d.py (reference).

CameraPoseEncoder turns the per-pixel Plücker embedding [b, 81, H, W, 6] into token-aligned camera features
[b, f*h*w, 2048].  It runs ONCE per sample, outside the 50-step loop: the 1x1 conv / GroupNorm /
temporal-pooling front end is a few hundred MFLOP and stays on torch; the token-sized tail (patch embedding and the
fc stack) runs on the fwb200 GEMM / LayerNorm kernels.  Same state_dict keys as the reference.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from fwb200 import engine as E
from fwb200 import ops


class CameraPoseEncoder(nn.Module):
    def __init__(self, context_dim: int = 2048, dim: int = 5120, patch_size: Tuple[int, int, int] = [1, 2, 2],
                 in_channels: int = 6, downscale_coef: int = 8, pose_inject_method='adaln', **kwargs):
        super().__init__()
        c0 = in_channels * downscale_coef ** 2
        self.pose_inject_method = pose_inject_method
        self.unshuffle = nn.PixelUnshuffle(downscale_coef)
        self.controlnet_encode_first = nn.Sequential(
            nn.Conv2d(c0, c0, kernel_size=1), nn.GroupNorm(2, c0), nn.Conv2d(c0, c0, kernel_size=1), nn.GroupNorm(2, c0),
            nn.ReLU())
        self.controlnet_encode_second = nn.Sequential(nn.Conv2d(c0, 2 * c0, kernel_size=1), nn.GroupNorm(2, 2 * c0), nn.ReLU())
        self.patch_embedding = nn.Conv3d(2 * c0, dim, kernel_size=patch_size, stride=patch_size)
        self.patch_size = tuple(patch_size)
        if pose_inject_method in ('adaln', 'latent_split'):
            self.fc = nn.Sequential(nn.Linear(dim, dim // 2), nn.LayerNorm(dim // 2), nn.GELU(),
                                    nn.Linear(dim // 2, context_dim), nn.LayerNorm(context_dim))

    @staticmethod
    def compress_time(x, num_frames):
        """Halve the frame axis by average pooling, keeping the first frame when the count is odd (81 -> 41 -> 21).
        ref: pose_adaptor_ac3d.py:58-73."""
        bf, c, h, w = x.shape
        b = bf // num_frames
        v = x.view(b, num_frames, c, h, w)
        if num_frames % 2 == 1:
            head, rest = v[:, :1], v[:, 1:]
            if rest.shape[1] > 0:
                rest = rest.reshape(b, rest.shape[1] // 2, 2, c, h, w).mean(dim=2)
            v = torch.cat([head, rest], dim=1)
        else:
            v = v.reshape(b, num_frames // 2, 2, c, h, w).mean(dim=2)
        return v.reshape(-1, c, h, w)

    def patchify(self, x: torch.Tensor):
        b, cin, F_, H_, W_ = x.shape
        pf, ph, pw = self.patch_size
        f, h, w = F_ // pf, H_ // ph, W_ // pw
        cols = E.as_bf16(x).view(b, cin, f, pf, h, ph, w, pw).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(b * f * h * w, -1)
        tok = E.lin(cols.contiguous(), self.patch_embedding, round_flags=ops.ROUND_AFTER_BIAS)
        return tok.view(b, f * h * w, -1), (f, h, w)

    def forward(self, x):
        b, nf = x.shape[0], x.shape[1]
        x = x.permute(0, 1, 4, 2, 3).reshape(b * nf, x.shape[4], x.shape[2], x.shape[3])
        x = self.controlnet_encode_first(self.unshuffle(x))
        x = self.compress_time(x, nf)
        nf = x.shape[0] // b
        x = self.controlnet_encode_second(x)
        x = self.compress_time(x, nf)
        nf = x.shape[0] // b
        x = x.view(b, nf, *x.shape[1:]).permute(0, 2, 1, 3, 4)
        x, _ = self.patchify(x)
        if self.pose_inject_method in ('adaln', 'latent_split'):
            l0, n1, _, l3, n4 = self.fc
            shp = x.shape
            h = E.lin(x.reshape(-1, shp[-1]), l0, round_flags=ops.ROUND_AFTER_BIAS)
            h = ops.ln_modulate(h, eps=n1.eps, w=E.f32(n1, "w", n1.weight), b=E.f32(n1, "b", n1.bias))
            h = F.gelu(h)
            h = E.lin(h, l3, round_flags=ops.ROUND_AFTER_BIAS)
            h = ops.ln_modulate(h, eps=n4.eps, w=E.f32(n4, "w", n4.weight), b=E.f32(n4, "b", n4.bias))
            x = h.view(*shp[:-1], -1)
        return x
