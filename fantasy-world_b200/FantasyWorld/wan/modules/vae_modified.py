"""Mirror of the parts of FantasyWorld/wan/modules/vae_modified.py that the geometry heads use (reference):
the causal 4x temporal up-sampler `WanVAE_(location="DPT")` (21 latent frames -> 81 frames) and
`ChannelExpandAndReshape`.  Same state_dict keys.

The reference decodes frame by frame with a per-convolution cache (vae_modified.py:454-476).  That streaming scheme is
algebraically a causal convolution over the whole clip in which the first frame by-passes the temporal up-sampling
('Rep' sentinel, :92-95) — so this mirror evaluates each stage once over the full sequence (no Python loop over
frames); frame selection / interleaving indices are identical.  Runs once per video, torch/cuDNN ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["CausalConv3d", "RMS_norm", "Resample", "ResidualBlock_Half", "Decoder3d_Simple", "WanVAE_",
           "ChannelExpandAndReshape"]


class CausalConv3d(nn.Conv3d):
    """Conv3d whose temporal padding is all on the past side (2*p frames in front, none behind).
    ref: vae_modified.py:17-36."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        pt, ph, pw = self.padding
        self._padding = (pw, pw, ph, ph, 2 * pt, 0)
        self.padding = (0, 0, 0)

    def forward(self, x, cache_x=None):
        pad = list(self._padding)
        if cache_x is not None and pad[4] > 0:
            x = torch.cat([cache_x.to(x.device), x], dim=2)
            pad[4] -= cache_x.shape[2]
        return super().forward(F.pad(x, pad))


class RMS_norm(nn.Module):
    """L2-normalise over channels, scale by sqrt(dim) * gamma.  ref: vae_modified.py:39-55."""

    def __init__(self, dim, channel_first=True, images=True, bias=False):
        super().__init__()
        tail = (1, 1) if images else (1, 1, 1)
        shape = (dim, *tail) if channel_first else (dim,)
        self.channel_first = channel_first
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(shape))
        self.bias = nn.Parameter(torch.zeros(shape)) if bias else 0.

    def forward(self, x):
        return F.normalize(x, dim=1 if self.channel_first else -1) * self.scale * self.gamma + self.bias


class Resample(nn.Module):
    """'upsample3d': causal temporal 2x up-sampling of every frame except the first.  ref: vae_modified.py:66-130."""

    def __init__(self, dim, mode):
        super().__init__()
        if mode != 'upsample3d':
            raise NotImplementedError("only 'upsample3d' is used by the geometry heads")
        self.dim, self.mode = dim, mode
        self.resample = nn.Identity()
        self.time_conv = CausalConv3d(dim, dim * 2, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x):
        """x [b, c, t, h, w] (whole clip) -> [b, c, 1 + 2 (t-1), h, w]."""
        b, c, t, h, w = x.shape
        if t == 1:
            return x
        y = self.time_conv(x[:, :, 1:])                       # zero history: frame 0 is not seen (the 'Rep' rule)
        y = y.view(b, 2, c, t - 1, h, w)
        y = torch.stack((y[:, 0], y[:, 1]), dim=3).reshape(b, c, 2 * (t - 1), h, w)   # (even, odd) interleave
        return torch.cat([x[:, :, :1], y], dim=2)


class ResidualBlock_Half(nn.Module):
    """x + causal_conv3(silu(rms_norm(x))).  ref: vae_modified.py:193-225."""

    def __init__(self, in_dim, out_dim, dropout=0.0):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.residual = nn.Sequential(RMS_norm(in_dim, images=False), nn.SiLU(), CausalConv3d(in_dim, out_dim, 3, padding=1))
        self.shortcut = CausalConv3d(in_dim, out_dim, 1) if in_dim != out_dim else nn.Identity()

    def forward(self, x):
        return self.residual(x) + self.shortcut(x)


class Decoder3d_Simple(nn.Module):
    """ref: vae_modified.py:373-408."""

    def __init__(self, dim=128, z_dim=4, dim_mult=[1, 2, 4, 4], attn_scales=[], temperal_upsample=[False, True, True],
                 dropout=0.0, residual=False):
        super().__init__()
        self.dim, self.z_dim = dim, z_dim
        layers = []
        for _ in range(2):
            layers.append(Resample(z_dim, mode="upsample3d"))
            if residual:
                layers.append(ResidualBlock_Half(z_dim, z_dim, dropout))
        self.upsamples = nn.Sequential(*layers)

    def forward(self, x):
        return self.upsamples(x)


class WanVAE_(nn.Module):
    """Temporal up-sampler wrapper.  ref: vae_modified.py:419-483 (only .decode is used)."""

    def __init__(self, dim=128, z_dim=4, dim_mult=[1, 2, 1, 1], attn_scales=[], temperal_downsample=[True, True, False],
                 dropout=0.0, location=None):
        super().__init__()
        self.temperal_upsample = temperal_downsample[::-1]
        self.conv2 = None
        if location in ("camera", "Rep"):
            self.decoder = Decoder3d_Simple(dim, z_dim, dim_mult, attn_scales, self.temperal_upsample, dropout, residual=False)
        elif location == "DPT":
            self.conv2 = CausalConv3d(z_dim, z_dim, 1)
            self.decoder = Decoder3d_Simple(dim, z_dim, dim_mult, attn_scales, self.temperal_upsample, dropout, residual=True)
        else:
            raise NotImplementedError(location)

    def decode(self, z):
        x = self.conv2(z) if self.conv2 is not None else z
        return self.decoder(x)


class ChannelExpandAndReshape(nn.Module):
    """[B, N, C] -> [B, 4N, C] through a 1x1 Conv1d C -> 4C whose output channels are re-read as (C, 4N).
    ref: vae_modified.py:555-572 (the reshape deliberately mixes channel and time indices; kept index-exact)."""

    def __init__(self, input_channels):
        super().__init__()
        self.expand_channels = nn.Conv1d(in_channels=input_channels, out_channels=input_channels * 4, kernel_size=1)

    def forward(self, x):
        B, N, C = x.shape
        y = self.expand_channels(x.transpose(1, 2))   # [B, 4C, N]
        return y.reshape(B, C, N * 4).transpose(1, 2)
