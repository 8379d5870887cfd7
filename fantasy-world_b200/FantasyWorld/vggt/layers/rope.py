"""Mirror of FantasyWorld/vggt/layers/rope.py (PositionGetter, RotaryPositionEmbedding2D).

In the fused attention path the rotation is applied by fwb_ln64_rope2d from per-token tables built by
fwb200.engine.rope2d_expanded with the reference's arithmetic (fp32 angles, integer F.embedding-style gather), and the
`int(positions.max())` host sync the reference pays on every call (rope.py:177; 96 per forward) is paid once per
resolution.  The classes below keep the reference API for external callers.
"""
from typing import Dict, Tuple

import torch
import torch.nn as nn


class PositionGetter:
    """(y, x) integer coordinates of a height x width patch grid, cached.  ref: rope.py:24-59."""

    def __init__(self):
        self.position_cache: Dict[Tuple[int, int], torch.Tensor] = {}

    def __call__(self, batch_size: int, height: int, width: int, device: torch.device) -> torch.Tensor:
        key = (height, width)
        if key not in self.position_cache:
            ys = torch.arange(height, device=device).repeat_interleave(width)
            xs = torch.arange(width, device=device).repeat(height)
            self.position_cache[key] = torch.stack([ys, xs], dim=-1)
        return self.position_cache[key].to(device).view(1, height * width, 2).expand(batch_size, -1, -1).clone()


class RotaryPositionEmbedding2D(nn.Module):
    """2-D rotate-half RoPE: the first half of the features rotates with y, the second with x.  ref: rope.py:62-188."""

    def __init__(self, frequency: float = 100.0, scaling_factor: float = 1.0):
        super().__init__()
        self.base_frequency = frequency
        self.scaling_factor = scaling_factor

    def forward(self, tokens: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """tokens [B, H, N, D] (any float dtype), positions int [B, N, 2].  Standalone torch form (index math only)."""
        assert tokens.size(-1) % 2 == 0 and positions.ndim == 3 and positions.shape[-1] == 2
        half = tokens.size(-1) // 2
        expo = torch.arange(0, half, 2, device=tokens.device).float() / half
        inv = 1.0 / (self.base_frequency ** expo)
        n_pos = int(positions.max()) + 1
        ang = torch.arange(n_pos, device=tokens.device, dtype=inv.dtype)[:, None] * inv[None, :]
        ang = torch.cat((ang.to(tokens.dtype), ang.to(tokens.dtype)), dim=-1)
        cos_t, sin_t = ang.cos(), ang.sin()

        def rot(feat, p):
            c, s = cos_t[p][:, None], sin_t[p][:, None]
            q = feat.shape[-1] // 2
            return feat * c + torch.cat((-feat[..., q:], feat[..., :q]), dim=-1) * s

        return torch.cat((rot(tokens[..., :half], positions[..., 0]), rot(tokens[..., half:], positions[..., 1])), dim=-1)
