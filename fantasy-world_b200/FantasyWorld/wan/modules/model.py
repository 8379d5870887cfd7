"""Mirror of the two helpers of FantasyWorld/wan/modules/model.py that the fusion path uses (the rest of that
file is an unused copy of the official Wan model and is out of scope).  No `diffusers` dependency."""
import torch

__all__ = ["sinusoidal_embedding_1d", "rope_params"]


def sinusoidal_embedding_1d(dim, position):
    """fp64 cos|sin embedding, NOT cast back (the VGGT caller does .float()).  ref: wan/modules/model.py:17-27."""
    assert dim % 2 == 0
    half = dim // 2
    pos = position.type(torch.float64)
    ang = torch.outer(pos, torch.pow(10000, -torch.arange(half).to(pos).div(half)))
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)


def rope_params(max_seq_len, dim, theta=10000):
    """complex128 [max_seq_len, dim/2].  ref: wan/modules/model.py:30-38."""
    assert dim % 2 == 0
    ang = torch.outer(torch.arange(max_seq_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(ang), ang)
