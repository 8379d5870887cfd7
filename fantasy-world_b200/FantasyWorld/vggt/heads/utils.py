"""Mirror of FantasyWorld/vggt/heads/utils.py: UV grid and sin/cos positional embedding used by the DPT head."""
import torch


def make_sincos_pos_embed(embed_dim: int, pos: torch.Tensor, omega_0: float = 100) -> torch.Tensor:
    """[M] positions -> [M, embed_dim] = sin | cos of pos * omega_0^(-2i/embed_dim), evaluated in fp64.  ref: utils.py:37-62."""
    assert embed_dim % 2 == 0
    half = embed_dim // 2
    expo = torch.arange(half, dtype=torch.double, device=pos.device) / (embed_dim / 2.0)
    ang = pos.reshape(-1)[:, None] * (1.0 / omega_0 ** expo)[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=1).float()


def position_grid_to_embed(pos_grid: torch.Tensor, embed_dim: int, omega_0: float = 100) -> torch.Tensor:
    """[H, W, 2] -> [H, W, embed_dim]: first half from x, second half from y.  ref: utils.py:11-34."""
    H, W, two = pos_grid.shape
    assert two == 2
    flat = pos_grid.reshape(-1, 2)
    emb = torch.cat([make_sincos_pos_embed(embed_dim // 2, flat[:, 0], omega_0),
                     make_sincos_pos_embed(embed_dim // 2, flat[:, 1], omega_0)], dim=-1)
    return emb.view(H, W, embed_dim)


def create_uv_grid(width: int, height: int, aspect_ratio: float = None, dtype: torch.dtype = None,
                   device: torch.device = None) -> torch.Tensor:
    """[height, width, 2] grid of (u, v) in a plane normalised by its diagonal.  ref: utils.py:68-109."""
    if aspect_ratio is None:
        aspect_ratio = float(width) / float(height)
    diag = (aspect_ratio ** 2 + 1.0) ** 0.5
    sx, sy = aspect_ratio / diag, 1.0 / diag
    xs = torch.linspace(-sx * (width - 1) / width, sx * (width - 1) / width, steps=width, dtype=dtype, device=device)
    ys = torch.linspace(-sy * (height - 1) / height, sy * (height - 1) / height, steps=height, dtype=dtype, device=device)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")
    return torch.stack((uu, vv), dim=-1)
