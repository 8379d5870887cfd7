"""B200-native mirror of FantasyWorld/vggt/layers/block.py (reference): AdaLN-modulated ViT block with the split
forward the IRG block needs, and CamTokenProjector.  Same state_dict keys.

Fused kernel sequence (rows = B*N tokens, C = 1024; the residual stream is fp32 once modulated — SURVEY Appendix A.3):
  ln_modulate(norm1, *(1+e1)+e0) -> qkv GEMM -> ln64_rope2d -> attention -> proj GEMM (+LayerScale, +resid)
  ln_modulate(norm2) -> fc1 GEMM (+GELU erf) -> fc2 GEMM (*(1+e4)+e3, *gamma*e5, +resid)      ref: block.py:73-116
"""
from typing import Callable

import torch
from torch import Tensor, nn

from fwb200 import engine as E

from .attention import Attention
from .layer_scale import LayerScale
from .mlp import Mlp


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = True, proj_bias: bool = True,
                 ffn_bias: bool = True, drop: float = 0.0, attn_drop: float = 0.0, init_values=None, drop_path: float = 0.0,
                 act_layer: Callable[..., nn.Module] = nn.GELU, norm_layer: Callable[..., nn.Module] = nn.LayerNorm,
                 attn_class: Callable[..., nn.Module] = Attention, ffn_layer: Callable[..., nn.Module] = Mlp,
                 qk_norm: bool = False, fused_attn: bool = True, rope=None) -> None:
        super().__init__()
        assert drop_path == 0.0 and drop == 0.0, "inference-only"
        self.norm1 = norm_layer(dim)
        self.attn = attn_class(dim, num_heads=num_heads, qkv_bias=qkv_bias, proj_bias=proj_bias, attn_drop=attn_drop,
                               proj_drop=drop, qk_norm=qk_norm, fused_attn=fused_attn, rope=rope)
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)
        self.ls1 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = ffn_layer(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop, bias=ffn_bias)
        self.ls2 = LayerScale(dim, init_values=init_values) if init_values else nn.Identity()
        self.drop_path2 = nn.Identity()
        self.sample_drop_ratio = drop_path

    def forward(self, x: Tensor, pos=None, e0=None, return_partial: bool = False, run_remaining: bool = False,
                modifiers: tuple | None = None):
        B, N, C = x.shape
        xs = x.reshape(B * N, C)
        if xs.dtype not in (torch.float32, torch.bfloat16):
            xs = xs.float()
        if run_remaining:
            assert modifiers is not None, "run_remaining need modifiers"
            return E.vggt_ffn_part(self, xs, modifiers).view(B, N, C)
        assert isinstance(self.ls1, LayerScale), "fused path expects LayerScale (init_values set), as the aggregator builds"
        mods = E.vggt_mod_vectors(self, e0) if e0 is not None else None
        at = self.attn
        if at.rope is not None and pos is not None:
            assert at.head_dim == 64 and isinstance(at.q_norm, nn.LayerNorm), "fused qk-norm/RoPE kernel: head_dim 64"
            tables = E.rope2d_expanded(pos, at.rope.base_frequency)
        else:
            assert isinstance(at.q_norm, nn.Identity), "qk-norm without RoPE is not a reference configuration"
            tables = None
        xs = E.vggt_attn_part(self, xs, tables, mods, B)
        if return_partial:
            return xs.view(B, N, C), mods
        if modifiers is not None:
            mods = modifiers
        return E.vggt_ffn_part(self, xs, mods).view(B, N, C)

    def forward_partial(self, *args, **kwargs):
        return self.forward(*args, **kwargs, return_partial=True)

    def forward_remaining(self, x: Tensor, e=None):
        return self.forward(x, run_remaining=True, modifiers=e)


class CamTokenProjector(nn.Module):
    """[B, V, 9] pose encodings -> one camera token per latent frame (groups of 4 views; the first view is repeated
    3x so that V+3 is a multiple of 4).  Tiny (V/4 rows): plain torch.  ref: block.py:276-297."""

    def __init__(self, out_dim: int, hidden: int = 128):
        super().__init__()
        self.out_dim = out_dim
        self.mlp = nn.Sequential(nn.Linear(36, hidden), nn.GELU(), nn.Linear(hidden, out_dim))

    def forward(self, cam: torch.Tensor) -> torch.Tensor:
        b = cam.shape[0]
        padded = torch.cat([cam, cam[:, :1].expand(b, 3, cam.shape[2])], dim=1)
        groups = padded.reshape(b * (padded.shape[1] // 4), 36)
        return self.mlp(groups).view(-1, 1, self.out_dim)
