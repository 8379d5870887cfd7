"""B200-native mirror of FantasyWorld/vggt/layers/attention.py (reference).

qkv GEMM -> fused per-head LayerNorm(64) + 2-D RoPE on q,k (fwb_ln64_rope2d, in place on the packed qkv buffer)
-> fwb_attn_fwd reading q/k/v as strided views of that buffer -> proj GEMM.  ref: attention.py:50-72.
"""
from torch import Tensor, nn

from fwb200 import engine as E
from fwb200 import ops


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, proj_bias: bool = True, attn_drop: float = 0.0,
                 proj_drop: float = 0.0, norm_layer: nn.Module = nn.LayerNorm, qk_norm: bool = False,
                 fused_attn: bool = True, rope=None) -> None:
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.fused_attn = fused_attn
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.k_norm = norm_layer(self.head_dim) if qk_norm else nn.Identity()
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim, bias=proj_bias)
        self.proj_drop = nn.Dropout(proj_drop)
        self.rope = rope

    def forward(self, x: Tensor, pos=None) -> Tensor:
        B, N, C = x.shape
        qkv = E.lin(E.as_bf16(x).reshape(B * N, C), self.qkv)
        if self.rope is not None and pos is not None:
            assert self.head_dim == 64 and isinstance(self.q_norm, nn.LayerNorm), "fused qk-norm/RoPE kernel: head_dim 64"
            cosT, sinT = E.rope2d_expanded(pos, self.rope.base_frequency)
            ops.ln64_rope2d_(qkv, self.num_heads, eps=self.q_norm.eps, qw=E.f32(self.q_norm, "w", self.q_norm.weight),
                             qb=E.f32(self.q_norm, "b", self.q_norm.bias), kw=E.f32(self.k_norm, "w", self.k_norm.weight),
                             kb=E.f32(self.k_norm, "b", self.k_norm.bias), cosT=cosT, sinT=sinT)
        else:
            assert isinstance(self.q_norm, nn.Identity), "qk-norm without RoPE is not a reference configuration"
        q5 = qkv.view(B, N, 3, self.num_heads, self.head_dim)
        o = ops.attention(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2])
        return E.lin(o.view(B * N, C), self.proj, round_flags=ops.ROUND_AFTER_BIAS).view(B, N, C)
