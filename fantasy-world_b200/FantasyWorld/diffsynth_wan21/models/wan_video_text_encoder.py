"""Mirror of FantasyWorld/diffsynth_wan21/models/wan_video_text_encoder.py (umT5-XXL encoder).

Same class names, constructor arguments and state_dict keys as the reference (wan_video_text_encoder.py:21-254), so
`models_t5_umt5-xxl-enc-bf16.pth` loads unchanged.  It runs once per prompt, outside the denoising loop
(inference_wan21.py:302-307 -> pipelines/wan_video.py:213-216 -> prompters/wan_prompter.py:98-109).

Where the work goes on the B200:
  * every nn.Linear (q/k/v/o, gate/fc1/fc2: 99.9 % of the FLOPs)  -> fwb_gemm_bf16; q|k|v share one launch, the gate's tanh
    GELU is its GEMM's epilogue, the residual add is the o / fc2 epilogue (the nn.Linear output is rounded to bf16 first, as the
    reference's bf16 module does);
  * T5LayerNorm (RMS, no mean)                                   -> fwb_rmsnorm_rope without the rotary part: same rounding
    points as the reference's bf16 run, bf16(bf16(x * rstd) * w) (wan_video_text_encoder.py:30-35);
  * the attention core: 64 heads x 512 x 512 with an additive per-head relative-position bias and a padding mask, no
    1/sqrt(d) scaling (wan_video_text_encoder.py:69-82).  The fwb200 attention kernel has no bias operand; at 512 tokens the core
    is 0.1 % of the encoder's FLOPs, so it stays on torch (bf16 matmul, fp32 softmax — the reference's arithmetic).
The model is bf16-only (the pipeline dtype, inference_wan21.py:165,226) and CUDA-only: there is no CPU fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from fwb200 import engine as E
from fwb200 import ops

BF16 = torch.bfloat16


def fp16_clamp(x):
    """fp16 overflow guard of the reference (wan_video_text_encoder.py:8-12); bf16 cannot overflow here, so it is the identity on
    this path."""
    if x.dtype == torch.float16 and torch.isinf(x).any():
        lim = torch.finfo(x.dtype).max - 1000
        x = x.clamp(-lim, lim)
    return x


class GELU(nn.Module):
    """tanh-approximated GELU (wan_video_text_encoder.py:15-19).  As a stand-alone module it is evaluated in fp32 and rounded
    once; inside T5FeedForward it is the epilogue of the gate GEMM."""

    def forward(self, x):
        xf = x.float()
        return (0.5 * xf * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (xf + 0.044715 * xf * xf * xf)))).to(x.dtype)


def _rows(x):
    return x.reshape(-1, x.shape[-1])


class T5LayerNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.dim, self.eps = dim, eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        ops.require_device()
        y = E.as_bf16(x).clone(memory_format=torch.contiguous_format)
        ops.rmsnorm_rope_(_rows(y), w=E.f32(self, "w", self.weight), eps=self.eps)
        return y


class T5Attention(nn.Module):
    def __init__(self, dim, dim_attn, num_heads, dropout=0.1):
        assert dim_attn % num_heads == 0
        super().__init__()
        self.dim, self.dim_attn, self.num_heads = dim, dim_attn, num_heads
        self.head_dim = dim_attn // num_heads
        self.q = nn.Linear(dim, dim_attn, bias=False)
        self.k = nn.Linear(dim, dim_attn, bias=False)
        self.v = nn.Linear(dim, dim_attn, bias=False)
        self.o = nn.Linear(dim_attn, dim, bias=False)
        self.dropout = nn.Dropout(dropout)          # inference only: never applied

    def _qkv_weight(self):
        return E.derived(self, "qkv", lambda a, b, c: torch.cat([a, b, c], 0).to(BF16).contiguous(), self.q.weight, self.k.weight,
                         self.v.weight)

    def core(self, q, k, v, mask=None, pos_bias=None):
        """softmax(q k^T + bias) v on [B, L, N, C] views — no 1/sqrt(C) (T5)."""
        b, lq, n, _ = q.shape
        scores = torch.matmul(q.permute(0, 2, 1, 3), k.permute(0, 2, 3, 1))          # [B, N, Lq, Lk] bf16
        bias = scores.new_zeros(b, n, lq, k.shape[1])
        if pos_bias is not None:
            bias = bias + pos_bias
        if mask is not None:
            assert mask.ndim in (2, 3)
            m = mask.view(b, 1, 1, -1) if mask.ndim == 2 else mask.unsqueeze(1)
            bias = bias.masked_fill(m == 0, torch.finfo(scores.dtype).min)
        p = torch.softmax((scores + bias).float(), dim=-1).to(scores.dtype)
        return torch.matmul(p, v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)             # [B, Lq, N, C]

    def forward(self, x, context=None, mask=None, pos_bias=None, resid=None):
        """x [B, L1, C]; context [B, L2, C] or None; mask [B, L2] / [B, L1, L2] or None.  `resid` (extension): added to the
        output inside the o-projection's epilogue."""
        b, n, c = x.size(0), self.num_heads, self.head_dim
        if context is None:
            qkv = ops.linear(x, self._qkv_weight()).view(b, -1, 3, n, c)
            q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        else:
            q = E.lin(x, self.q).view(b, -1, n, c)
            k = E.lin(context, self.k).view(b, -1, n, c)
            v = E.lin(context, self.v).view(b, -1, n, c)
        a = self.core(q, k, v, mask, pos_bias).reshape(b, -1, n * c)
        return E.lin(a, self.o, resid=resid, round_flags=E.ROUND_AFTER_BIAS if resid is not None else 0)


class T5FeedForward(nn.Module):
    def __init__(self, dim, dim_ffn, dropout=0.1):
        super().__init__()
        self.dim, self.dim_ffn = dim, dim_ffn
        self.gate = nn.Sequential(nn.Linear(dim, dim_ffn, bias=False), GELU())
        self.fc1 = nn.Linear(dim, dim_ffn, bias=False)
        self.fc2 = nn.Linear(dim_ffn, dim, bias=False)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, resid=None):
        g = E.lin(x, self.gate[0], act=E.ACT_GELU_TANH)
        h = E.lin(x, self.fc1) * g
        return E.lin(h, self.fc2, resid=resid, round_flags=E.ROUND_AFTER_BIAS if resid is not None else 0)


class T5RelativeEmbedding(nn.Module):
    def __init__(self, num_buckets, num_heads, bidirectional, max_dist=128):
        super().__init__()
        self.num_buckets, self.num_heads, self.bidirectional, self.max_dist = num_buckets, num_heads, bidirectional, max_dist
        self.embedding = nn.Embedding(num_buckets, num_heads)

    def forward(self, lq, lk):
        dev = self.embedding.weight.device
        rel = torch.arange(lk, device=dev)[None, :] - torch.arange(lq, device=dev)[:, None]
        table = self.embedding(self._relative_position_bucket(rel))                  # [Lq, Lk, N]
        return table.permute(2, 0, 1).unsqueeze(0).contiguous()                      # [1, N, Lq, Lk]

    def _relative_position_bucket(self, rel_pos):
        """T5 bucketing (wan_video_text_encoder.py:170-190): exact buckets for small distances, log-spaced beyond, the sign in
        the upper half when bidirectional."""
        if self.bidirectional:
            nb = self.num_buckets // 2
            base = (rel_pos > 0).long() * nb
            dist = rel_pos.abs()
        else:
            nb = self.num_buckets
            base = torch.zeros_like(rel_pos)
            dist = (-rel_pos).clamp(min=0)
        exact = nb // 2
        far = exact + (torch.log(dist.float() / exact) / math.log(self.max_dist / exact) * (nb - exact)).long()
        far = far.clamp(max=nb - 1)
        return base + torch.where(dist < exact, dist, far)


class T5SelfAttention(nn.Module):
    def __init__(self, dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos=True, dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_buckets, self.shared_pos = num_heads, num_buckets, shared_pos
        self.norm1 = T5LayerNorm(dim)
        self.attn = T5Attention(dim, dim_attn, num_heads, dropout)
        self.norm2 = T5LayerNorm(dim)
        self.ffn = T5FeedForward(dim, dim_ffn, dropout)
        self.pos_embedding = None if shared_pos else T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True)

    def forward(self, x, mask=None, pos_bias=None):
        e = pos_bias if self.shared_pos else self.pos_embedding(x.size(1), x.size(1))
        x = fp16_clamp(self.attn(self.norm1(x), mask=mask, pos_bias=e, resid=x))
        return fp16_clamp(self.ffn(self.norm2(x), resid=x))


def init_weights(m):
    """The reference's initialisation (wan_video_text_encoder.py:192-208)."""
    if isinstance(m, T5LayerNorm):
        nn.init.ones_(m.weight)
    elif isinstance(m, T5FeedForward):
        nn.init.normal_(m.gate[0].weight, std=m.dim ** -0.5)
        nn.init.normal_(m.fc1.weight, std=m.dim ** -0.5)
        nn.init.normal_(m.fc2.weight, std=m.dim_ffn ** -0.5)
    elif isinstance(m, T5Attention):
        nn.init.normal_(m.q.weight, std=(m.dim * m.dim_attn) ** -0.5)
        nn.init.normal_(m.k.weight, std=m.dim ** -0.5)
        nn.init.normal_(m.v.weight, std=m.dim ** -0.5)
        nn.init.normal_(m.o.weight, std=(m.num_heads * m.dim_attn) ** -0.5)
    elif isinstance(m, T5RelativeEmbedding):
        nn.init.normal_(m.embedding.weight, std=(2 * m.num_buckets * m.num_heads) ** -0.5)


class WanTextEncoder(nn.Module):
    def __init__(self, vocab=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32,
                 shared_pos=False, dropout=0.1):
        super().__init__()
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets, self.shared_pos = num_heads, num_layers, num_buckets, shared_pos
        self.token_embedding = vocab if isinstance(vocab, nn.Embedding) else nn.Embedding(vocab, dim)
        self.pos_embedding = T5RelativeEmbedding(num_buckets, num_heads, bidirectional=True) if shared_pos else None
        self.dropout = nn.Dropout(dropout)
        self.blocks = nn.ModuleList([T5SelfAttention(dim, dim_attn, dim_ffn, num_heads, num_buckets, shared_pos, dropout)
                                     for _ in range(num_layers)])
        self.norm = T5LayerNorm(dim)
        if not any(p.is_meta for p in self.parameters()):
            self.apply(init_weights)

    @torch.no_grad()
    def forward(self, ids, mask=None):
        ops.require_device()
        assert not self.training, "WanTextEncoder mirror: inference only (dropout is never applied)"
        x = E.as_bf16(self.token_embedding(ids))
        e = self.pos_embedding(x.size(1), x.size(1)) if self.shared_pos else None
        for block in self.blocks:
            x = block(x, mask, pos_bias=e)
        return self.norm(x)

    @staticmethod
    def state_dict_converter():
        return WanTextEncoderStateDictConverter()


class WanTextEncoderStateDictConverter:
    def from_diffusers(self, state_dict):
        return state_dict

    def from_civitai(self, state_dict):
        return state_dict
