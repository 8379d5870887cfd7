"""fw_oracle.py — CPU restatement (torch fp32, functional) of the FantasyWorld denoising hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product path (fantasy-world_b200/) never does.

Every function restates one piece of the reference algorithm and cites the reference file:line it follows
(paths relative to /root/reference/FantasyWorld).  The restatement is functional: it takes the reference's own
state_dict (flat {key: tensor}) plus a key prefix, so parity with the reference is checked by feeding both the same
weights.  It is pinned against the real reference, imported through tools/ref_shim.py, by tools/make_golden.py
(fixtures under tests/golden/) and tests/test_oracle_golden.py.  The reference itself ships no tests or golden vectors
(SURVEY.md §4), so these generated fixtures are the pin.

`emulate_bf16=True` inserts bf16 roundings at the points where the reference's CUDA-autocast run rounds
(SURVEY.md Appendix A); with False everything is plain fp32 (the mode the fixtures were generated in).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class Numerics:
    emulate_bf16: bool = False

    def r(self, x: torch.Tensor) -> torch.Tensor:
        """Round to bf16 (and come back to fp32) when emulating the CUDA autocast run."""
        if self.emulate_bf16:
            return x.to(torch.bfloat16).to(torch.float32)
        return x


FP32 = Numerics(False)
BF16 = Numerics(True)


# ------------------------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------------------------
def linear(sd, pfx, x, nm=FP32):
    """nn.Linear under autocast: bf16 operands, fp32 accumulate, bf16 result."""
    w = sd[pfx + ".weight"].float()
    b = sd.get(pfx + ".bias")
    y = nm.r(x) @ nm.r(w).t()
    if b is not None:
        y = y + nm.r(b.float())
    return nm.r(y)


def layer_norm(x, eps, w=None, b=None):
    """F.layer_norm — fp32 under autocast."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    y = (x - mu) * torch.rsqrt(var + eps)
    if w is not None:
        y = y * w.float() + b.float()
    return y


def rms_norm(x, w, eps, nm=FP32):
    """RMSNorm over the FULL channel dim — diffsynth_wan21/models/wan_video_dit.py:135-146."""
    y = x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps)
    return nm.r(nm.r(y) * nm.r(w.float()))


USE_TORCH_SDPA = False   # bench.py's timed CPU legs set this: call F.scaled_dot_product_attention (what the reference itself calls
                         # on CPU, wan_video_dit.py:60-65) instead of the explicit restatement below; fp32 mode only


def sdpa(q, k, v, nm=FP32):
    """Non-causal softmax attention, scale 1/sqrt(D); q,k,v [B,H,L,D] — wan_video_dit.py:60-65, F.sdpa."""
    q, k, v = nm.r(q), nm.r(k), nm.r(v)
    if USE_TORCH_SDPA and nm is FP32:
        return torch.nn.functional.scaled_dot_product_attention(q, k, v)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    p = torch.softmax(s, dim=-1)
    return nm.r(p @ v)


def heads_split(x, n):  # [B,L,(n d)] -> [B,n,L,d]
    B, L, C = x.shape
    return x.view(B, L, n, C // n).transpose(1, 2)


def heads_merge(x):  # [B,n,L,d] -> [B,L,(n d)]
    B, n, L, d = x.shape
    return x.transpose(1, 2).reshape(B, L, n * d)


# ------------------------------------------------------------------------------------------------------------------
# timestep embeddings and RoPE tables
# ------------------------------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim, position):
    """wan_video_dit.py:73-77 (result cast to position.dtype) / wan/modules/model.py:17-27 (kept fp64)."""
    half = dim // 2
    sinusoid = torch.outer(position.double(), torch.pow(10000.0, -torch.arange(half, dtype=torch.float64) / half))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def freqs_cis_1d(dim, end=1024, theta=10000.0):
    """wan_video_dit.py:88-94 precompute_freqs_cis: complex128 table [end, dim/2]."""
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].double() / dim))
    ang = torch.outer(torch.arange(end).double(), inv)
    return torch.polar(torch.ones_like(ang), ang)


def freqs_cis_3d(head_dim):
    """wan_video_dit.py:80-85: head_dim split as (d - 2*(d//3), d//3, d//3) over (f, h, w)."""
    return (freqs_cis_1d(head_dim - 2 * (head_dim // 3)), freqs_cis_1d(head_dim // 3), freqs_cis_1d(head_dim // 3))


def rope_table_3d(head_dim, f, h, w):
    """model_wan21.py:132-136 — per-token complex table [f*h*w, head_dim/2] in (f h w) order."""
    tf, th, tw = freqs_cis_3d(head_dim)
    return torch.cat([
        tf[:f].view(f, 1, 1, -1).expand(f, h, w, -1),
        th[:h].view(1, h, 1, -1).expand(f, h, w, -1),
        tw[:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(f * h * w, -1)


def rope_table_3d_with_extra(head_dim, f, h, w, n_extra):
    """wan_video_dit.py:105-132 build_freqs_3d_with_extra_cis: n_extra identity rotations in front of every frame."""
    patch = rope_table_3d(head_dim, f, h, w).view(f, h * w, -1)
    extra = torch.ones(f, n_extra, patch.shape[-1], dtype=patch.dtype)
    return torch.cat([extra, patch], dim=1).reshape(f * (n_extra + h * w), -1)


def rope_apply(x, table, n_heads, nm=FP32):
    """wan_video_dit.py:97-102: interleaved (2i,2i+1) complex pairs, fp64 multiply, cast back."""
    B, L, C = x.shape
    xc = torch.view_as_complex(x.double().reshape(B, L, n_heads, -1, 2))
    out = torch.view_as_real(xc * table.view(1, L, 1, -1)).flatten(3).reshape(B, L, C)
    return nm.r(out.float())


def rope2d_tables(dim_half, max_pos, base=100.0, nm=FP32):
    """vggt/layers/rope.py:82-110 (fp32 tokens): cos/sin [max_pos, dim_half] with the angle vector duplicated.
    Under CUDA autocast the einsum that forms the angles runs in bf16 (einsum is an autocast-to-bf16 op and the table is
    built on first use inside the sampler's autocast region): angle = bf16(bf16(pos) * bf16(inv_freq))."""
    exponents = torch.arange(0, dim_half, 2).float() / dim_half
    inv_freq = 1.0 / (base ** exponents)
    positions = torch.arange(max_pos, dtype=inv_freq.dtype)
    if nm.emulate_bf16:
        ang = (positions.to(torch.bfloat16)[:, None] * inv_freq.to(torch.bfloat16)[None, :]).float()
    else:
        ang = torch.einsum("i,j->ij", positions, inv_freq)
    ang = torch.cat((ang, ang), dim=-1)
    return ang.cos(), ang.sin()


def rope2d_apply(t, pos, base=100.0, nm=FP32):
    """vggt/layers/rope.py:133-188: t [B,H,N,D]; first D/2 features rotate with y, last D/2 with x; rotate-half."""
    D2 = t.shape[-1] // 2
    cos_t, sin_t = rope2d_tables(D2, int(pos.max()) + 1, base, nm)

    def one(feat, p):
        cos = F.embedding(p, cos_t)[:, None]
        sin = F.embedding(p, sin_t)[:, None]
        h = feat.shape[-1] // 2
        rot = torch.cat((-feat[..., h:], feat[..., :h]), dim=-1)
        return feat * cos + rot * sin

    return torch.cat((one(t[..., :D2], pos[..., 0]), one(t[..., D2:], pos[..., 1])), dim=-1)


# ------------------------------------------------------------------------------------------------------------------
# DiT block (video branch)
# ------------------------------------------------------------------------------------------------------------------
DIT_HEADS = 40


def dit_self_attn(sd, pfx, x, rope_tab, nm=FP32):
    """SelfAttention.forward — wan_video_dit.py:175-182."""
    q = rms_norm(linear(sd, pfx + ".q", x, nm), sd[pfx + ".norm_q.weight"], 1e-6, nm)
    k = rms_norm(linear(sd, pfx + ".k", x, nm), sd[pfx + ".norm_k.weight"], 1e-6, nm)
    v = linear(sd, pfx + ".v", x, nm)
    q = rope_apply(q, rope_tab, DIT_HEADS, nm)
    k = rope_apply(k, rope_tab, DIT_HEADS, nm)
    o = heads_merge(sdpa(heads_split(q, DIT_HEADS), heads_split(k, DIT_HEADS), heads_split(v, DIT_HEADS), nm))
    return linear(sd, pfx + ".o", o, nm)


def dit_cross_attn(sd, pfx, x, context, plucker_fea, nm=FP32):
    """CrossAttentionProcessor / CrossAttentionAdapterProcessor('adaln') —
    wan_video_dit.py:185-201, camera_control.py:92-148.  context = [clip(257) | text] when the block has image input
    (Wan2.1-I2V: k_img / v_img weights present), else text only (Wan2.2-Fun, diffsynth_wan22/models/wan_video_dit.py:203-223)."""
    has_img = (pfx + ".k_img.weight") in sd
    img, ctx = (context[:, :257], context[:, 257:]) if has_img else (None, context)
    q = rms_norm(linear(sd, pfx + ".q", x, nm), sd[pfx + ".norm_q.weight"], 1e-6, nm)
    k = rms_norm(linear(sd, pfx + ".k", ctx, nm), sd[pfx + ".norm_k.weight"], 1e-6, nm)
    v = linear(sd, pfx + ".v", ctx, nm)
    qh = heads_split(q, DIT_HEADS)
    o = heads_merge(sdpa(qh, heads_split(k, DIT_HEADS), heads_split(v, DIT_HEADS), nm))
    if has_img:
        k_img = rms_norm(linear(sd, pfx + ".k_img", img, nm), sd[pfx + ".norm_k_img.weight"], 1e-6, nm)
        v_img = linear(sd, pfx + ".v_img", img, nm)
        o_img = heads_merge(sdpa(qh, heads_split(k_img, DIT_HEADS), heads_split(v_img, DIT_HEADS), nm))
        o = nm.r(o + o_img)
    has_adapter = (pfx + ".processor.k_proj.group1.weight") in sd
    if has_adapter and plucker_fea is not None and not bool(torch.all(plucker_fea == 0)):
        # GroupLinearDualK (camera_control.py:24-39), GroupLinearDualV (:42-63): scale is the float 0.0
        p1 = linear(sd, pfx + ".processor.k_proj.group1", plucker_fea, nm)
        h = nm.r(F.relu(linear(sd, pfx + ".processor.k_proj.group2.0", o, nm)))
        comb = nm.r(linear(sd, pfx + ".processor.k_proj.group2.2", h, nm) + p1)
        h2 = nm.r(F.relu(linear(sd, pfx + ".processor.v_proj.group2.0", comb, nm)))
        shift = linear(sd, pfx + ".processor.v_proj.group2.2", h2, nm)
        o = nm.r(o + shift)  # x * (0.0 + 1.) + shift
    return linear(sd, pfx + ".o", o, nm)


def dit_modulation(sd, pfx, t_mod, nm=FP32):
    """(modulation + t_mod).chunk(6) — wan_video_dit.py:296-299 (bf16 add under autocast)."""
    m = nm.r(nm.r(sd[pfx + ".modulation"].float()) + nm.r(t_mod))
    return m.chunk(6, dim=1)


def modulate(xn, shift, scale, nm=FP32):
    """wan_video_dit.py:69-70: (1 + scale) is rounded to bf16 before the fp32 multiply."""
    return xn * nm.r(1 + scale) + shift


def dit_block_partial(sd, pfx, x, context, t_mod, rope_tab, plucker_fea, nm=FP32):
    """DiTBlock.forward(return_partial=True) — wan_video_dit.py:296-306."""
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = dit_modulation(sd, pfx, t_mod, nm)
    h = modulate(layer_norm(x, 1e-6), shift_msa, scale_msa, nm)
    a = dit_self_attn(sd, pfx + ".self_attn", h, rope_tab, nm)
    x = nm.r(x + nm.r(gate_msa * a))
    n3 = layer_norm(x, 1e-6, sd[pfx + ".norm3.weight"], sd[pfx + ".norm3.bias"])
    x = nm.r(x + dit_cross_attn(sd, pfx + ".cross_attn", n3, context, plucker_fea, nm))
    return x, (shift_mlp, scale_mlp, gate_mlp)


def dit_block_remaining(sd, pfx, x, mods, nm=FP32):
    """DiTBlock.forward(run_remaining=True) — wan_video_dit.py:288-294; FFN = Linear, GELU(tanh), Linear (:274-275)."""
    shift_mlp, scale_mlp, gate_mlp = mods
    h = modulate(layer_norm(x, 1e-6), shift_mlp, scale_mlp, nm)
    h = nm.r(F.gelu(linear(sd, pfx + ".ffn.0", h, nm), approximate="tanh"))
    y = linear(sd, pfx + ".ffn.2", h, nm)
    return nm.r(x + nm.r(gate_mlp * y))


def dit_block(sd, pfx, x, context, t_mod, rope_tab, plucker_fea, nm=FP32):
    x, mods = dit_block_partial(sd, pfx, x, context, t_mod, rope_tab, plucker_fea, nm)
    return dit_block_remaining(sd, pfx, x, mods, nm)


# ------------------------------------------------------------------------------------------------------------------
# VGGT block (geometry branch)
# ------------------------------------------------------------------------------------------------------------------
VGGT_HEADS = 16


def vggt_attention(sd, pfx, x, pos, nm=FP32):
    """Attention.forward — vggt/layers/attention.py:50-72 (per-head LayerNorm qk-norm, 2-D RoPE, SDPA, proj)."""
    B, N, C = x.shape
    qkv = linear(sd, pfx + ".qkv", x, nm).reshape(B, N, 3, VGGT_HEADS, C // VGGT_HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    q = layer_norm(q, 1e-5, sd[pfx + ".q_norm.weight"], sd[pfx + ".q_norm.bias"])
    k = layer_norm(k, 1e-5, sd[pfx + ".k_norm.weight"], sd[pfx + ".k_norm.bias"])
    q, k = rope2d_apply(q, pos, nm=nm), rope2d_apply(k, pos, nm=nm)
    o = sdpa(q, k, v, nm).transpose(1, 2).reshape(B, N, C)
    return linear(sd, pfx + ".proj", o, nm)


def vggt_modulation(sd, pfx, e0, batch):
    """vggt/layers/block.py:95-104: e0 [B0,6,C] fp32 repeated over frames; modulation(bf16)+e0 -> fp32."""
    B0 = e0.shape[0]
    if B0 != batch:
        e0 = e0.unsqueeze(1).repeat(1, batch // B0, 1, 1).reshape(batch, 6, -1)
    return (sd[pfx + ".modulation"].float() + e0).chunk(6, dim=1)


def vggt_block_partial(sd, pfx, x, pos, e0, nm=FP32):
    """Block.forward(return_partial=True) — vggt/layers/block.py:73-76, 106-109."""
    e = vggt_modulation(sd, pfx, e0, x.shape[0])
    h = layer_norm(x, 1e-5, sd[pfx + ".norm1.weight"], sd[pfx + ".norm1.bias"]) * (1 + e[1]) + e[0]
    a = vggt_attention(sd, pfx + ".attn", h, pos, nm)
    x = x + nm.r(a * nm.r(sd[pfx + ".ls1.gamma"].float()))
    return x, e


def vggt_block_remaining(sd, pfx, x, e, nm=FP32):
    """Block.forward(run_remaining=True) — vggt/layers/block.py:78-81: modulation applied AFTER the MLP."""
    h = layer_norm(x, 1e-5, sd[pfx + ".norm2.weight"], sd[pfx + ".norm2.bias"])
    h = nm.r(F.gelu(linear(sd, pfx + ".mlp.fc1", h, nm)))
    y = linear(sd, pfx + ".mlp.fc2", h, nm)
    y = (y * (1 + e[4]) + e[3]) * nm.r(sd[pfx + ".ls2.gamma"].float()) * e[5]
    return x + y


def vggt_block(sd, pfx, x, pos, e0, nm=FP32):
    x, e = vggt_block_partial(sd, pfx, x, pos, e0, nm)
    return vggt_block_remaining(sd, pfx, x, e, nm)


# ------------------------------------------------------------------------------------------------------------------
# bidirectional adapter and the IRG block
# ------------------------------------------------------------------------------------------------------------------
BI_HEADS = 12


def bicross(sd, pfx, x1, x2, tab_dit, tab_agg, nm=FP32):
    """CrossModalityBiAttentionBlock.forward + BiMultiHeadAttention.forward_sdpa —
    fusion/layer/block.py:179-221, 532-625.  Attention scale is SDPA's default 1/sqrt(96)."""
    c = pfx + ".cross_attn"
    n1, n2 = layer_norm(x1, 1e-6), layer_norm(x2, 1e-6)
    q = rope_apply(linear(sd, c + ".m1_proj", n1, nm), tab_dit, BI_HEADS, nm)
    k = rope_apply(linear(sd, c + ".m2_proj", n2, nm), tab_agg, BI_HEADS, nm)
    v1 = linear(sd, c + ".values_m1_proj", n1, nm)
    v2 = linear(sd, c + ".values_m2_proj", n2, nm)
    qh, kh = heads_split(q, BI_HEADS), heads_split(k, BI_HEADS)
    o1 = heads_merge(sdpa(qh, kh, heads_split(v2, BI_HEADS), nm))
    o2 = heads_merge(sdpa(kh, qh, heads_split(v1, BI_HEADS), nm))
    d1 = linear(sd, c + ".out_m1_proj", o1, nm)
    d2 = linear(sd, c + ".out_m2_proj", o2, nm)
    x1 = nm.r(x1 + nm.r(nm.r(sd[pfx + ".gamma_m1"].float()) * d1))
    x2 = x2 + nm.r(nm.r(sd[pfx + ".gamma_m2"].float()) * d2)
    return x1, x2


def irg_block(sd, pfx, x_dit, x_agg, context, t_mod, rope_tab, tab_bi_dit, tab_bi_agg, pos, e0, plucker_fea,
              uncond=False, nm=FP32):
    """IRGBlock._forward_impl — fusion/layer/block.py:43-94.  x_agg [(b s), p, d], pos [(b s), p, 2]."""
    S, P, D = x_agg.shape
    B = x_dit.shape[0]
    xd, mods = dit_block_partial(sd, pfx + ".x_dit", x_dit, context, t_mod, rope_tab, plucker_fea, nm)
    pos_g = pos.reshape(B, -1, 2)
    xa = x_agg.reshape(B, -1, D)
    xa, e = vggt_block_partial(sd, pfx + ".x_agg", xa, pos_g, e0, nm)
    if not uncond:
        xd, xa = bicross(sd, pfx + ".bicross_attention", xd, xa, tab_bi_dit, tab_bi_agg, nm)
    xd = dit_block_remaining(sd, pfx + ".x_dit", xd, mods, nm)
    xa = vggt_block_remaining(sd, pfx + ".x_agg", xa, e, nm)
    return xd, xa, xa.view(B, -1, P, D)


# ------------------------------------------------------------------------------------------------------------------
# embeddings, patchify, head
# ------------------------------------------------------------------------------------------------------------------
def dit_time_embed(sd, pfx, timestep, nm=FP32):
    """model_wan21.py:119-122: t = time_embedding(sinusoid(256, t)), t_mod = time_projection(t) [1,6,5120]."""
    s = sinusoidal_embedding_1d(256, timestep).float()
    s = nm.r(s)  # cast to the timestep dtype (bf16 in the CUDA run)
    t = linear(sd, pfx + ".time_embedding.2", nm.r(F.silu(linear(sd, pfx + ".time_embedding.0", s, nm))), nm)
    t_mod = linear(sd, pfx + ".time_projection.1", nm.r(F.silu(t)), nm).unflatten(1, (6, 5120))
    return t, t_mod


def vggt_time_embed(sd, pfx, timestep):
    """vggt/models/vggt.py:126-130: fp32 throughout (autocast(dtype=float32))."""
    s = sinusoidal_embedding_1d(256, timestep).float()
    e = F.linear(F.silu(F.linear(s, sd[pfx + ".time_embedding.0.weight"].float(), sd[pfx + ".time_embedding.0.bias"].float())),
                 sd[pfx + ".time_embedding.2.weight"].float(), sd[pfx + ".time_embedding.2.bias"].float())
    e0 = F.linear(F.silu(e), sd[pfx + ".time_projection.1.weight"].float(), sd[pfx + ".time_projection.1.bias"].float())
    return e0.unflatten(1, (6, 1024))


def text_embed(sd, pfx, context, nm=FP32):
    """WanModel.text_embedding — wan_video_dit.py:387-391."""
    return linear(sd, pfx + ".text_embedding.2",
                  nm.r(F.gelu(linear(sd, pfx + ".text_embedding.0", context, nm), approximate="tanh")), nm)


def img_embed(sd, pfx, clip, nm=FP32):
    """MLP(1280 -> 5120) — wan_video_dit.py:324-341: LN, Linear, GELU(erf), Linear, LN."""
    p = pfx + ".img_emb.proj"
    h = layer_norm(clip, 1e-5, sd[p + ".0.weight"], sd[p + ".0.bias"])
    h = nm.r(F.gelu(linear(sd, p + ".1", h, nm)))
    h = linear(sd, p + ".3", h, nm)
    return layer_norm(h, 1e-5, sd[p + ".4.weight"], sd[p + ".4.bias"])


def patchify(sd, pfx, x, nm=FP32):
    """WanModel.patchify — wan_video_dit.py:424-435: Conv3d(36->5120, k=s=(1,2,2)) then 'b c f h w -> b (f h w) c'."""
    w = sd[pfx + ".patch_embedding.weight"].float()
    b = sd[pfx + ".patch_embedding.bias"].float()
    y = nm.r(F.conv3d(nm.r(x), nm.r(w), nm.r(b), stride=(1, 2, 2)))
    B, C, f, h, ww = y.shape
    return y.permute(0, 2, 3, 4, 1).reshape(B, f * h * ww, C), (f, h, ww)


def dit_head(sd, pfx, x, t, nm=FP32):
    """Head.forward — wan_video_dit.py:353-358 (called with t, not t_mod: model_wan21.py:214)."""
    m = nm.r(nm.r(sd[pfx + ".head.modulation"].float()) + nm.r(t).unsqueeze(1))
    shift, scale = m.chunk(2, dim=1)
    return linear(sd, pfx + ".head.head", layer_norm(x, 1e-6) * nm.r(1 + scale) + shift, nm)


def unpatchify(x, grid):
    """wan_video_dit.py:437-442: 'b (f h w) (x y z c) -> b c (f x) (h y) (w z)' with (x,y,z) = (1,2,2)."""
    f, h, w = grid
    B = x.shape[0]
    c = x.shape[-1] // 4
    x = x.view(B, f, h, w, 1, 2, 2, c).permute(0, 7, 1, 4, 2, 5, 3, 6)
    return x.reshape(B, c, f, h * 2, w * 2)


def aggregator_input(sd, pfx, patch_token):
    """Aggregator._process_aggregator_input + slice_expand_and_flatten — vggt/models/aggregator.py:261-306.
    patch_token [B,T,h,w,1024] -> tokens [(B T), 5+h*w, 1024], pos int64 [(B T), 5+h*w, 2] (patches +1, specials 0)."""
    B, T, h, w, C = patch_token.shape
    pt = patch_token.reshape(B * T, h * w, C)

    def special(tok):  # [1,2,X,C]: index 0 for the first frame, index 1 for the rest
        first = tok[:, 0:1].expand(B, 1, *tok.shape[2:])
        rest = tok[:, 1:].expand(B, T - 1, *tok.shape[2:])
        return torch.cat([first, rest], dim=1).reshape(B * T, *tok.shape[2:])

    cam = special(sd[pfx + ".camera_token"].float())
    reg = special(sd[pfx + ".register_token"].float())
    tokens = torch.cat([cam, reg, pt], dim=1)
    ys, xs = torch.arange(h), torch.arange(w)
    grid = torch.cartesian_prod(ys, xs).view(1, h * w, 2).expand(B * T, -1, -1) + 1
    pos = torch.cat([torch.zeros(B * T, 5, 2, dtype=grid.dtype), grid], dim=1)
    return tokens, pos


# ------------------------------------------------------------------------------------------------------------------
# joint_forward (without the geometry heads) and the sampler step
# ------------------------------------------------------------------------------------------------------------------
def control_adapter(sd, pfx, control, nm=FP32):
    """SimpleAdapter — diffsynth_wan22/models/wan_video_camera_controller.py:8-47, 64-76: PixelUnshuffle(8) ->
    Conv2d(k = s = 2) -> x + conv2(relu(conv1(x))); frames folded into the batch.  Returns tokens [1, f*h*w, dim]."""
    b, c, f, H, W = control.shape
    z = F.pixel_unshuffle(control.permute(0, 2, 1, 3, 4).reshape(b * f, c, H, W), 8)
    p = pfx + ".control_adapter"
    z = nm.r(F.conv2d(nm.r(z), nm.r(sd[p + ".conv.weight"].float()), nm.r(sd[p + ".conv.bias"].float()), stride=2))
    r = p + ".residual_blocks.0"
    h1 = nm.r(F.relu(nm.r(F.conv2d(z, nm.r(sd[r + ".conv1.weight"].float()), nm.r(sd[r + ".conv1.bias"].float()), padding=1))))
    z = nm.r(z + nm.r(F.conv2d(h1, nm.r(sd[r + ".conv2.weight"].float()), nm.r(sd[r + ".conv2.bias"].float()), padding=1)))
    return z.view(b, f, *z.shape[1:]).permute(0, 1, 3, 4, 2).reshape(b, -1, z.shape[1])


def joint_forward(sd, x, timestep, context, clip_feature, y, plucker_fea, start_index, n_irg, nm=FP32,
                  collect_intermediates=False, control=None):
    """FantasyWorldFusionModel.joint_forward — fusion/model_wan21.py:104-224 (heads excluded); with clip_feature=None and
    `control` given it is the Wan2.2 variant, fusion/model_wan22.py:226-348 (no CLIP context, control adapter added to the
    patch embedding, no camera AdaLN).  Returns the predicted latent [1,16,f,H,W] and, optionally, the per-layer
    [B,S,P,2C] intermediates."""
    dit = "pipe.dit"
    t, t_mod = dit_time_embed(sd, dit, timestep, nm)
    ctx = text_embed(sd, dit, context, nm)
    x = torch.cat([x, y], dim=1)
    if clip_feature is not None:
        ctx = torch.cat([nm.r(img_embed(sd, dit, clip_feature, nm)), ctx], dim=1)
    x, (f, h, w) = patchify(sd, dit, x, nm)
    if control is not None:
        x = nm.r(x + control_adapter(sd, dit, control, nm))
    tab = rope_table_3d(128, f, h, w)
    tab_bi_dit = rope_table_3d(96, f, h, w)
    tab_bi_agg = rope_table_3d_with_extra(96, f, h, w, 5)
    for i in range(start_index):
        x = dit_block(sd, f"{dit}.blocks.{i}", x, ctx, t_mod, tab, plucker_fea, nm)
    # VGGT._process_wan_input — vggt/models/vggt.py:118-131: Conv3d 1x1x1 5120->1024 over tokens
    wproj = sd["vggt.projection_head.weight"].float().view(1024, 5120)
    patch = nm.r(nm.r(x) @ nm.r(wproj).t() + nm.r(sd["vggt.projection_head.bias"].float()))
    patch_token = patch.view(x.shape[0], f, h, w, 1024)
    e0 = vggt_time_embed(sd, "vggt", timestep)
    tokens, pos = aggregator_input(sd, "vggt.aggregator", patch_token)
    B, S, P = x.shape[0], f, tokens.shape[1]
    inter = []
    for i in range(n_irg):
        tokens = vggt_block(sd, f"vggt.aggregator.frame_blocks.{i}", tokens.reshape(B * S, P, 1024), pos, e0, nm)
        frame_inter = tokens.view(B, S, P, 1024)
        x, tokens, g_inter = irg_block(sd, f"IRGBlock.{i}", x, tokens, ctx, t_mod, tab, tab_bi_dit, tab_bi_agg, pos, e0,
                                       plucker_fea, False, nm)
        if collect_intermediates:
            inter.append(torch.cat([frame_inter, g_inter], dim=-1))
    out = unpatchify(dit_head(sd, dit, x, t, nm), (f, h, w))
    return out, inter, patch_token


def flow_match_sigmas(num_steps, shift=5.0, sigma_min=0.0, sigma_max=1.0):
    """FlowMatchScheduler.set_timesteps(extra_one_step=True) — diffsynth_wan21/schedulers/flow_match.py:18-31."""
    s = torch.linspace(sigma_max, sigma_min, num_steps + 1)[:-1]
    s = shift * s / (1 + (shift - 1) * s)
    return s, s * 1000.0


def denoise_step(sd, latents, step, sigmas, timesteps, ctx_pos, ctx_neg, clip_feature, y, plucker_fea, start_index,
                 n_irg, cfg_scale=5.0, nm=FP32):
    """One iteration of generate_video's loop — fusion/model_wan21.py:289-322 + flow_match.py:43-53."""
    t = nm.r(timesteps[step].reshape(1).float())  # cast to bf16 before the sinusoid in the CUDA run (:292-293)
    pos, _, _ = joint_forward(sd, latents, t, ctx_pos, clip_feature, y, plucker_fea, start_index, n_irg, nm)
    neg, _, _ = joint_forward(sd, latents, t, ctx_neg, clip_feature, y, plucker_fea, start_index, n_irg, nm)
    pred = nm.r(neg + nm.r(cfg_scale * nm.r(pos - neg)))
    sigma = sigmas[step]
    sigma_next = sigmas[step + 1] if step + 1 < len(sigmas) else torch.zeros(())
    return nm.r(latents + nm.r(pred * (sigma_next - sigma)))
