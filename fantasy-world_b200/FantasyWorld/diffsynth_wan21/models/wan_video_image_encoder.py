"""Mirror of FantasyWorld/diffsynth_wan21/models/wan_video_image_encoder.py: the open-clip XLM-Roberta / ViT-H-14 image tower
that produces `clip_feature` for the I2V cross-attention.

Only what `WanImageEncoder.encode_image` touches is mirrored — the vision transformer (wan_video_image_encoder.py:203-478), the
`XLMRobertaCLIP` shell whose `textual` is None in the reference too (:642-706), `clip_xlm_roberta_vit_h_14` (:822-849) and
`WanImageEncoder` (:852-899) — with the reference's class names, constructor arguments and state_dict keys (`model.visual.*`,
`model.log_scale`), so `models_clip_open-clip-xlm-roberta-large-vit-huge-14.pth` loads through the same converter.  The
XLM-Roberta text tower classes (:14-200, :617-640) are never instantiated on this path and are not mirrored.

Where the work goes on the B200 (once per sample: 257 tokens x 31 blocks at width 1280):
  * patch embedding (14x14 stride-14 convolution)  -> unfold + fwb_gemm_bf16 (K = 588, zero-padded to 592: exact);
  * LayerNorms                                     -> fwb_ln_modulate;
  * to_qkv / proj / mlp                            -> fwb_gemm_bf16 (bias, erf-GELU / quick-GELU and the residual add in the epilogue);
  * attention, head_dim 80                         -> fwb_attn_fwd on its head_dim-96 instance.  The 16 zero columns per head are
    produced by the qkv GEMM itself (zero rows spliced into the cached weight) and consumed by the projection (zero columns), so
    nothing is padded or copied at run time; softmax scale stays 1/sqrt(80).
bf16 and CUDA only, like the pipeline it is used from (inference_wan21.py:165,226).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from fwb200 import engine as E
from fwb200 import ops

BF16 = torch.bfloat16


def padded_head_dim(d: int) -> int:
    for width in (64, 96, 128):
        if d <= width:
            return width
    raise ValueError(f"head_dim {d} > 128 is not supported by fwb_attn_fwd")


def pos_interpolate(pos, seq_len):
    """Bicubic resampling of the patch-position table for another grid (wan_video_image_encoder.py:203-219); host-side, fp32."""
    if pos.size(1) == seq_len:
        return pos
    src, tar = int(math.sqrt(pos.size(1))), int(math.sqrt(seq_len))
    n = pos.size(1) - src * src
    grid = pos[:, n:].float().reshape(1, src, src, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(tar, tar), mode="bicubic", align_corners=False)
    return torch.cat([pos[:, :n], grid.flatten(2).transpose(1, 2)], dim=1)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        ops.require_device()
        return ops.ln_modulate(E.as_bf16(x), eps=self.eps, w=E.f32(self, "w", self.weight), b=E.f32(self, "b", self.bias)).to(x.dtype)


def _splice_heads(t: torch.Tensor, groups: int, heads: int, d: int, dp: int) -> torch.Tensor:
    """[groups*heads*d, ...] -> [groups*heads*dp, ...] with zero rows after each head's d rows."""
    rest = t.shape[1:]
    out = torch.zeros((groups, heads, dp) + rest, device=t.device, dtype=t.dtype)
    out[:, :, :d] = t.reshape((groups, heads, d) + rest)
    return out.reshape((groups * heads * dp,) + rest)


class SelfAttention(nn.Module):
    def __init__(self, dim, num_heads, causal=False, attn_dropout=0.0, proj_dropout=0.0):
        assert dim % num_heads == 0
        super().__init__()
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.causal, self.attn_dropout, self.proj_dropout = causal, attn_dropout, proj_dropout
        self.to_qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def _padded(self):
        n, d = self.num_heads, self.head_dim
        dp = padded_head_dim(d)
        wqkv = E.derived(self, "wqkv", lambda w: _splice_heads(w.detach().to(BF16), 3, n, d, dp).contiguous(), self.to_qkv.weight)
        bqkv = E.derived(self, "bqkv", lambda b: _splice_heads(b.detach().float(), 3, n, d, dp).contiguous(), self.to_qkv.bias)
        wproj = E.derived(self, "wproj", lambda w: _splice_heads(w.detach().to(BF16).t().contiguous(), 1, n, d, dp).t().contiguous(),
                          self.proj.weight)
        return dp, wqkv, bqkv, wproj

    def forward(self, x, resid=None):
        """x [B, L, C] -> [B, L, C] (+ resid, added in the projection's epilogue)."""
        assert not self.causal, "causal attention is only used by the (unmirrored) text tower"
        b, s, _ = x.shape
        n = self.num_heads
        dp, wqkv, bqkv, wproj = self._padded()
        qkv = ops.linear(E.as_bf16(x), wqkv, bias=bqkv).view(b, s, 3, n, dp)
        a = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale=1.0 / math.sqrt(self.head_dim))
        return ops.linear(a.view(b, s, n * dp), wproj, bias=E.f32(self.proj, "b", self.proj.bias), resid=resid,
                          round_flags=E.ROUND_AFTER_BIAS if resid is not None else 0)


class SwiGLU(nn.Module):
    def __init__(self, dim, mid_dim):
        super().__init__()
        self.dim, self.mid_dim = dim, mid_dim
        self.fc1 = nn.Linear(dim, mid_dim)
        self.fc2 = nn.Linear(dim, mid_dim)
        self.fc3 = nn.Linear(mid_dim, dim)

    def forward(self, x, resid=None):
        h = E.lin(x, self.fc1, act=E.ACT_SILU, round_flags=E.ROUND_AFTER_BIAS) * E.lin(x, self.fc2)
        return E.lin(h, self.fc3, resid=resid, round_flags=E.ROUND_AFTER_BIAS if resid is not None else 0)


def _mlp(seq: nn.Sequential, x, resid=None):
    """Linear -> (Quick)GELU -> Linear (-> Dropout): the activation is the first GEMM's epilogue where the kernel has it."""
    act = seq[1]
    if isinstance(act, nn.GELU):
        assert act.approximate == "none"
        h = E.lin(x, seq[0], act=E.ACT_GELU_ERF, round_flags=E.ROUND_AFTER_BIAS)
    else:
        h = act(E.lin(x, seq[0]))
    return E.lin(h, seq[2], resid=resid, round_flags=E.ROUND_AFTER_BIAS if resid is not None else 0)


class AttentionBlock(nn.Module):
    def __init__(self, dim, mlp_ratio, num_heads, post_norm=False, causal=False, activation="quick_gelu", attn_dropout=0.0,
                 proj_dropout=0.0, norm_eps=1e-5):
        assert activation in ("quick_gelu", "gelu", "swi_glu")
        super().__init__()
        self.dim, self.mlp_ratio, self.num_heads = dim, mlp_ratio, num_heads
        self.post_norm, self.causal, self.norm_eps = post_norm, causal, norm_eps
        self.norm1 = LayerNorm(dim, eps=norm_eps)
        self.attn = SelfAttention(dim, num_heads, causal, attn_dropout, proj_dropout)
        self.norm2 = LayerNorm(dim, eps=norm_eps)
        if activation == "swi_glu":
            self.mlp = SwiGLU(dim, int(dim * mlp_ratio))
        else:
            self.mlp = nn.Sequential(nn.Linear(dim, int(dim * mlp_ratio)), QuickGELU() if activation == "quick_gelu" else nn.GELU(),
                                     nn.Linear(int(dim * mlp_ratio), dim), nn.Dropout(proj_dropout))

    def _ffn(self, x, resid=None):
        return self.mlp(x, resid=resid) if isinstance(self.mlp, SwiGLU) else _mlp(self.mlp, x, resid)

    def forward(self, x):
        if self.post_norm:
            x = x + self.norm1(self.attn(x))
            return x + self.norm2(self._ffn(x))
        x = self.attn(self.norm1(x), resid=x)
        return self._ffn(self.norm2(x), resid=x)


class AttentionPool(nn.Module):
    def __init__(self, dim, mlp_ratio, num_heads, activation="gelu", proj_dropout=0.0, norm_eps=1e-5):
        assert dim % num_heads == 0
        super().__init__()
        self.dim, self.mlp_ratio, self.num_heads, self.head_dim = dim, mlp_ratio, num_heads, dim // num_heads
        self.proj_dropout, self.norm_eps = proj_dropout, norm_eps
        gain = 1.0 / math.sqrt(dim)
        self.cls_embedding = nn.Parameter(gain * torch.randn(1, 1, dim))
        self.to_q = nn.Linear(dim, dim)
        self.to_kv = nn.Linear(dim, dim * 2)
        self.proj = nn.Linear(dim, dim)
        self.norm = LayerNorm(dim, eps=norm_eps)
        self.mlp = nn.Sequential(nn.Linear(dim, int(dim * mlp_ratio)), QuickGELU() if activation == "quick_gelu" else nn.GELU(),
                                 nn.Linear(int(dim * mlp_ratio), dim), nn.Dropout(proj_dropout))

    def forward(self, x):
        """One learned query over the tokens (wan_video_image_encoder.py:363-383).  A single query row: the core is a [1, L]
        softmax per head and runs on torch; the projections and the MLP are fwb200 GEMMs."""
        b, s, c = x.shape
        n, d = self.num_heads, self.head_dim
        q = E.lin(E.as_bf16(self.cls_embedding), self.to_q).view(1, 1, n, d).expand(b, -1, -1, -1)
        kv = E.lin(E.as_bf16(x), self.to_kv).view(b, s, 2, n, d)
        a = F.scaled_dot_product_attention(q.transpose(1, 2), kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2))
        y = E.lin(a.transpose(1, 2).reshape(b, 1, c), self.proj)
        y = _mlp(self.mlp, self.norm(y), resid=y)
        return y[:, 0]


class VisionTransformer(nn.Module):
    def __init__(self, image_size=224, patch_size=16, dim=768, mlp_ratio=4, out_dim=512, num_heads=12, num_layers=12,
                 pool_type="token", pre_norm=True, post_norm=False, activation="quick_gelu", attn_dropout=0.0, proj_dropout=0.0,
                 embedding_dropout=0.0, norm_eps=1e-5):
        if image_size % patch_size != 0:
            print("[WARNING] image_size is not divisible by patch_size", flush=True)
        assert pool_type in ("token", "token_fc", "attn_pool")
        out_dim = out_dim or dim
        super().__init__()
        self.image_size, self.patch_size = image_size, patch_size
        self.num_patches = (image_size // patch_size) ** 2
        self.dim, self.mlp_ratio, self.out_dim = dim, mlp_ratio, out_dim
        self.num_heads, self.num_layers, self.pool_type = num_heads, num_layers, pool_type
        self.post_norm, self.norm_eps = post_norm, norm_eps

        gain = 1.0 / math.sqrt(dim)
        self.patch_embedding = nn.Conv2d(3, dim, kernel_size=patch_size, stride=patch_size, bias=not pre_norm)
        has_cls = pool_type in ("token", "token_fc")
        if has_cls:
            self.cls_embedding = nn.Parameter(gain * torch.randn(1, 1, dim))
        self.pos_embedding = nn.Parameter(gain * torch.randn(1, self.num_patches + (1 if has_cls else 0), dim))
        self.dropout = nn.Dropout(embedding_dropout)

        self.pre_norm = LayerNorm(dim, eps=norm_eps) if pre_norm else None
        self.transformer = nn.Sequential(*[AttentionBlock(dim, mlp_ratio, num_heads, post_norm, False, activation, attn_dropout,
                                                          proj_dropout, norm_eps) for _ in range(num_layers)])
        self.post_norm = LayerNorm(dim, eps=norm_eps)      # (the reference overwrites the bool with the module, :441)

        if pool_type == "token":
            self.head = nn.Parameter(gain * torch.randn(dim, out_dim))
        elif pool_type == "token_fc":
            self.head = nn.Linear(dim, out_dim)
        else:
            self.head = AttentionPool(dim, mlp_ratio, num_heads, activation, proj_dropout, norm_eps)

    def patches(self, x):
        """[B, 3, H, W] -> [B, (H/p)*(W/p), 3*p*p] in the convolution's (c, ky, kx) weight order."""
        b, c, h, w = x.shape
        p = self.patch_size
        gh, gw = h // p, w // p
        x = x[:, :, :gh * p, :gw * p].reshape(b, c, gh, p, gw, p)
        return x.permute(0, 2, 4, 1, 3, 5).reshape(b, gh * gw, c * p * p)

    @torch.no_grad()
    def forward(self, x, interpolation=False, use_31_block=False):
        ops.require_device()
        assert not self.training, "VisionTransformer mirror: inference only"
        b = x.size(0)
        x = E.lin(E.as_bf16(self.patches(x)), self.patch_embedding)
        if self.pool_type in ("token", "token_fc"):
            x = torch.cat([self.cls_embedding.expand(b, -1, -1).to(dtype=x.dtype, device=x.device), x], dim=1)
        e = pos_interpolate(self.pos_embedding, x.size(1)) if interpolation else self.pos_embedding
        x = x + e.to(dtype=x.dtype, device=x.device)
        if self.pre_norm is not None:
            x = self.pre_norm(x)
        blocks = self.transformer[:-1] if use_31_block else self.transformer
        for blk in blocks:
            x = blk(x)
        return x


class XLMRobertaCLIP(nn.Module):
    """The CLIP shell (wan_video_image_encoder.py:642-706): a vision tower, `textual = None`, `log_scale`."""

    def __init__(self, embed_dim=1024, image_size=224, patch_size=14, vision_dim=1280, vision_mlp_ratio=4, vision_heads=16,
                 vision_layers=32, vision_pool="token", vision_pre_norm=True, vision_post_norm=False, activation="gelu",
                 vocab_size=250002, max_text_len=514, type_size=1, pad_id=1, text_dim=1024, text_heads=16, text_layers=24,
                 text_post_norm=True, text_dropout=0.1, attn_dropout=0.0, proj_dropout=0.0, embedding_dropout=0.0, norm_eps=1e-5):
        super().__init__()
        self.embed_dim, self.image_size, self.patch_size = embed_dim, image_size, patch_size
        self.vision_dim, self.vision_mlp_ratio, self.vision_heads, self.vision_layers = (vision_dim, vision_mlp_ratio, vision_heads,
                                                                                         vision_layers)
        self.vision_pre_norm, self.vision_post_norm, self.activation = vision_pre_norm, vision_post_norm, activation
        self.vocab_size, self.max_text_len, self.type_size, self.pad_id = vocab_size, max_text_len, type_size, pad_id
        self.text_dim, self.text_heads, self.text_layers, self.text_post_norm = text_dim, text_heads, text_layers, text_post_norm
        self.norm_eps = norm_eps
        self.visual = VisionTransformer(image_size=image_size, patch_size=patch_size, dim=vision_dim, mlp_ratio=vision_mlp_ratio,
                                        out_dim=embed_dim, num_heads=vision_heads, num_layers=vision_layers, pool_type=vision_pool,
                                        pre_norm=vision_pre_norm, post_norm=vision_post_norm, activation=activation,
                                        attn_dropout=attn_dropout, proj_dropout=proj_dropout, embedding_dropout=embedding_dropout,
                                        norm_eps=norm_eps)
        self.textual = None
        self.log_scale = nn.Parameter(math.log(1 / 0.07) * torch.ones([]))

    def forward(self, imgs, txt_ids):
        raise NotImplementedError("the CLIP text tower is not part of the I2V path (textual is None in the reference as well)")


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class _Normalize(nn.Module):
    """Channel normalisation (the last entry of the reference's torchvision transform list, :783-790)."""

    def __init__(self, mean, std):
        super().__init__()
        self.mean, self.std = tuple(mean), tuple(std)

    def forward(self, x):
        mean = torch.tensor(self.mean, device=x.device, dtype=x.dtype).view(1, -1, 1, 1)
        std = torch.tensor(self.std, device=x.device, dtype=x.dtype).view(1, -1, 1, 1)
        return (x - mean) / std


def clip_xlm_roberta_vit_h_14(pretrained=False, pretrained_name="open-clip-xlm-roberta-large-vit-huge-14", return_transforms=False,
                              return_tokenizer=False, dtype=torch.float32, device="cpu", **kwargs):
    """Random-init ViT-H/14 CLIP shell (+ the normalisation transform).  `pretrained=True` (bucket download in the reference,
    :749-771) and the tokenizer are I/O and out of scope."""
    if pretrained or return_tokenizer:
        raise NotImplementedError("pretrained download / tokenizer: out of scope, load a state_dict instead")
    cfg = dict(embed_dim=1024, image_size=224, patch_size=14, vision_dim=1280, vision_mlp_ratio=4, vision_heads=16,
               vision_layers=32, vision_pool="token", activation="gelu", vocab_size=250002, max_text_len=514, type_size=1,
               pad_id=1, text_dim=1024, text_heads=16, text_layers=24, text_post_norm=True, text_dropout=0.1, attn_dropout=0.0,
               proj_dropout=0.0, embedding_dropout=0.0)
    cfg.update(**kwargs)
    with torch.device(device):
        model = XLMRobertaCLIP(**cfg)
    if return_transforms:
        return model, nn.Sequential(_Normalize(CLIP_MEAN, CLIP_STD))
    return model


class WanImageEncoder(nn.Module):
    def __init__(self, device="cpu", **clip_kwargs):
        """The reference takes no arguments (ViT-H/14 on the CPU); `device` ("meta" for checkpoint loading) and the CLIP config
        overrides (reduced towers for tests) are extensions."""
        super().__init__()
        self.model, self.transforms = clip_xlm_roberta_vit_h_14(pretrained=False, return_transforms=True, return_tokenizer=False,
                                                                dtype=torch.float32, device=device, **clip_kwargs)

    @torch.no_grad()
    def encode_image(self, videos):
        """videos: list of [1, 3, H, W] images in [-1, 1] -> [B, 1 + 16*16, 1280] penultimate-block tokens (:864-880)."""
        size = (self.model.image_size,) * 2
        imgs = torch.cat([F.interpolate(u, size=size, mode="bicubic", align_corners=False) for u in videos])
        imgs = self.transforms[-1](imgs * 0.5 + 0.5)
        dtype = next(iter(self.model.visual.parameters())).dtype
        return self.model.visual(imgs.to(dtype), use_31_block=True)

    @staticmethod
    def state_dict_converter():
        return WanImageEncoderStateDictConverter()


class WanImageEncoderStateDictConverter:
    def from_diffusers(self, state_dict):
        return state_dict

    def from_civitai(self, state_dict):
        """Drop the text tower, prefix the rest with `model.` (wan_video_image_encoder.py:894-901)."""
        return {"model." + k: v for k, v in state_dict.items() if not k.startswith("textual.")}
