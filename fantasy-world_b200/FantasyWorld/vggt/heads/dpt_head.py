"""Mirror of FantasyWorld/vggt/heads/dpt_head.py (reference): DPT dense-prediction head with a causal 4x temporal
up-sampler (21 latent frames -> 81 frames), used for the depth and world-point outputs on the LAST denoising step only
(SURVEY §8 a19).  Same state_dict keys; stage 1 (LayerNorm + 1x1 projections of the 2048-wide tokens) runs on the fwb200 kernels, the spatial
convolution / interpolation work stays on torch (cuDNN / ATen),
frame chunking (4 / 16) and every index selection follow the reference so the outputs line up element for element.
"""
from __future__ import annotations

from typing import List, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from FantasyWorld.wan.modules.vae_modified import WanVAE_

from .head_act import activate_head
from .utils import create_uv_grid, position_grid_to_embed


def custom_interpolate(x: torch.Tensor, size: Tuple[int, int] = None, scale_factor: float = None, mode: str = "bilinear",
                       align_corners: bool = True) -> torch.Tensor:
    """F.interpolate, split along the batch when the output would exceed the 32-bit index range.  ref: dpt_head.py:536-566."""
    if size is None:
        size = (int(x.shape[-2] * scale_factor), int(x.shape[-1] * scale_factor))
    limit = 1610612736
    total = size[0] * size[1] * x.shape[0] * x.shape[1]
    if total <= limit:
        return F.interpolate(x, size=size, mode=mode, align_corners=align_corners)
    parts = torch.chunk(x, chunks=total // limit + 1, dim=0)
    return torch.cat([F.interpolate(p, size=size, mode=mode, align_corners=align_corners) for p in parts], dim=0).contiguous()


class ResidualConvUnit(nn.Module):
    """relu(x) + conv2(relu(conv1(relu(x)))).  The reference's activation is nn.ReLU(inplace=True), so by the time the
    skip connection is added its input has already been rectified in place — the skip carries relu(x), not x.
    ref: dpt_head.py:395-451, 342-343."""

    def __init__(self, features, activation, bn, groups=1):
        super().__init__()
        self.bn, self.groups = bn, groups
        self.conv1 = nn.Conv2d(features, features, kernel_size=3, stride=1, padding=1, bias=True, groups=groups)
        self.conv2 = nn.Conv2d(features, features, kernel_size=3, stride=1, padding=1, bias=True, groups=groups)
        self.norm1 = None
        self.norm2 = None
        self.activation = activation

    def forward(self, x):
        r = F.relu(x)
        y = self.conv2(F.relu(self.conv1(r)))
        return y + r


class FeatureFusionBlock(nn.Module):
    """(coarse + RCU(skip)) -> RCU -> bilinear up-sample (align_corners) -> 1x1 conv.  ref: dpt_head.py:454-533."""

    def __init__(self, features, activation, deconv=False, bn=False, expand=False, align_corners=True, size=None,
                 has_residual=True, groups=1):
        super().__init__()
        self.deconv, self.align_corners, self.groups, self.expand = deconv, align_corners, groups, expand
        out_features = features // 2 if expand else features
        self.out_conv = nn.Conv2d(features, out_features, kernel_size=1, stride=1, padding=0, bias=True, groups=groups)
        if has_residual:
            self.resConfUnit1 = ResidualConvUnit(features, activation, bn, groups=groups)
        self.has_residual = has_residual
        self.resConfUnit2 = ResidualConvUnit(features, activation, bn, groups=groups)
        self.size = size

    def forward(self, *xs, size=None):
        y = xs[0]
        if self.has_residual:
            y = y + self.resConfUnit1(xs[1])
        y = self.resConfUnit2(y)
        if size is not None:
            y = custom_interpolate(y, size=size, mode="bilinear", align_corners=self.align_corners)
        elif self.size is not None:
            y = custom_interpolate(y, size=self.size, mode="bilinear", align_corners=self.align_corners)
        else:
            y = custom_interpolate(y, scale_factor=2, mode="bilinear", align_corners=self.align_corners)
        return self.out_conv(y)


def _make_fusion_block(features: int, size: int = None, has_residual: bool = True, groups: int = 1) -> nn.Module:
    return FeatureFusionBlock(features, nn.ReLU(inplace=True), deconv=False, bn=False, expand=False, align_corners=True,
                              size=size, has_residual=has_residual, groups=groups)


def _make_scratch(in_shape: List[int], out_shape: int, groups: int = 1, expand: bool = False) -> nn.Module:
    scratch = nn.Module()
    mult = [1, 2, 4, 8] if expand else [1, 1, 1, 1]
    for i, cin in enumerate(in_shape[:4]):
        setattr(scratch, f"layer{i + 1}_rn",
                nn.Conv2d(cin, out_shape * mult[i], kernel_size=3, stride=1, padding=1, bias=False, groups=groups))
    return scratch


class TokenLayerNorm(nn.LayerNorm):
    """nn.LayerNorm over the 2048-wide aggregated tokens (same parameters and state_dict keys).  On a CUDA tensor the fwb200 row
    kernel computes it (fp32 statistics, affine) and returns bf16 — under the reference's autocast the fp32 LayerNorm output is
    cast to bf16 by the convolution that consumes it, so the values entering the projection are the same."""

    def forward(self, x):
        if not x.is_cuda:
            return super().forward(x)
        from fwb200 import engine as E
        from fwb200 import ops
        rows = x.reshape(-1, x.shape[-1])
        if rows.dtype not in (torch.float32, torch.bfloat16):
            rows = rows.float()
        y = ops.ln_modulate(rows.contiguous(), eps=self.eps, w=E.f32(self, "w", self.weight), b=E.f32(self, "b", self.bias))
        return y.view(*x.shape)


class DPTHead_3D_Causal(nn.Module):
    """ref: dpt_head.py:13-320."""

    def __init__(self, dim_in: int, patch_size: int = 14, output_dim: int = 4, activation: str = "inv_log",
                 conf_activation: str = "expp1", features: int = 256, out_channels: List[int] = [256, 512, 1024, 1024],
                 intermediate_layer_idx: List[int] = [23, 17, 11, 7], pos_embed: bool = True, feature_only: bool = False,
                 down_ratio: int = 1, temporal_scale: int = 4) -> None:
        super().__init__()
        self.patch_size, self.activation, self.conf_activation = patch_size, activation, conf_activation
        self.pos_embed, self.feature_only, self.down_ratio = pos_embed, feature_only, down_ratio
        self.intermediate_layer_idx, self.temporal_scale = intermediate_layer_idx, temporal_scale
        self.norm = TokenLayerNorm(dim_in)
        self.projects = nn.ModuleList([nn.Conv2d(dim_in, oc, kernel_size=1) for oc in out_channels])
        self.resize_layers = nn.ModuleList([
            nn.ConvTranspose2d(out_channels[0], out_channels[0], kernel_size=4, stride=4, padding=0),
            nn.ConvTranspose2d(out_channels[1], out_channels[1], kernel_size=2, stride=2, padding=0),
            nn.Identity(),
            nn.Conv2d(out_channels[3], out_channels[3], kernel_size=3, stride=2, padding=1)])
        self.temporal_upsamplers = nn.ModuleList([WanVAE_(z_dim=oc, location="DPT") for oc in out_channels])
        self.scratch = _make_scratch(out_channels, features, expand=False)
        self.scratch.stem_transpose = None
        self.scratch.refinenet1 = _make_fusion_block(features)
        self.scratch.refinenet2 = _make_fusion_block(features)
        self.scratch.refinenet3 = _make_fusion_block(features)
        self.scratch.refinenet4 = _make_fusion_block(features, has_residual=False)
        if feature_only:
            self.scratch.output_conv1 = nn.Conv2d(features, features, kernel_size=3, stride=1, padding=1)
        else:
            self.scratch.output_conv1 = nn.Conv2d(features, features // 2, kernel_size=3, stride=1, padding=1)
            self.scratch.output_conv2 = nn.Sequential(nn.Conv2d(features // 2, 32, kernel_size=3, stride=1, padding=1),
                                                      nn.ReLU(inplace=True), nn.Conv2d(32, output_dim, kernel_size=1))

    # ---- positional embedding ------------------------------------------------------------------------------------------
    def _apply_pos_embed(self, x: torch.Tensor, W: int, H: int, ratio: float = 0.1) -> torch.Tensor:
        """ref: dpt_head.py:264-287."""
        ph, pw = x.shape[-2], x.shape[-1]
        grid = create_uv_grid(pw, ph, aspect_ratio=W / H, dtype=x.dtype, device=x.device)
        emb = position_grid_to_embed(grid, x.shape[1]) * ratio
        return x + emb.permute(2, 0, 1)[None].expand(x.shape[0], -1, -1, -1)

    # ---- stage 1: tokens -> 4 feature pyramids, per chunk of latent frames ------------------------------------------------
    def _tokens_to_pyramid(self, tokens_list, images, patch_start_idx, f0, f1):
        """ref: dpt_head.py:204-238."""
        B = images.shape[0]
        S = f1 - f0
        gh, gw = images.shape[2], images.shape[3]
        H, W = gh * self.patch_size, gw * self.patch_size
        feats = []
        for level, layer_idx in enumerate(self.intermediate_layer_idx):
            x = tokens_list[layer_idx][:, f0:f1, patch_start_idx:]
            x = self.norm(x.reshape(B * S, -1, x.shape[-1]))          # CUDA: fwb_ln_modulate, bf16 out (see TokenLayerNorm)
            if x.is_cuda:
                # the 1x1 projection as a tcgen05 GEMM over tokens [S*gh*gw, 2048] x [oc, 2048]^T, conv output rounded
                # to bf16 as under autocast; the token-major result is re-laid as NCHW for the convolutions behind it
                from fwb200 import engine as E
                from fwb200 import ops
                y = E.lin(E.as_bf16(x).reshape(-1, x.shape[-1]), self.projects[level], round_flags=ops.ROUND_AFTER_BIAS)
                x = y.view(B * S, gh, gw, -1).permute(0, 3, 1, 2).contiguous()
            else:   # CPU: plain torch, only used by the index / golden unit tests of the host logic
                x = x.permute(0, 2, 1).reshape(B * S, x.shape[-1], gh, gw)
                x = self.projects[level](x)
            if self.pos_embed:
                x = self._apply_pos_embed(x, W, H)
            feats.append(self.resize_layers[level](x))
        return feats

    # ---- stage 3: fusion + output convs, per chunk of video frames --------------------------------------------------------
    def scratch_forward(self, features: List[torch.Tensor]) -> torch.Tensor:
        """ref: dpt_head.py:289-320."""
        l1, l2, l3, l4 = features
        r1, r2, r3, r4 = (self.scratch.layer1_rn(l1), self.scratch.layer2_rn(l2), self.scratch.layer3_rn(l3),
                          self.scratch.layer4_rn(l4))
        y = self.scratch.refinenet4(r4, size=r3.shape[2:])
        y = self.scratch.refinenet3(y, r3, size=r2.shape[2:])
        y = self.scratch.refinenet2(y, r2, size=r1.shape[2:])
        y = self.scratch.refinenet1(y, r1)
        return self.scratch.output_conv1(y)

    def _pyramid_to_output(self, feats, images):
        """ref: dpt_head.py:240-262."""
        B, _, gh, gw, _ = images.shape
        H, W = gh * self.patch_size, gw * self.patch_size
        y = self.scratch_forward(feats)
        y = custom_interpolate(y, (int(H / self.down_ratio), int(W / self.down_ratio)), mode="bilinear", align_corners=True)
        if self.pos_embed:
            y = self._apply_pos_embed(y, W, H)
        if self.feature_only:
            return y.view(B, -1, *y.shape[1:])
        y = self.scratch.output_conv2(y)
        preds, conf = activate_head(y, activation=self.activation, conf_activation=self.conf_activation)
        n = y.shape[0] // B
        return preds.view(B, n, *preds.shape[1:]), conf.view(B, n, *conf.shape[1:])

    def forward(self, aggregated_tokens_list: List[torch.Tensor], images: torch.Tensor, patch_start_idx: int,
                frames_chunk_size_first: int = 4, frames_chunk_size_second: int = 16
                ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        """`images` is the [B, S, h, w, 1024] patch-token tensor (only its shape is used).  ref: dpt_head.py:133-202."""
        B, S = images.shape[0], images.shape[1]
        levels = [[] for _ in range(4)]
        for f0 in range(0, S, frames_chunk_size_first):
            f1 = min(f0 + frames_chunk_size_first, S)
            for lv, feat in enumerate(self._tokens_to_pyramid(aggregated_tokens_list, images, patch_start_idx, f0, f1)):
                levels[lv].append(feat.view(B, f1 - f0, *feat.shape[1:]).permute(0, 2, 1, 3, 4))
        # stage 2: causal temporal 4x up-sampling of every pyramid level: S -> 4 (S - 1) + 1 frames
        clips = [self.temporal_upsamplers[lv].decode(torch.cat(levels[lv], dim=2)) for lv in range(4)]
        n_out = (S - 1) * 4 + 1
        preds, confs = [], []
        for t0 in range(0, n_out, frames_chunk_size_second):
            t1 = min(t0 + frames_chunk_size_second, n_out)
            sub = [c[:, :, t0:t1].permute(0, 2, 1, 3, 4).reshape(-1, c.shape[1], c.shape[3], c.shape[4]) for c in clips]
            out = self._pyramid_to_output(sub, images)
            if self.feature_only:
                preds.append(out)
            else:
                preds.append(out[0])
                confs.append(out[1])
        if self.feature_only:
            return torch.cat(preds, dim=1)
        return torch.cat(preds, dim=1), torch.cat(confs, dim=1)
