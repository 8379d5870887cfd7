"""Mirror of FantasyWorld/diffsynth_wan21/models/wan_video_vae.py (reference): the Wan 3-D causal VAE — what the sampler calls
AFTER the denoising loop, `pipe.vae.decode(latents, tiled=True, tile_size=(30, 52), tile_stride=(15, 26))`
(inference_wan21.py:324-330), plus the encoder (first-frame conditioning, N3).  Same class names, constructor
arguments and state_dict keys (`model.encoder.*`, `model.conv1/2.*`, `model.decoder.*`), so the released Wan2.1 VAE checkpoint
loads unchanged.

What differs is the execution plan, not the arithmetic:
  * The reference decodes ONE latent frame at a time and threads a per-convolution cache of the last two frames through ~30 causal
    convolutions (wan_video_vae.py:552-575, CACHE_T = 2).  That streaming scheme is algebraically a causal convolution over the whole
    clip (zero history in front), with two documented exceptions that are reproduced: the first frame by-passes every temporal
    up-sampling ('Rep' sentinel, :126-129) and the temporal up-sampler's history starts at frame 1.  Here every stage runs ONCE over
    the full clip: 21x fewer, 21x larger cuDNN launches.
  * `tiled_decode` keeps its task list, masks and accumulation ORDER (:643-694) but accumulates on the GPU instead of the host, and
    can shard the tiles over the ranks of a process group (round-robin; partial sums are all-reduced).  One rank: same op order as
    the reference.
Convolutions / interpolation stay on cuDNN / ATen (once per video; the hand-written kernels of this repo cover the 50-step loop).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ...wan.modules.vae_modified import CausalConv3d, RMS_norm

CACHE_T = 2


class Upsample(nn.Upsample):
    """nearest-exact 2x in fp32, cast back (the reference's bf16 work-around, wan_video_vae.py:73-79)."""

    def forward(self, x):
        return super().forward(x.float()).type_as(x)


def _per_frame(fn, x):
    """Apply a 2-D module to every frame of [b, c, t, h, w]."""
    b, c, t, h, w = x.shape
    y = fn(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))
    return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)


class Resample(nn.Module):
    """Spatial (and, for the 3-D modes, causal temporal) 2x re-sampling.  ref: wan_video_vae.py:82-174."""

    def __init__(self, dim, mode):
        assert mode in ('none', 'upsample2d', 'upsample3d', 'downsample2d', 'downsample3d')
        super().__init__()
        self.dim, self.mode = dim, mode
        if mode in ('upsample2d', 'upsample3d'):
            self.resample = nn.Sequential(Upsample(scale_factor=(2., 2.), mode='nearest-exact'), nn.Conv2d(dim, dim // 2, 3, padding=1))
            if mode == 'upsample3d':
                self.time_conv = CausalConv3d(dim, dim * 2, (3, 1, 1), padding=(1, 0, 0))
        elif mode in ('downsample2d', 'downsample3d'):
            self.resample = nn.Sequential(nn.ZeroPad2d((0, 1, 0, 1)), nn.Conv2d(dim, dim, 3, stride=(2, 2)))
            if mode == 'downsample3d':
                self.time_conv = CausalConv3d(dim, dim, (3, 1, 1), stride=(2, 1, 1), padding=(0, 0, 0))
        else:
            self.resample = nn.Identity()

    def forward(self, x):
        """Whole clip.  upsample3d: t -> 1 + 2 (t - 1) frames (frame 0 is not doubled and is not part of the history of the
        frames behind it); downsample3d: 1 + 4k -> 1 + 2k frames (frame 0 passes, the rest see one frame of history)."""
        b, c, t, h, w = x.shape
        if self.mode == 'upsample3d' and t > 1:
            y = self.time_conv(x[:, :, 1:]).view(b, 2, c, t - 1, h, w)
            y = torch.stack((y[:, 0], y[:, 1]), dim=3).reshape(b, c, 2 * (t - 1), h, w)      # (first, second) half-frames interleaved
            x = torch.cat([x[:, :, :1], y], dim=2)
        x = _per_frame(self.resample, x)
        if self.mode == 'downsample3d' and x.shape[2] > 1:
            # streaming rule (:160-172): frame 0 is kept; every later chunk is convolved (kernel 3, stride 2, no padding) together with
            # the last frame before it -> over the whole clip: a stride-2 convolution over frames 0.. whose outputs start at frame 2
            x = torch.cat([x[:, :, :1], self.time_conv(x)], dim=2)        # padding (0,0,0): no causal pad is added
        return x


class ResidualBlock(nn.Module):
    """ref: wan_video_vae.py:198-232."""

    def __init__(self, in_dim, out_dim, dropout=0.0):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.residual = nn.Sequential(RMS_norm(in_dim, images=False), nn.SiLU(), CausalConv3d(in_dim, out_dim, 3, padding=1),
                                      RMS_norm(out_dim, images=False), nn.SiLU(), nn.Dropout(dropout),
                                      CausalConv3d(out_dim, out_dim, 3, padding=1))
        self.shortcut = CausalConv3d(in_dim, out_dim, 1) if in_dim != out_dim else nn.Identity()

    def forward(self, x):
        return self.residual(x) + self.shortcut(x)


class AttentionBlock(nn.Module):
    """Single-head self-attention over the h*w positions of every frame.  ref: wan_video_vae.py:235-273."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.norm = RMS_norm(dim)
        self.to_qkv = nn.Conv2d(dim, dim * 3, 1)
        self.proj = nn.Conv2d(dim, dim, 1)
        nn.init.zeros_(self.proj.weight)

    def forward(self, x):
        def frame_attn(f):                                          # f [n, c, h, w]
            n, c, h, w = f.shape
            q, k, v = self.to_qkv(self.norm(f)).reshape(n, 1, 3 * c, h * w).permute(0, 1, 3, 2).contiguous().chunk(3, dim=-1)
            o = F.scaled_dot_product_attention(q, k, v)
            return self.proj(o.squeeze(1).permute(0, 2, 1).reshape(n, c, h, w))

        return x + _per_frame(frame_attn, x)


def _run(layers, x):
    for layer in layers:
        x = layer(x)
    return x


class Encoder3d(nn.Module):
    """ref: wan_video_vae.py:276-376."""

    def __init__(self, dim=128, z_dim=4, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[True, True, False],
                 dropout=0.0):
        super().__init__()
        self.dim, self.z_dim, self.dim_mult, self.num_res_blocks = dim, z_dim, dim_mult, num_res_blocks
        self.attn_scales, self.temperal_downsample = attn_scales, temperal_downsample
        dims = [dim * u for u in [1] + dim_mult]
        scale = 1.0
        self.conv1 = CausalConv3d(3, dims[0], 3, padding=1)
        downs = []
        for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
            for _ in range(num_res_blocks):
                downs.append(ResidualBlock(in_dim, out_dim, dropout))
                if scale in attn_scales:
                    downs.append(AttentionBlock(out_dim))
                in_dim = out_dim
            if i != len(dim_mult) - 1:
                downs.append(Resample(out_dim, mode='downsample3d' if temperal_downsample[i] else 'downsample2d'))
                scale /= 2.0
        self.downsamples = nn.Sequential(*downs)
        self.middle = nn.Sequential(ResidualBlock(out_dim, out_dim, dropout), AttentionBlock(out_dim), ResidualBlock(out_dim, out_dim, dropout))
        self.head = nn.Sequential(RMS_norm(out_dim, images=False), nn.SiLU(), CausalConv3d(out_dim, z_dim, 3, padding=1))

    def forward(self, x):
        return _run(self.head, _run(self.middle, _run(self.downsamples, self.conv1(x))))


class Decoder3d(nn.Module):
    """ref: wan_video_vae.py:379-481."""

    def __init__(self, dim=128, z_dim=4, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_upsample=[False, True, True],
                 dropout=0.0):
        super().__init__()
        self.dim, self.z_dim, self.dim_mult, self.num_res_blocks = dim, z_dim, dim_mult, num_res_blocks
        self.attn_scales, self.temperal_upsample = attn_scales, temperal_upsample
        dims = [dim * u for u in [dim_mult[-1]] + dim_mult[::-1]]
        scale = 1.0 / 2 ** (len(dim_mult) - 2)
        self.conv1 = CausalConv3d(z_dim, dims[0], 3, padding=1)
        self.middle = nn.Sequential(ResidualBlock(dims[0], dims[0], dropout), AttentionBlock(dims[0]), ResidualBlock(dims[0], dims[0], dropout))
        ups = []
        for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
            if i in (1, 2, 3):
                in_dim = in_dim // 2                                   # the up-sampler in front halved the channels
            for _ in range(num_res_blocks + 1):
                ups.append(ResidualBlock(in_dim, out_dim, dropout))
                if scale in attn_scales:
                    ups.append(AttentionBlock(out_dim))
                in_dim = out_dim
            if i != len(dim_mult) - 1:
                ups.append(Resample(out_dim, mode='upsample3d' if temperal_upsample[i] else 'upsample2d'))
                scale *= 2.0
        self.upsamples = nn.Sequential(*ups)
        self.head = nn.Sequential(RMS_norm(out_dim, images=False), nn.SiLU(), CausalConv3d(out_dim, 3, 3, padding=1))

    def forward(self, x):
        return _run(self.head, _run(self.upsamples, _run(self.middle, self.conv1(x))))


class VideoVAE_(nn.Module):
    """ref: wan_video_vae.py:492-596."""

    def __init__(self, dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2, attn_scales=[], temperal_downsample=[False, True, True],
                 dropout=0.0):
        super().__init__()
        self.dim, self.z_dim, self.dim_mult, self.num_res_blocks = dim, z_dim, dim_mult, num_res_blocks
        self.attn_scales, self.temperal_downsample = attn_scales, temperal_downsample
        self.temperal_upsample = temperal_downsample[::-1]
        self.encoder = Encoder3d(dim, z_dim * 2, dim_mult, num_res_blocks, attn_scales, self.temperal_downsample, dropout)
        self.conv1 = CausalConv3d(z_dim * 2, z_dim * 2, 1)
        self.conv2 = CausalConv3d(z_dim, z_dim, 1)
        self.decoder = Decoder3d(dim, z_dim, dim_mult, num_res_blocks, attn_scales, self.temperal_upsample, dropout)

    @staticmethod
    def _scale(scale, like):
        if isinstance(scale[0], torch.Tensor):
            return [s.to(dtype=like.dtype, device=like.device).view(1, -1, 1, 1, 1) for s in scale]
        return scale

    def encode(self, x, scale):
        """video [b, 3, 1 + 4k, H, W] in [-1, 1] -> normalised latent mean [b, z, 1 + k, H/8, W/8].  ref: :525-550."""
        mu, _ = self.conv1(self.encoder(x)).chunk(2, dim=1)
        s = self._scale(scale, mu)
        return (mu - s[0]) * s[1]

    def decode(self, z, scale):
        """normalised latents [b, z, t, h, w] -> video [b, 3, 1 + 4 (t-1), 8h, 8w].  ref: :552-575."""
        s = self._scale(scale, z)
        return self.decoder(self.conv2(z / s[1] + s[0]))

    def forward(self, x):
        raise NotImplementedError("training forward (reparameterised sampling) is outside the inference path")


class WanVideoVAE(nn.Module):
    """ref: wan_video_vae.py:599-787."""

    def __init__(self, z_dim=16):
        super().__init__()
        self.mean = torch.tensor([-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
                                  0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921])
        self.std = torch.tensor([2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
                                 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160])
        self.scale = [self.mean, 1.0 / self.std]
        self.model = VideoVAE_(z_dim=z_dim).eval().requires_grad_(False)
        self.upsampling_factor = 8
        self.z_dim = z_dim

    # ---- blending masks (ref: :621-641) --------------------------------------------------------------------------------------
    @staticmethod
    def build_1d_mask(length, left_bound, right_bound, border_width):
        x = torch.ones((length,))
        ramp = (torch.arange(border_width) + 1) / border_width
        if not left_bound:
            x[:border_width] = ramp
        if not right_bound:
            x[-border_width:] = torch.flip(ramp, dims=(0,))
        return x

    def build_mask(self, data, is_bound, border_width):
        H, W = data.shape[-2:]
        h = self.build_1d_mask(H, is_bound[0], is_bound[1], border_width[0])[:, None].expand(H, W)
        w = self.build_1d_mask(W, is_bound[2], is_bound[3], border_width[1])[None, :].expand(H, W)
        return torch.minimum(h, w)[None, None, None]

    @staticmethod
    def tile_tasks(H, W, tile_size, tile_stride) -> List[Tuple[int, int, int, int]]:
        """Tile rectangles (h0, h1, w0, w1) in the reference's order (ref: :648-656): stride steps, skipping a start whose
        predecessor already reaches the border."""
        (size_h, size_w), (stride_h, stride_w) = tile_size, tile_stride
        tasks = []
        for h in range(0, H, stride_h):
            if h - stride_h >= 0 and h - stride_h + size_h >= H:
                continue
            for w in range(0, W, stride_w):
                if w - stride_w >= 0 and w - stride_w + size_w >= W:
                    continue
                tasks.append((h, h + size_h, w, w + size_w))
        return tasks

    def tiled_decode(self, hidden_states, device, tile_size, tile_stride, group=None):
        """Overlapping latent tiles decoded independently and blended with linear ramps (ref: :643-694).  `group` (extension): a
        torch.distributed process group — tiles are dealt round-robin to its ranks and the two accumulators are all-reduced."""
        _, _, T, H, W = hidden_states.shape
        f = self.upsampling_factor
        tasks = self.tile_tasks(H, W, tile_size, tile_stride)
        rank, world = 0, 1
        if group is not None:
            import torch.distributed as dist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        out_T = T * 4 - 3
        dt = hidden_states.dtype
        weight = torch.zeros((1, 1, out_T, H * f, W * f), dtype=dt, device=device)
        values = torch.zeros((1, 3, out_T, H * f, W * f), dtype=dt, device=device)
        border = ((tile_size[0] - tile_stride[0]) * f, (tile_size[1] - tile_stride[1]) * f)
        for n, (h0, h1, w0, w1) in enumerate(tasks):
            if n % world != rank:
                continue
            tile = self.model.decode(hidden_states[:, :, :, h0:h1, w0:w1].to(device), self.scale)
            mask = self.build_mask(tile, is_bound=(h0 == 0, h1 >= H, w0 == 0, w1 >= W), border_width=border).to(dtype=dt, device=device)
            th, tw = h0 * f, w0 * f
            values[:, :, :, th:th + tile.shape[3], tw:tw + tile.shape[4]] += tile * mask
            weight[:, :, :, th:th + tile.shape[3], tw:tw + tile.shape[4]] += mask
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(values, group=group)
            dist.all_reduce(weight, group=group)
        return (values / weight).clamp_(-1, 1)

    def single_decode(self, hidden_state, device):
        return self.model.decode(hidden_state.to(device), self.scale).clamp_(-1, 1)

    def single_encode(self, video, device):
        return self.model.encode(video.to(device), self.scale)

    def tiled_encode(self, video, device, tile_size, tile_stride):
        """The encoder on overlapping PIXEL tiles (sizes in pixels), blended on the latent grid with the same linear ramps as the tiled
        decode (ref: :695-743; used by the Wan2.2 conditioning call, inference_wan22.py:345-351 with tiled=True).  Accumulates on
        `device` (the reference accumulates on the CPU: same additions in the same order)."""
        _, _, T, H, W = video.shape
        f = self.upsampling_factor
        dt = video.dtype
        out_T = (T + 3) // 4
        z = self.model.z_dim
        weight = torch.zeros((1, 1, out_T, H // f, W // f), dtype=dt, device=device)
        values = torch.zeros((1, z, out_T, H // f, W // f), dtype=dt, device=device)
        border = ((tile_size[0] - tile_stride[0]) // f, (tile_size[1] - tile_stride[1]) // f)
        for h0, h1, w0, w1 in self.tile_tasks(H, W, tile_size, tile_stride):
            tile = self.model.encode(video[:, :, :, h0:h1, w0:w1].to(device), self.scale)
            mask = self.build_mask(tile, is_bound=(h0 == 0, h1 >= H, w0 == 0, w1 >= W), border_width=border).to(dtype=dt, device=device)
            th, tw = h0 // f, w0 // f
            values[:, :, :, th:th + tile.shape[3], tw:tw + tile.shape[4]] += tile * mask
            weight[:, :, :, th:th + tile.shape[3], tw:tw + tile.shape[4]] += mask
        return values / weight

    def encode(self, videos: Sequence[torch.Tensor], device, tiled=False, tile_size=(34, 34), tile_stride=(18, 16)):
        """videos: iterable of [3, T, H, W] -> [n, 16, 1 + (T-1)/4, H/8, W/8]; `tile_size` / `tile_stride` in latent cells.  ref: :758-774."""
        f = self.upsampling_factor
        out = []
        for v in videos:
            v = v.unsqueeze(0)
            if tiled:
                z = self.tiled_encode(v, device, (tile_size[0] * f, tile_size[1] * f), (tile_stride[0] * f, tile_stride[1] * f))
            else:
                z = self.single_encode(v, device)
            out.append(z.squeeze(0))
        return torch.stack(out)

    def decode(self, hidden_states, device, tiled=False, tile_size=(34, 34), tile_stride=(18, 16), group=None):
        """ref: :776-783 (call site inference_wan21.py:324-330)."""
        if tiled:
            return self.tiled_decode(hidden_states, device, tile_size, tile_stride, group=group)
        return self.single_decode(hidden_states, device)

    @staticmethod
    def state_dict_converter():
        return WanVideoVAEStateDictConverter()


class WanVideoVAEStateDictConverter:
    """ref: :789-800 — `Wan2.1_VAE.pth` holds the bare VideoVAE_ keys (optionally under 'model_state'): prefix with `model.`."""

    def from_civitai(self, state_dict):
        inner = state_dict.get("model_state", state_dict)
        return {"model." + k: v for k, v in inner.items()}
