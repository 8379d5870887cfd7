"""Synthetic weights and inputs (there are no checkpoints or datasets in this environment).

Pure torch: importing this module loads NO native library, so the CPU reference arm of bench.py, the reference runner
(oracle/ref_runner.py) and the golden generators can share the generators without mapping libfwb200.so.

`synth_init` fills every parameter of a module from a per-key seeded generator, so two models with the same
state_dict keys (the reference imported in the build container, and this repo's mirror) receive IDENTICAL weights
regardless of construction order — that is how the golden vectors under tests/golden/ are tied to the oracle and to the
CUDA path.  Zero / tiny initialised tensors of the reference (adapter gammas, LayerScale, camera-shift last layer, ...)
get O(1) or N(0, 0.02) values so that no sub-path is multiplied away.
"""
from __future__ import annotations

import math
import zlib

import torch
import torch.nn as nn


def _gen(key: str, seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return g


def synth_tensor(key: str, shape, seed: int = 0, device="cpu") -> torch.Tensor:
    """fp32 values for parameter `key` (rule chosen from the key name / rank)."""
    dev = torch.device(device)
    g = _gen(key, seed, dev)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]

    def randn(std=1.0):
        return torch.randn(shape, generator=g, device=dev, dtype=torch.float32) * std

    def rand(lo, hi):
        return torch.rand(shape, generator=g, device=dev, dtype=torch.float32) * (hi - lo) + lo

    if leaf in ("gamma_m1", "gamma_m2") or key.endswith("ls1.gamma") or key.endswith("ls2.gamma"):
        return rand(0.5, 1.5)
    if leaf == "gamma":                       # RMS_norm gains of the temporal up-samplers
        return 1.0 + randn(0.1)
    if leaf == "modulation":
        return randn(1.0 / math.sqrt(shape[-1]))
    if leaf in ("camera_token", "register_token", "empty_pose_tokens", "emb_pos"):
        return randn(0.02)
    if leaf == "bias":
        return randn(0.02)
    if leaf == "weight" and len(shape) == 1:  # norm gains
        return 1.0 + randn(0.1)
    if leaf == "weight":
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return randn(1.0 / math.sqrt(max(fan_in, 1)))
    return randn(0.02)


@torch.no_grad()
def synth_init(module: nn.Module, seed: int = 0, gen_device=None) -> nn.Module:
    """Overwrite every parameter of `module` in place.  gen_device='cpu' gives machine-independent values
    (tests, goldens); the default generates on the parameter's own device (fast path for the 14B benchmark)."""
    for key, p in module.named_parameters():
        dev = gen_device or p.device
        p.copy_(synth_tensor(key, p.shape, seed, dev).to(device=p.device, dtype=p.dtype))
    return module


def materialize(module: nn.Module, device, dtype=torch.bfloat16) -> nn.Module:
    """Give storage to a module built under torch.device('meta'): floating parameters / buffers become `dtype`."""
    for mod in module.modules():
        for name, p in list(mod._parameters.items()):
            if p is not None and p.is_meta:
                dt = dtype if p.is_floating_point() else p.dtype
                mod._parameters[name] = nn.Parameter(torch.empty(p.shape, device=device, dtype=dt), requires_grad=False)
        for name, b in list(mod._buffers.items()):
            if b is not None and b.is_meta:
                dt = dtype if b.is_floating_point() else b.dtype
                mod._buffers[name] = torch.zeros(b.shape, device=device, dtype=dt)
    return module


VGGT_CFG = dict(img_size=518, patch_size=16, embed_dim=1024, number_frame=81, freq_dim=256, enable_camera=True,
                enable_depth=True, enable_point=True, enable_track=False, DPT_patch_size=16)
CAMERA_CFG = dict(pose_in_dim=1024, plucker_fea_dim=2048, pose_inject_method="adaln", use_info="plucker")
WAN21_I2V_14B = dict(dim=5120, in_dim=36, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6,
                     patch_size=(1, 2, 2), num_heads=40, num_layers=40, has_image_input=True)



def synth_inputs(f: int, h: int, w: int, device="cuda", seed: int = 1024, text_len: int = 512, dtype=torch.bfloat16):
    """Synthetic sampler inputs for a latent grid of f x (2h) x (2w) (token grid f x h x w), SURVEY §8d:
    latents, y (4 mask + 16 latent channels), context_pos/neg, clip_feature, token-aligned camera features."""
    g = torch.Generator(device="cpu").manual_seed(seed)

    def rn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    H, W = 2 * h, 2 * w
    lat = rn(1, 16, f, H, W)
    mask = torch.zeros(1, 4, f, H, W)
    mask[:, :, 0] = 1.0
    y = torch.cat([mask, rn(1, 16, f, H, W)], dim=1)
    out = dict(latents=lat, y=y, context_pos=rn(1, text_len, 4096), context_neg=rn(1, text_len, 4096),
               clip_feature=rn(1, 257, 1280), plucker_fea=rn(1, f * h * w, 2048))
    return {k: v.to(device=device, dtype=dtype) for k, v in out.items()}


def synth_block_inputs(f: int, h: int, w: int, text_len: int = 512, seed: int = 1024):
    """Inputs for running single blocks in isolation (one PCB DiT block, one VGGT frame block, one IRG block) at a token
    grid f x h x w — SURVEY §8d "C1"-style inputs at any size.  fp32 CPU tensors holding bf16-representable values (so an
    fp32 run and a bf16 run see the same numbers); `e0` is genuinely fp32, as on the real path (vggt.py:126-130)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L, P = f * h * w, 5 + h * w

    def rn(*shape, s=1.0):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * s).to(torch.bfloat16).float()

    return dict(x_dit=rn(1, L, 5120), x_agg=rn(f, P, 1024), context=rn(1, 257 + text_len, 5120), t_mod=rn(1, 6, 5120, s=0.1),
                e0=torch.randn(1, 6, 1024, generator=g, dtype=torch.float32) * 0.1, plucker=rn(1, L, 2048))
