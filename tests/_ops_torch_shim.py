"""TEST-ONLY torch restatement of the four fwb200 ops the encoder mirrors call (linear, attention, ln_modulate, rmsnorm_rope_).

Purpose: check the HOST LOGIC of the conditioning front-end mirrors (weight splicing for head_dim 80 -> 96, fused q|k|v weights,
residual / rounding wiring, mask folding) against the reference goldens in the CPU suite, and — on the GPU — give a same-weights
fp32-accumulating yardstick for the real kernels at full model size.  It follows the arithmetic include/fwb200.h documents for
each entry point (fp32 accumulation, bf16 rounding where `round_flags` says so, bf16 output).  Like oracle/, it is test
infrastructure: nothing under fantasy-world_b200/ imports it, and the product path has no CPU fallback.

    with torch_ops():          # monkeypatches fwb200.ops inside the block only
        y = mirror(x)
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def _r(t):
    return t.to(BF16).float()


def linear(x, w, *, bias=None, act=0, scale1=None, shift1=None, scale2=None, resid=None, out=None, out_dtype=BF16, round_flags=0):
    from fwb200 import ops
    assert x.dtype == BF16 and w.dtype == BF16, "fwb_gemm_bf16 takes bf16 operands"
    y = x.float() @ w.float().t()
    if bias is not None:
        y = y + bias.float()
    if round_flags & ops.ROUND_AFTER_BIAS:
        y = _r(y)
    if act == ops.ACT_GELU_TANH:
        y = F.gelu(y, approximate="tanh")
    elif act == ops.ACT_GELU_ERF:
        y = F.gelu(y)
    elif act == ops.ACT_RELU:
        y = F.relu(y)
    elif act == ops.ACT_SILU:
        y = F.silu(y)
    if round_flags & ops.ROUND_AFTER_ACT:
        y = _r(y)
    if scale1 is not None:
        y = y * scale1.float()
    if shift1 is not None:
        y = y + shift1.float()
    if round_flags & ops.ROUND_AFTER_AFFINE:
        y = _r(y)
    if scale2 is not None:
        y = y * scale2.float()
    if resid is not None:
        y = y + resid.float().reshape(y.shape)
    y = y.to(out_dtype)
    if out is not None:
        out.view(y.shape).copy_(y)
        return out
    return y


def attention(q, k, v, *, scale=None, out=None, accumulate=False):
    assert q.dtype == BF16 and q.dim() == 4          # [B, L, H, D]
    scale = 1.0 / math.sqrt(q.shape[-1]) if scale is None else scale
    s = torch.einsum("blhd,bmhd->bhlm", q.float(), k.float()) * scale
    o = torch.einsum("bhlm,bmhd->blhd", torch.softmax(s, dim=-1), v.float()).contiguous()
    if out is None:
        return o.to(BF16)
    out.copy_((o + out.float()) if accumulate else o)
    return out


def ln_modulate(x, *, eps, w=None, b=None, mul=None, add=None, out=None):
    xf = x.float()
    y = (xf - xf.mean(-1, keepdim=True)) * torch.rsqrt(xf.var(-1, unbiased=False, keepdim=True) + eps)
    if w is not None:
        y = y * w.float() + b.float()
    if mul is not None:
        y = y * mul.float()
    if add is not None:
        y = y + add.float()
    y = y.to(BF16)
    if out is not None:
        out.view(y.shape).copy_(y)
        return out
    return y


def rmsnorm_rope_(x, *, w=None, eps=1e-6, cos_sin=None, head_dim=0):
    assert cos_sin is None, "shim: the encoders use the RMSNorm part only"
    xf = x.float()
    if w is not None:
        rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
        x.copy_(_r(_r(xf * rstd) * w.float()))
    return x


@contextlib.contextmanager
def torch_ops():
    from fwb200 import ops
    saved = {n: getattr(ops, n) for n in ("linear", "attention", "ln_modulate", "rmsnorm_rope_", "require_device")}
    ops.linear, ops.attention, ops.ln_modulate, ops.rmsnorm_rope_ = linear, attention, ln_modulate, rmsnorm_rope_
    ops.require_device = lambda: None
    try:
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
